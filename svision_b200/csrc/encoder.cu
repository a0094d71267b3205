// Segment-pair -> 227x227x3 similarity image, bit-exact with the reference's rasteriser.
//
// Replaces (reference paths): src/network/create_batch.py:103-152 (row -> Segments -> image ->
// float32 - mean) and src/segmentplot/plot_segment.py:9-73 (scale ratio, cv2.line per segment,
// overlap channel), including OpenCV's cv::line = clipLine + 8-connected LineIterator with
// leftToRight=true (third-party; behaviour pinned by tests/golden/encoder_golden.npz).
//
// One CTA per image (grid-stride).  The image is a 3-bit-per-pixel bitmap, so it
// is built as three 227x256-bit planes in shared memory:
//   1. two threads do the fp64 end-point scaling and the Cohen-Sutherland clip of one segment
//      each (IEEE double division + truncation, exactly as Python/OpenCV do it);
//   2. every pixel of a Bresenham line has a closed form (# minor-axis steps before pixel i =
//      floor((2*dy*i + dx - 1) / (2*dx))), so the pixels are drawn in parallel with atomicOr;
//   3. channel 1 = channel 0 AND (columns holding >= 2 pixels): a carry-save "ones/twos"
//      reduction over the rows, 32 columns per word;
//   4. the CTA streams the image to HBM: a branch-free pass of 16-byte coalesced stores of the
//      background pattern, then one scalar store per lit channel-pixel (see write_nhwc).
// Output layouts: NHWC fp32 (what the reference materialises), NHWC fp16 (same values, lossless)
// and the conv1 operand layout used by the fused path: space-to-depth 4x4 -> [57*57][64] fp16
// (48 real channels (dy*4+dx)*3+c, 16 zero), see gemm layouts in DESIGN.md.
#include "common.cuh"
#include "encoder_bitmap.cuh"
#include "kernels.h"

namespace svx {

namespace {

using namespace bitmap;

// ---- value helpers --------------------------------------------------------------------------
template <typename T> struct Levels;
template <> struct Levels<float> {
    static __device__ __forceinline__ float get(int ch, bool lit) {
        return ch == 0 ? (lit ? 151.f : -104.f) : ch == 1 ? (lit ? 138.f : -117.f)
                                                          : (lit ? 131.f : -124.f);
    }
};
// fp16 bit patterns: 151=0x58B8 -104=0xD680 138=0x5850 -117=0xD750 131=0x5818 -124=0xD7C0
template <> struct Levels<__half> {
    static __device__ __forceinline__ uint32_t get(int ch, bool lit) {
        return ch == 0 ? (lit ? 0x58B8u : 0xD680u) : ch == 1 ? (lit ? 0x5850u : 0xD750u)
                                                             : (lit ? 0x5818u : 0xD7C0u);
    }
};

__device__ __forceinline__ bool bm_bit(const uint32_t* bm, int ch, int r, int c) {
    return (bm[ch * PLANE + r * BMW + (c >> 5)] >> (c & 31)) & 1u;
}

// 4 consecutive pixels (row-major, wrapping to the next image row) starting at (r, c)
__device__ __forceinline__ uint32_t window4(const uint32_t* plane, int r, int c) {
    const uint32_t* rowp = plane + r * BMW;
    const int wi = c >> 5, sh = c & 31;
    const uint32_t lo = rowp[wi];
    const uint32_t hi = wi < BMW - 1 ? rowp[wi + 1] : 0u;
    uint32_t w = __funnelshift_r(lo, hi, sh) & 0xFu;
    if (c > IMG - 4) w |= (rowp[BMW] << (IMG - c)) & 0xFu;   // row r+1 (row 227 is all zero)
    return w;
}

// element e (flat NHWC index inside the image) -> value
template <typename T>
__device__ __forceinline__ T scalar_value(const uint32_t* bm, int e) {
    const int p = e / 3, ch = e - 3 * p;
    const int r = p / IMG, c = p - r * IMG;
    const bool lit = bm_bit(bm, ch, r, c);
    if constexpr (sizeof(T) == 4) {
        return Levels<float>::get(ch, lit);
    } else {
        return __ushort_as_half((unsigned short)Levels<__half>::get(ch, lit));
    }
}

// One 16-byte vector of the NHWC stream starting at element e (phase ph = e % 3), given the
// 12-bit lit mask L (bit pix*3+ch for the 4 pixels starting at pixel e/3).
template <typename T>
__device__ __forceinline__ uint4 nhwc_vector(int ph, uint32_t L) {
    uint4 out;
    if constexpr (sizeof(T) == 4) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = ph + j;
            const int ch = k >= 3 ? k - 3 : k;                 // k in [0, 5]
            v[j] = Levels<float>::get(ch, (L >> k) & 1u);
        }
        out = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
                         __float_as_uint(v[3]));
    } else {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = ph + j;                              // k in [0, 9]
            const int ch = k - 3 * (k / 3);
            v[j] = Levels<__half>::get(ch, (L >> k) & 1u);
        }
        out = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16),
                         v[6] | (v[7] << 16));
    }
    return out;
}

// The image is > 98 % background, so it is written in two passes: (1) a branch-free stream of
// 16-byte vectors holding the periodic background pattern (period 3 elements), (2) after a CTA
// barrier, one scalar store per lit channel-pixel (~160 per image), which merges in L2 with the
// line written microseconds earlier.  The stream loop is ~12 instructions per 512-byte warp
// store; the first version decided lit/background per vector (~120 instructions) and was
// instruction-bound at 27-50 % of the HBM roofline.
template <typename T>
__device__ __forceinline__ T level_value(int ch, bool lit) {
    if constexpr (sizeof(T) == 4) return Levels<float>::get(ch, lit);
    else return __ushort_as_half((unsigned short)Levels<__half>::get(ch, lit));
}

template <typename T, int NT>
__device__ void write_nhwc(const uint32_t* bm, T* __restrict__ out_all, long long img) {
    constexpr int ENC_THREADS = NT;
    constexpr int EPV = 16 / (int)sizeof(T);                    // elements per 16-byte vector
    const int tid = threadIdx.x;
    const long long e_begin = img * (long long)NEL;
    const long long e_end = e_begin + NEL;
    // the vector stream starts on a 128-byte line so every 512-byte warp store is 4 whole lines
    constexpr int EPL = 128 / (int)sizeof(T);                   // elements per cache line
    const long long v_first = ((e_begin + EPL - 1) / EPL) * (EPL / EPV);
    const long long v_last = e_end / EPV;                       // exclusive
    const int head = (int)(v_first * EPV - e_begin);            // < EPL <= 64 scalar elements
    const int tail = (int)(e_end - v_last * EPV);
    const int nvec = (int)(v_last - v_first);

    // ---- pass 1: background everywhere ----
    if (tid < head) out_all[e_begin + tid] = level_value<T>(tid % 3, false);
    if (tid >= 64 && tid < 64 + tail) {
        const int e = NEL - tail + (tid - 64);
        out_all[e_begin + e] = level_value<T>(e % 3, false);
    }
    const uint4 bg0 = nhwc_vector<T>(0, 0u), bg1 = nhwc_vector<T>(1, 0u), bg2 = nhwc_vector<T>(2, 0u);
    constexpr int DPH = (ENC_THREADS * EPV) % 3;                // phase step of the thread's stride
    uint4* __restrict__ outv = reinterpret_cast<uint4*>(out_all) + v_first;
    int ph = (head + tid * EPV) % 3;
#pragma unroll 4
    for (int i = tid; i < nvec; i += ENC_THREADS) {
        outv[i] = ph == 0 ? bg0 : (ph == 1 ? bg1 : bg2);
        ph += DPH;
        if (ph >= 3) ph -= 3;
    }
    __syncthreads();                                            // background lands before the patches

    // ---- pass 2: lit pixels (channel 0 plane drives; channels 1 and 2 are subsets of it) ----
    T* __restrict__ o = out_all + e_begin;
    for (int w = tid; w < IMG * BMW; w += ENC_THREADS) {
        uint32_t bits = bm[w];
        if (!bits) continue;
        const uint32_t b1 = bm[PLANE + w], b2 = bm[2 * PLANE + w];
        const int r = w >> 3, cbase = (w & 7) << 5;
        while (bits) {
            const int k = __ffs(bits) - 1;
            bits &= bits - 1;
            const int e = 3 * (r * IMG + cbase + k);
            o[e] = level_value<T>(0, true);
            if ((b1 >> k) & 1u) o[e + 1] = level_value<T>(1, true);
            if ((b2 >> k) & 1u) o[e + 2] = level_value<T>(2, true);
        }
    }
}

// conv1 operand layout: [57*57 s2d pixels][64 ch] fp16, ch = (dy*4+dx)*3 + c, 48..63 zero.
template <int NT>
__device__ void write_s2d(const uint32_t* bm, __half* __restrict__ out_all, long long img) {
    constexpr int ENC_THREADS = NT;
    constexpr int S2D = 57;
    const int tid = threadIdx.x;
    uint4* __restrict__ outv =
        reinterpret_cast<uint4*>(out_all + img * (long long)(S2D * S2D * 64));
    const int q = tid & 7;                       // channel octet handled by this thread (fixed)
    for (int v = tid; v < S2D * S2D * 8; v += ENC_THREADS) {
        const int sp = v >> 3;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (q < 6) {
            const int Y = sp / S2D, X = sp - Y * S2D;
            const int c0 = 4 * X, wi = c0 >> 5, sh = c0 & 31;   // nibble never straddles a word
            uint32_t m[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                uint32_t acc = 0;
#pragma unroll
                for (int dy = 0; dy < 4; ++dy)
                    acc |= ((bm[ch * PLANE + (4 * Y + dy) * BMW + wi] >> sh) & 0xFu) << (4 * dy);
                m[ch] = acc;                                    // bit dy*4+dx
            }
            uint32_t h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = 8 * q + j;                        // (dy*4+dx)*3 + c
                const int d = k / 3, ch = k - 3 * d;
                const uint32_t mm = ch == 0 ? m[0] : ch == 1 ? m[1] : m[2];
                h[j] = Levels<__half>::get(ch, (mm >> d) & 1u);
            }
            // pixels of image row/col 227 do not exist: their conv1 weights are zero, emit 0
            o = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16),
                           h[6] | (h[7] << 16));
            if (Y == S2D - 1 || X == S2D - 1) {
                uint32_t hh[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = 8 * q + j, d = k / 3;
                    const bool pad = (Y == S2D - 1 && (d >> 2) == 3) || (X == S2D - 1 && (d & 3) == 3);
                    hh[j] = pad ? 0u : h[j];
                }
                o = make_uint4(hh[0] | (hh[1] << 16), hh[2] | (hh[3] << 16), hh[4] | (hh[5] << 16),
                               hh[6] | (hh[7] << 16));
            }
        }
        outv[v] = o;
    }
}

// Threads per CTA (= per image) are chosen so that the images being written concurrently span
// ~90 MB: with ~1200 small CTAs each streaming its own image the write window (370-730 MB) exceeds
// the TLB reach (~256 MB) and the store path stalls at 4.1-4.5 TB/s; measured 5.9 TB/s (fp16,
// 512 threads x 2 CTAs/SM) and 6.2 TB/s (fp32, 1024 threads x 1 CTA/SM).
template <int MODE> struct EncCfg { static constexpr int NT = 512, CTAS = 2; };
template <> struct EncCfg<0> { static constexpr int NT = 1024, CTAS = 1; };

template <int MODE>   // 0: NHWC f32, 1: NHWC f16, 2: conv1 operand (s2d f16)
__global__ void __launch_bounds__(EncCfg<MODE>::NT, EncCfg<MODE>::CTAS)
encode_kernel(const int32_t* __restrict__ rows, long long n, void* __restrict__ out) {
    __shared__ __align__(16) uint32_t bm[3 * PLANE];
    __shared__ LineParams lines[2];
    constexpr int ENC_THREADS = EncCfg<MODE>::NT;
    __shared__ uint32_t red[(ENC_THREADS / 32) * 8 * 2];
    __shared__ uint32_t colmask[8];
    // The 48-byte row of the NEXT image is fetched while the current image streams out: a global
    // load issued behind a saturated store queue takes microseconds, and the line set-up (and with
    // it the whole CTA, at the barrier) would otherwise wait for it on every image.
    __shared__ int32_t rowbuf[2][12];
    if (threadIdx.x < 12 && blockIdx.x < n) rowbuf[0][threadIdx.x] = rows[(long long)blockIdx.x * 12 + threadIdx.x];
    __syncthreads();
    int cur = 0;
    for (long long img = blockIdx.x; img < n; img += gridDim.x) {
        const long long nxt = img + gridDim.x;
        int32_t pre = 0;
        if (threadIdx.x < 12 && nxt < n) pre = __ldg(rows + nxt * 12 + threadIdx.x);
        build_bitmap<ENC_THREADS>(rowbuf[cur], bm, lines, red, colmask);
        if constexpr (MODE == 0) write_nhwc<float, ENC_THREADS>(bm, reinterpret_cast<float*>(out), img);
        if constexpr (MODE == 1) write_nhwc<__half, ENC_THREADS>(bm, reinterpret_cast<__half*>(out), img);
        if constexpr (MODE == 2) write_s2d<ENC_THREADS>(bm, reinterpret_cast<__half*>(out), img);
        if (threadIdx.x < 12) rowbuf[cur ^ 1][threadIdx.x] = pre;
        cur ^= 1;
        __syncthreads();
    }
}

}  // namespace

int launch_encode(const int32_t* rows_dev, long long n, void* out, int mode, int num_sms,
                  cudaStream_t stream) {
    if (n <= 0) return 0;
    if (mode < 0 || mode > 2) return fail(-1, "launch_encode: bad mode");
    // One wave of resident CTAs, each looping over images: the grid is SMs x (CTAs that really fit
    // per SM with the maximum shared-memory carveout), so the grid-stride loop has no ragged
    // second wave (a fixed 8 x SMs grid ran as 1.6 waves: only 5 CTAs were resident).
    static int blocks_cache[64][3] = {};             // per device (function attributes are per device)
    int dev = 0;
    cudaGetDevice(&dev);
    int* blocks_per_sm = blocks_cache[(dev >= 0 && dev < 64) ? dev : 0];
    if (blocks_per_sm[mode] == 0) {
        const void* fn = mode == 0 ? (const void*)encode_kernel<0>
                       : mode == 1 ? (const void*)encode_kernel<1> : (const void*)encode_kernel<2>;
        cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        int nb = 0;
        const int nt = mode == 0 ? EncCfg<0>::NT : mode == 1 ? EncCfg<1>::NT : EncCfg<2>::NT;
        const int want = mode == 0 ? EncCfg<0>::CTAS : mode == 1 ? EncCfg<1>::CTAS : EncCfg<2>::CTAS;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, nt, 0) != cudaSuccess || nb < 1) nb = 1;
        blocks_per_sm[mode] = nb < want ? nb : want;
    }
    long long blocks = (long long)num_sms * blocks_per_sm[mode];
    if (blocks > n) blocks = n;
    switch (mode) {
        case 0: encode_kernel<0><<<(unsigned)blocks, EncCfg<0>::NT, 0, stream>>>(rows_dev, n, out); break;
        case 1: encode_kernel<1><<<(unsigned)blocks, EncCfg<1>::NT, 0, stream>>>(rows_dev, n, out); break;
        default: encode_kernel<2><<<(unsigned)blocks, EncCfg<2>::NT, 0, stream>>>(rows_dev, n, out); break;
    }
    SVX_LAUNCH_CHECK("encode_kernel");
    return 0;
}

}  // namespace svx
