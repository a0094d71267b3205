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
//   4. the CTA streams the image to HBM with 16-byte coalesced stores.  Background vectors (98 %
//      of them) take a fast path: no per-element work.
// Output layouts: NHWC fp32 (what the reference materialises), NHWC fp16 (same values, lossless)
// and the conv1 operand layout used by the fused path: space-to-depth 4x4 -> [57*57][64] fp16
// (48 real channels (dy*4+dx)*3+c, 16 zero), see gemm layouts in DESIGN.md.
#include "common.cuh"
#include "kernels.h"

namespace svx {

namespace {

constexpr int IMG = 227;
constexpr int NPIX = IMG * IMG;          // 51529
constexpr int NEL = NPIX * 3;            // 154587 elements per image
constexpr int BMW = 8;                   // 32-bit words per bitmap row (256 >= 227 columns)
constexpr int BMROWS = 228;              // +1 all-zero row (space-to-depth pad row / straddle reads)
constexpr int PLANE = BMROWS * BMW;      // words per channel plane
constexpr int ENC_THREADS = 256;

struct LineParams {
    int x1, y1, dx, dy, sy, vert, count, rev;
};

__device__ __forceinline__ bool clip_line(long long& x1, long long& y1, long long& x2,
                                          long long& y2) {
    const long long right = IMG - 1, bottom = IMG - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// Segment s of a packed row -> draw parameters (count == 0 when the line is rejected).
__device__ __forceinline__ LineParams setup_line(const int32_t* __restrict__ row, int s) {
    const int la = row[10], lb = row[11];
    double ratio = (double)(la > lb ? la : lb) / 227.0;          // plot_segment.py:12
    if (ratio < 1.0) ratio = 1.0;                                // plot_segment.py:14-15
    const int32_t* g = row + 5 * s;
    const long long xs = g[0], ys = g[2], ye = g[3];
    const bool fwd = g[4] == 1;
    const long long len = ye - ys;                               // create_batch.py:118,132
    const long long xe = fwd ? xs + (len - 1) : xs - (len - 1);  // segmentplot/classes.py:50-53
    const long long ye2 = ys + (len - 1);                        // segmentplot/classes.py:54
    // (col, row) = (ref, read); int(v / ratio): fp64 division, truncation toward zero
    long long sx = (long long)((double)ys / ratio), sy_ = (long long)((double)xs / ratio);
    long long ex = (long long)((double)ye2 / ratio), ey = (long long)((double)xe / ratio);
    long long x1, y1, x2, y2;
    if (fwd) { x1 = sx; y1 = sy_; x2 = ex; y2 = ey; }            // plot_segment.py:46-47
    else     { x1 = ex; y1 = ey; x2 = sx; y2 = sy_; }            // plot_segment.py:49-52
    LineParams L;
    L.rev = fwd ? 0 : 1;
    L.count = 0;
    L.x1 = L.y1 = L.dx = L.dy = L.vert = 0;
    L.sy = 1;
    if (x1 < 0 || x1 >= IMG || y1 < 0 || y1 >= IMG || x2 < 0 || x2 >= IMG || y2 < 0 || y2 >= IMG) {
        if (!clip_line(x1, y1, x2, y2)) return L;
    }
    int dx = (int)(x2 - x1), dy = (int)(y2 - y1);
    int px = (int)x1, py = (int)y1;
    if (dx < 0) { dx = -dx; dy = -dy; px = (int)x2; py = (int)y2; }   // leftToRight
    if (dy < 0) { dy = -dy; L.sy = -1; }
    L.vert = dy > dx;
    if (L.vert) { int t = dx; dx = dy; dy = t; }
    L.x1 = px; L.y1 = py; L.dx = dx; L.dy = dy;
    L.count = dx + 1;
    return L;
}

// Builds the three bit planes of one image in shared memory.  All threads of the CTA call it.
__device__ void build_bitmap(const int32_t* __restrict__ row, uint32_t* bm, LineParams* lines,
                             uint32_t* red /* [8 warps][8 words][2] */, uint32_t* colmask) {
    const int tid = threadIdx.x;
    // zero the planes (3*228*8 words = 1368 uint4)
    uint4* bz = reinterpret_cast<uint4*>(bm);
    for (int i = tid; i < 3 * PLANE / 4; i += ENC_THREADS) bz[i] = make_uint4(0, 0, 0, 0);
    if (tid < 2) lines[tid] = setup_line(row, tid);
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const LineParams L = lines[s];
        for (int i = tid; i < L.count; i += ENC_THREADS) {
            const int st = L.dx > 0 ? (2 * L.dy * i + L.dx - 1) / (2 * L.dx) : 0;
            const int c = L.vert ? L.x1 + st : L.x1 + i;
            const int r = L.vert ? L.y1 + L.sy * i : L.y1 + L.sy * st;
            const uint32_t bit = 1u << (c & 31);
            atomicOr(&bm[r * BMW + (c >> 5)], bit);
            if (L.rev) atomicOr(&bm[2 * PLANE + r * BMW + (c >> 5)], bit);
        }
    }
    __syncthreads();
    // columns with >= 2 lit pixels: per 32-column word, (ones, twos) carry-save over rows
    {
        const int w = tid & 7, chunk = tid >> 3;           // 32 chunks of 8 rows (last: 3 rows)
        uint32_t ones = 0, twos = 0;
        const int r0 = chunk * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int r = r0 + k;
            const uint32_t v = r < IMG ? bm[r * BMW + w] : 0u;
            twos |= ones & v;
            ones |= v;
        }
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
            const uint32_t o2 = __shfl_xor_sync(0xffffffffu, ones, off);
            const uint32_t t2 = __shfl_xor_sync(0xffffffffu, twos, off);
            twos |= t2 | (ones & o2);
            ones |= o2;
        }
        if ((tid & 31) < 8) {
            red[((tid >> 5) * 8 + w) * 2 + 0] = ones;
            red[((tid >> 5) * 8 + w) * 2 + 1] = twos;
        }
    }
    __syncthreads();
    if (tid < 8) {
        uint32_t ones = 0, twos = 0;
#pragma unroll
        for (int q = 0; q < ENC_THREADS / 32; ++q) {
            const uint32_t o2 = red[(q * 8 + tid) * 2 + 0], t2 = red[(q * 8 + tid) * 2 + 1];
            twos |= t2 | (ones & o2);
            ones |= o2;
        }
        colmask[tid] = twos;
    }
    __syncthreads();
    for (int i = tid; i < IMG * BMW; i += ENC_THREADS)
        bm[PLANE + i] = bm[i] & colmask[i & 7];             // plot_segment.py:59-65
    __syncthreads();
}

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

template <typename T>
__device__ void write_nhwc(const uint32_t* bm, T* __restrict__ out_all, long long img) {
    constexpr int EPV = 16 / (int)sizeof(T);
    const int tid = threadIdx.x;
    const long long e_begin = img * (long long)NEL;
    const long long e_end = e_begin + NEL;
    const long long v_first = (e_begin + EPV - 1) / EPV;
    const long long v_last = e_end / EPV;                       // exclusive
    const int head = (int)(v_first * EPV - e_begin);
    const int tail = (int)(e_end - v_last * EPV);
    if (tid < head) out_all[e_begin + tid] = scalar_value<T>(bm, tid);
    if (tid >= 32 && tid < 32 + tail) {
        const int e = NEL - tail + (tid - 32);
        out_all[e_begin + e] = scalar_value<T>(bm, e);
    }
    uint4* __restrict__ outv = reinterpret_cast<uint4*>(out_all);
    const int nvec = (int)(v_last - v_first);
    for (int i = tid; i < nvec; i += ENC_THREADS) {
        const int e = head + i * EPV;
        const int p0 = e / 3, ph = e - 3 * p0;
        const int r = p0 / IMG, c = p0 - r * IMG;
        const uint32_t w0 = window4(bm, r, c);
        const uint32_t w1 = window4(bm + PLANE, r, c);
        const uint32_t w2 = window4(bm + 2 * PLANE, r, c);
        uint32_t L = 0;
        if (w0 | w2) {                                          // ch1 is a subset of ch0
#pragma unroll
            for (int px = 0; px < 4; ++px)
                L |= (((w0 >> px) & 1u) | (((w1 >> px) & 1u) << 1) | (((w2 >> px) & 1u) << 2))
                     << (3 * px);
        }
        outv[v_first + i] = nhwc_vector<T>(ph, L);
    }
}

// conv1 operand layout: [57*57 s2d pixels][64 ch] fp16, ch = (dy*4+dx)*3 + c, 48..63 zero.
__device__ void write_s2d(const uint32_t* bm, __half* __restrict__ out_all, long long img) {
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

template <int MODE>   // 0: NHWC f32, 1: NHWC f16, 2: conv1 operand (s2d f16)
__global__ void __launch_bounds__(ENC_THREADS)
encode_kernel(const int32_t* __restrict__ rows, long long n, void* __restrict__ out) {
    __shared__ __align__(16) uint32_t bm[3 * PLANE];
    __shared__ LineParams lines[2];
    __shared__ uint32_t red[(ENC_THREADS / 32) * 8 * 2];
    __shared__ uint32_t colmask[8];
    for (long long img = blockIdx.x; img < n; img += gridDim.x) {
        build_bitmap(rows + img * 12, bm, lines, red, colmask);
        if constexpr (MODE == 0) write_nhwc<float>(bm, reinterpret_cast<float*>(out), img);
        if constexpr (MODE == 1) write_nhwc<__half>(bm, reinterpret_cast<__half*>(out), img);
        if constexpr (MODE == 2) write_s2d(bm, reinterpret_cast<__half*>(out), img);
        __syncthreads();
    }
}

}  // namespace

int launch_encode(const int32_t* rows_dev, long long n, void* out, int mode, int num_sms,
                  cudaStream_t stream) {
    if (n <= 0) return 0;
    // 22 KB smem + 256 threads per CTA -> 8 CTAs/SM resident; size the grid as a multiple of the
    // SM count so the grid-stride loop has no ragged tail.
    long long blocks = (long long)num_sms * 8;
    if (blocks > n) blocks = n;
    switch (mode) {
        case 0: encode_kernel<0><<<(unsigned)blocks, ENC_THREADS, 0, stream>>>(rows_dev, n, out); break;
        case 1: encode_kernel<1><<<(unsigned)blocks, ENC_THREADS, 0, stream>>>(rows_dev, n, out); break;
        case 2: encode_kernel<2><<<(unsigned)blocks, ENC_THREADS, 0, stream>>>(rows_dev, n, out); break;
        default: return fail(-1, "launch_encode: bad mode");
    }
    SVX_LAUNCH_CHECK("encode_kernel");
    return 0;
}

}  // namespace svx
