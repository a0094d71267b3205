// Fused front end of the classify path: packed site row -> lit-pixel bitmap -> conv1 + ReLU ->
// 3x3/2 max-pool -> LRN -> conv2 operand (fp16 hi/lo, padded layout), in ONE kernel; the 227x227x3
// image, conv1's input tensor and conv1's 55x55x96 output never touch HBM.
//
// Replaces, for images produced by this library's own encoder:
//   src/network/create_batch.py:103-152 + src/segmentplot/plot_segment.py:9-73   (image)
//   src/network/alexnet.py:29-31: conv1 (11x11/4 VALID, 96) + ReLU, pool1, norm1
//
// Why this is exact and cheap.  After mean subtraction every pixel of channel c is lo_c or
// lo_c + 255, so for every conv1 output
//     conv1[Y,X,n] = base[n] + 255 * sum_{lit (r,c,ch) in the 11x11 window} W[r-4Y, c-4X, ch, n],
//     base[n]      = bias[n] + sum_{kh,kw,ch} lo_ch * W[kh,kw,ch,n]          (host, in double)
// -- the same dense convolution with its terms regrouped; only fp32 rounding order differs.  An
// image has <= ~1100 lit channel-pixels of 154 587, so > 80 % of the pooled 27x27 positions see
// pure background (one precomputed 96-vector) and the rest need a few dozen FMAs per channel
// instead of 3 267 MACs.  The dense tcgen05 conv1 (layer_tc.cu) remains the path for arbitrary
// images (svx_forward) and the parity tests cross-check the two.
//
// One CTA per site: bitmap in shared memory (encoder_bitmap.cuh, bit-exact with the reference
// rasteriser); the lit pixels mark the conv1 positions whose 11x11 window they fall in (~350 of
// 3 025) and those mark the pooled positions that see them (~140 of 729).  Background pooled
// positions are streamed out with 16-byte stores.  Phase A: a quarter-warp per dirty conv1 position
// walks the lit pixels of its window once, lane = 12 consecutive output channels (16-byte weight
// loads), result to a per-CTA scratch that stays in L2.  Phase B: one warp per dirty pooled position
// (lanes 0..23 = 4 channels each) takes the max over its 3x3 conv positions (scratch value, or the
// background value), ReLU, LRN by shuffle, fp16 hi/lo split.  ncu (round 2) had phase B at ~45 % of
// the kernel's instructions with 3-channel lanes: every lane repeated the nine slot lookups and
// issued 27 four-byte loads that used 14 of every 32 bytes fetched; now lanes 0..8 look the slots up
// and broadcast them, and a position's nine conv vectors are nine 16-byte loads per lane.
#include "common.cuh"
#include "encoder_bitmap.cuh"
#include "kernels.h"

#include <math_constants.h>

namespace svx {

namespace {

using namespace bitmap;

constexpr int POOLED = 27;
constexpr int NPOS = POOLED * POOLED;        // 729
constexpr int G2W = 29, G2POS = G2W * G2W;   // conv2 operand grid (27 + 2 shared pad)

// LRN over channels with lane = 4 consecutive channels, lanes 0..23 (radius 2, alpha 2e-5, beta .75,
// bias 1); lanes 24..31 must hold zeros (they are the zero padding above channel 95)
__device__ __forceinline__ void lrn4(const float (&m)[4], float (&out)[4], int lane) {
    float sq[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) sq[j + 2] = m[j] * m[j];
    const float l0 = __shfl_up_sync(0xffffffffu, sq[4], 1);       // lane-1's channels 2, 3
    const float l1 = __shfl_up_sync(0xffffffffu, sq[5], 1);
    const float r0 = __shfl_down_sync(0xffffffffu, sq[2], 1);     // lane+1's channels 0, 1
    const float r1 = __shfl_down_sync(0xffffffffu, sq[3], 1);
    sq[0] = lane > 0 ? l0 : 0.f;
    sq[1] = lane > 0 ? l1 : 0.f;
    sq[6] = lane < 31 ? r0 : 0.f;
    sq[7] = lane < 31 ? r1 : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float s5 = sq[j] + sq[j + 1] + sq[j + 2] + sq[j + 3] + sq[j + 4];
        out[j] = m[j] * pow_m075(1.0f + 2e-5f * s5);
    }
}

constexpr int CONV_W = 55, NCONV = CONV_W * CONV_W;      // conv1 output grid, 3025 positions
constexpr unsigned short CLEAN = 0xFFFF;
// Scratch slots per CTA.  A segment is a digital line of <= 227 pixels, i.e. <= 57 steps on the conv1
// grid, and every pixel marks at most a 3x3 block of positions, so one line dirties < 300 positions and
// a site < 600.  Keeping the per-CTA regions small keeps the ~600 concurrent regions inside the TLB
// reach; positions beyond the capacity (not reachable with two segments) are recomputed in phase B.
constexpr int SCRATCH_SLOTS = 640;

// bits [c, c+11) of bitmap row r (c + 10 <= 226)
__device__ __forceinline__ uint32_t window11(const uint32_t* plane, int r, int c) {
    const uint32_t* rowp = plane + r * BMW;
    const int wi = c >> 5;
    return __funnelshift_r(rowp[wi], rowp[wi + 1], c & 31) & 0x7FFu;
}

// conv1 value (before ReLU) of position (Y, X) for NCH consecutive channels starting at c0
template <int NCH>
__device__ __forceinline__ void conv1_at(const uint32_t* bm, const FrontParams& P, int Y, int X, int c0,
                                         float (&acc)[NCH]) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) acc[j] = __ldg(P.base + c0 + j);
    for (int kh = 0; kh < 11; ++kh) {
        uint32_t w0 = window11(bm, 4 * Y + kh, 4 * X);
        if (!w0) continue;
        const uint32_t w1 = window11(bm + PLANE, 4 * Y + kh, 4 * X);
        const uint32_t w2 = window11(bm + 2 * PLANE, 4 * Y + kh, 4 * X);
        while (w0) {
            const int kw = __ffs(w0) - 1;
            w0 &= w0 - 1;
            const float* wp = P.w255 + ((kh * 11 + kw) * 3) * 96 + c0;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                if (ch == 1 && !((w1 >> kw) & 1u)) continue;
                if (ch == 2 && !((w2 >> kw) & 1u)) continue;
#pragma unroll
                for (int j = 0; j < NCH; ++j) acc[j] += __ldg(wp + ch * 96 + j);
            }
        }
    }
}

#ifndef SVX_FRONT_CTAS
#define SVX_FRONT_CTAS 4          // resident CTAs per SM the register budget is set for
#endif
#ifndef SVX_FRONT_CARVE
#define SVX_FRONT_CARVE 72        // shared-memory carve-out in percent (the rest is L1 for the conv1 weights)
#endif

__global__ void __launch_bounds__(FRONT_THREADS, SVX_FRONT_CTAS)
front_kernel(const int32_t* __restrict__ rows, long long n, const FrontParams P) {
    __shared__ __align__(16) uint32_t bm[3 * PLANE];
    __shared__ LineParams lines[2];
    __shared__ uint32_t red[(FRONT_THREADS / 32) * 8 * 2];
    __shared__ uint32_t colmask[8];
    __shared__ __align__(16) unsigned short bg[2][96];     // background vector: hi plane, lo plane
    __shared__ uint32_t cdirty[CONV_W * 2];                // dirty conv1 positions: 55 rows x 64 bits
    __shared__ unsigned short cslot[NCONV];                // conv position -> scratch slot (CLEAN if none)
    __shared__ unsigned short clist[NCONV];                // scratch slot -> conv position
    __shared__ uint32_t dirty_mask[(NPOS + 31) / 32];
    __shared__ unsigned short dirty_list[NPOS];
    __shared__ int dirty_count, conv_count;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // phase B and the background vector: lanes 0..23 own 4 consecutive channels (16-byte loads)
    const int c4 = 4 * lane;
    const bool chan = lane < 24;
    float* __restrict__ scratch = P.scratch + (size_t)blockIdx.x * SCRATCH_SLOTS * 96;

    float4 base4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (chan) base4 = __ldg(reinterpret_cast<const float4*>(P.base + c4));
    // background value of every pooled position: LRN(ReLU(base)) (max-pool of a constant)
    if (warp == 0) {
        const float m[4] = {fmaxf(base4.x, 0.f), fmaxf(base4.y, 0.f), fmaxf(base4.z, 0.f), fmaxf(base4.w, 0.f)};
        float o[4];
        lrn4(m, o, lane);
        if (chan) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __half h = __float2half_rn(o[j]);
                bg[0][c4 + j] = __half_as_ushort(h);
                bg[1][c4 + j] = __half_as_ushort(__float2half_rn(o[j] - __half2float(h)));
            }
        }
    }

    // prefetch of the next site's row: see encode_kernel
    __shared__ int32_t rowbuf[2][12];
    if (tid < 12 && blockIdx.x < n) rowbuf[0][tid] = rows[(long long)blockIdx.x * 12 + tid];
    __syncthreads();
    int cur = 0;
    for (long long img = blockIdx.x; img < n; img += gridDim.x) {
        const long long nxt = img + gridDim.x;
        int32_t pre = 0;
        if (tid < 12 && nxt < n) pre = __ldg(rows + nxt * 12 + tid);
        if (tid < (NPOS + 31) / 32) dirty_mask[tid] = 0;
        if (tid < CONV_W * 2) cdirty[tid] = 0;
        if (tid == 0) { dirty_count = 0; conv_count = 0; }
        build_bitmap<FRONT_THREADS>(rowbuf[cur], bm, lines, red, colmask);   // ends with __syncthreads()

        // ---- conv1 positions whose 11x11 window holds a lit pixel: pixel (r,c) touches
        //      Y in [ceil((r-10)/4), r/4] x X in [ceil((c-10)/4), c/4]
        for (int w = tid; w < IMG * BMW; w += FRONT_THREADS) {
            uint32_t bits = bm[w];
            const int r = w >> 3, cbase = (w & 7) << 5;
            const int y_lo = max((r - 7) >> 2, 0), y_hi = min(r >> 2, CONV_W - 1);
            while (bits) {
                const int c = cbase + __ffs(bits) - 1;
                bits &= bits - 1;
                const int x_lo = max((c - 7) >> 2, 0), x_hi = min(c >> 2, CONV_W - 1);
                // x_hi - x_lo <= 2: build the (up to 3-bit) column mask once
                const unsigned long long m = ((2ull << x_hi) - (1ull << x_lo));
                for (int y = y_lo; y <= y_hi; ++y) {
                    if ((uint32_t)m) atomicOr(&cdirty[2 * y], (uint32_t)m);
                    if ((uint32_t)(m >> 32)) atomicOr(&cdirty[2 * y + 1], (uint32_t)(m >> 32));
                }
            }
        }
        __syncthreads();
        // ---- compact them: slot <-> position
        for (int q = tid; q < NCONV; q += FRONT_THREADS) {
            const int y = q / CONV_W, x = q - y * CONV_W;
            unsigned short s = CLEAN;
            if ((cdirty[2 * y + (x >> 5)] >> (x & 31)) & 1u) {
                s = (unsigned short)atomicAdd(&conv_count, 1);
                clist[s] = (unsigned short)q;
            }
            cslot[q] = s;
        }
        // ---- pooled positions that see a dirty conv position (= a lit pixel in their 19x19 field)
        for (int p = tid; p < NPOS; p += FRONT_THREADS) {
            const int py = p / POOLED, px = p - py * POOLED;
            uint32_t any = 0;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int y = 2 * py + a, x0 = 2 * px;
                const unsigned long long rowbits = ((unsigned long long)cdirty[2 * y + 1] << 32) | cdirty[2 * y];
                any |= (uint32_t)((rowbits >> x0) & 7ull);
            }
            if (any) {
                atomicOr(&dirty_mask[p >> 5], 1u << (p & 31));
                dirty_list[atomicAdd(&dirty_count, 1)] = (unsigned short)((py << 5) | px);
            }
        }
        __syncthreads();

        // ---- background positions, 16-byte stores.  Four (plane, channel group) streams of 64
        //      threads; within a stream 6 threads (octets of 8 channels) cover one position and 10
        //      positions go out per pass, so a thread's value, plane and column never change and
        //      the loop body is a bit test, one multiply-add and the store (the first version
        //      derived plane / position / octet from a flat index: 27 % of the kernel's instructions)
        __half* const planes[2] = {P.x2_hi, P.x2_lo};
        const long long img_row0 = img * G2POS;
        {
            static_assert(FRONT_THREADS == 256, "four streams of 64 threads");
            const int stream = tid >> 6, t64 = tid & 63;
            if (t64 < 60) {
                const int plane = stream >> 1, g = stream & 1;
                const int pl = t64 / 6, q = t64 - 6 * pl;
                const uint4 val = *reinterpret_cast<const uint4*>(&bg[plane][48 * g + 8 * q]);
                __half* const dst = planes[plane] + (long long)g * P.group_elems + img_row0 * P.ld + 8 * q;
                int py = 0, px = pl;
                for (int p = pl; p < NPOS; p += 10) {
                    if (!((dirty_mask[p >> 5] >> (p & 31)) & 1u))
                        *reinterpret_cast<uint4*>(dst + (long long)((py * G2W + px) * P.ld)) = val;
                    px += 10;
                    if (px >= POOLED) { px -= POOLED; ++py; }
                }
            }
        }

        // ---- phase A: every dirty conv1 position once; a quarter-warp (8 lanes x 12 channels,
        //      16-byte weight loads) per position, so one instruction stream serves four positions
        const int nc = min(conv_count, SCRATCH_SLOTS);
        {
            const int sub = lane >> 3, c12 = 12 * (lane & 7);
            for (int i0 = warp * 4; i0 < nc; i0 += (FRONT_THREADS / 32) * 4) {
                const int i = i0 + sub;
                if (i >= nc) continue;
                const int q = clist[i];
                const int Y = q / CONV_W, X = q - Y * CONV_W;
                float4 a0 = __ldg(reinterpret_cast<const float4*>(P.base + c12));
                float4 a1 = __ldg(reinterpret_cast<const float4*>(P.base + c12 + 4));
                float4 a2 = __ldg(reinterpret_cast<const float4*>(P.base + c12 + 8));
                for (int kh = 0; kh < 11; ++kh) {
                    uint32_t w0 = window11(bm, 4 * Y + kh, 4 * X);
                    if (!w0) continue;
                    const uint32_t w1 = window11(bm + PLANE, 4 * Y + kh, 4 * X);
                    const uint32_t w2 = window11(bm + 2 * PLANE, 4 * Y + kh, 4 * X);
                    while (w0) {
                        const int kw = __ffs(w0) - 1;
                        w0 &= w0 - 1;
                        const float4* wp = reinterpret_cast<const float4*>(P.w255 + ((kh * 11 + kw) * 3) * 96 + c12);
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            if (ch == 1 && !((w1 >> kw) & 1u)) continue;
                            if (ch == 2 && !((w2 >> kw) & 1u)) continue;
                            const float4 d0 = __ldg(wp + ch * 24), d1 = __ldg(wp + ch * 24 + 1), d2 = __ldg(wp + ch * 24 + 2);
                            a0.x += d0.x; a0.y += d0.y; a0.z += d0.z; a0.w += d0.w;
                            a1.x += d1.x; a1.y += d1.y; a1.z += d1.z; a1.w += d1.w;
                            a2.x += d2.x; a2.y += d2.y; a2.z += d2.z; a2.w += d2.w;
                        }
                    }
                }
                float4* o = reinterpret_cast<float4*>(scratch + i * 96 + c12);
                o[0] = a0; o[1] = a1; o[2] = a2;
            }
        }
        __syncthreads();                                    // scratch visible to the whole CTA

        // ---- phase B: flagged pooled positions: max over the 3x3 conv positions, LRN, store.  One warp
        //      per position.  The nine scratch slots are looked up by lanes 0..8 and broadcast; if none
        //      overflowed the scratch (never, with two segments) the nine conv vectors are fetched with
        //      independent predicated 16-byte loads: one L2 round trip per pooled position
        const int nd = dirty_count;
        for (int i = warp; i < nd; i += FRONT_THREADS / 32) {
            const int code = dirty_list[i];
            const int py = code >> 5, px = code & 31;
            int my_slot = CLEAN;
            if (lane < 9) my_slot = cslot[(2 * py + lane / 3) * CONV_W + 2 * px + (lane % 3)];
            const bool overflow = __any_sync(0xffffffffu, my_slot != CLEAN && my_slot >= SCRATCH_SLOTS);
            float m[4] = {0.f, 0.f, 0.f, 0.f}, o[4];         // ReLU folded into the max with 0
            if (!overflow) {
                float4 v[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const int sl = __shfl_sync(0xffffffffu, my_slot, k);
                    v[k] = base4;                            // clean conv positions hold the background value
                    if (sl != CLEAN && chan) v[k] = *reinterpret_cast<const float4*>(scratch + sl * 96 + c4);
                }
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    m[0] = fmaxf(m[0], v[k].x); m[1] = fmaxf(m[1], v[k].y);
                    m[2] = fmaxf(m[2], v[k].z); m[3] = fmaxf(m[3], v[k].w);
                }
            } else {
#pragma unroll 1
                for (int k = 0; k < 9; ++k) {
                    const int sl = __shfl_sync(0xffffffffu, my_slot, k);
                    float4 v = base4;
                    if (chan) {
                        if (sl != CLEAN && sl < SCRATCH_SLOTS) {
                            v = *reinterpret_cast<const float4*>(scratch + sl * 96 + c4);
                        } else if (sl != CLEAN) {            // beyond the scratch capacity: recompute here
                            float r[4];
                            conv1_at<4>(bm, P, 2 * py + k / 3, 2 * px + (k % 3), c4, r);
                            v = make_float4(r[0], r[1], r[2], r[3]);
                        }
                    }
                    m[0] = fmaxf(m[0], v.x); m[1] = fmaxf(m[1], v.y);
                    m[2] = fmaxf(m[2], v.z); m[3] = fmaxf(m[3], v.w);
                }
            }
            lrn4(m, o, lane);                                // lanes 24..31 carry zeros
            if (chan) {
                const long long off =
                    (long long)(c4 / 48) * P.group_elems + (img_row0 + py * G2W + px) * P.ld + (c4 % 48);
                uint32_t ph[2], pl[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const __half h0 = __float2half_rn(o[2 * j]), h1 = __float2half_rn(o[2 * j + 1]);
                    const __half e0 = __float2half_rn(o[2 * j] - __half2float(h0));
                    const __half e1 = __float2half_rn(o[2 * j + 1] - __half2float(h1));
                    ph[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                    pl[j] = (uint32_t)__half_as_ushort(e0) | ((uint32_t)__half_as_ushort(e1) << 16);
                }
                *reinterpret_cast<uint2*>(P.x2_hi + off) = make_uint2(ph[0], ph[1]);
                *reinterpret_cast<uint2*>(P.x2_lo + off) = make_uint2(pl[0], pl[1]);
            }
        }
        if (tid < 12) rowbuf[cur ^ 1][tid] = pre;
        cur ^= 1;
        __syncthreads();
    }
}

}  // namespace

int launch_front(const int32_t* rows_dev, long long n, const FrontParams& P, int num_sms,
                 cudaStream_t stream) {
    if (n <= 0) return 0;
    static int blocks_cache[64] = {};             // per device; one wave of resident CTAs (see launch_encode)
    int dev = 0;
    cudaGetDevice(&dev);
    int& blocks_per_sm = blocks_cache[(dev >= 0 && dev < 64) ? dev : 0];
    if (blocks_per_sm == 0) {
        // 4 CTAs x 38 KB of shared memory fit the 164 KB configuration, which leaves ~90 KB of L1 for
        // the conv1 weight vectors phase A walks (the channel-0 vectors alone are 46 KB); with the
        // maximum carve-out L1 was ~25 KB and half of those loads went to L2
        cudaFuncSetAttribute(front_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, SVX_FRONT_CARVE);
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, front_kernel, bitmap::FRONT_THREADS, 0) !=
                cudaSuccess || nb < 1)
            nb = 4;
        blocks_per_sm = nb;
    }
    // per-site work varies with the number of lit pixels: oversubscribe (8 waves) so that the
    // block scheduler balances the load instead of a static stride over one resident wave
    long long blocks = (long long)num_sms * blocks_per_sm * 8;
    if (blocks > n) blocks = n;
    if (blocks > P.scratch_blocks) blocks = P.scratch_blocks;
    front_kernel<<<(unsigned)blocks, bitmap::FRONT_THREADS, 0, stream>>>(rows_dev, n, P);
    SVX_LAUNCH_CHECK("front_kernel");
    return 0;
}

}  // namespace svx
