// CUDA-core kernels around the tensor-core layers: max-pool (+LRN), fc8 + softmax + argmax,
// and the layout/precision converters.
//
// Replaces (reference): src/network/alexnet.py:158-166 (`max_pool` 3x3/2 VALID, `lrn` radius 2,
// alpha 2e-5, beta 0.75, bias 1), :58 fc8 (`xw_plus_b`, no ReLU), and
// src/network/predict.py:209 (`tf.argmax(score, 1)`, `tf.nn.softmax(score)`).
#include "common.cuh"
#include "kernels.h"

#include <math_constants.h>

namespace svx {

namespace {

__device__ __forceinline__ void split_store(float v, __half* hi, __half* lo, long long idx) {
    const __half h = __float2half_rn(v);
    hi[idx] = h;
    lo[idx] = __float2half_rn(v - __half2float(h));
}

// One warp per output position; lane l owns CPL consecutive channels (16-byte loads), so the
// LRN window (c-2..c+2) needs only the two edge values of each neighbouring lane (shuffles).
template <int CPL>
__global__ void __launch_bounds__(256) pool_kernel(const PoolParams p, long long total_pos) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const long long warp0 = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * warps_per_block;
    const int per_img = p.out_h * p.out_w;
    const int c0 = lane * CPL;
    const bool active = c0 < p.C;
    for (long long pos = warp0; pos < total_pos; pos += nwarps) {
        const long long img = pos / per_img;
        const int rem = (int)(pos - img * per_img);
        const int y = rem / p.out_w, x = rem - y * p.out_w;
        float m[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) m[j] = active ? -CUDART_INF_F : 0.f;
        if (active) {
            const float* base =
                p.in + ((img * p.in_pos_per_img + (long long)(2 * y) * p.in_grid_w + 2 * x) * p.C + c0);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float4* q = reinterpret_cast<const float4*>(
                        base + (long long)(i * p.in_grid_w + j) * p.C);
#pragma unroll
                    for (int v = 0; v < CPL / 4; ++v) {
                        const float4 t = __ldg(q + v);
                        m[4 * v + 0] = fmaxf(m[4 * v + 0], t.x);
                        m[4 * v + 1] = fmaxf(m[4 * v + 1], t.y);
                        m[4 * v + 2] = fmaxf(m[4 * v + 2], t.z);
                        m[4 * v + 3] = fmaxf(m[4 * v + 3], t.w);
                    }
                }
        }
        float out[CPL];
        if (p.lrn) {
            float sq[CPL + 4];
#pragma unroll
            for (int j = 0; j < CPL; ++j) sq[j + 2] = m[j] * m[j];
            // neighbours' edge squares (inactive lanes hold zeros = the zero padding of the window)
            const float l0 = __shfl_up_sync(0xffffffffu, sq[CPL], 1);       // lane-1's channel CPL-2
            const float l1 = __shfl_up_sync(0xffffffffu, sq[CPL + 1], 1);   // lane-1's channel CPL-1
            const float r0 = __shfl_down_sync(0xffffffffu, sq[2], 1);       // lane+1's channel 0
            const float r1 = __shfl_down_sync(0xffffffffu, sq[3], 1);       // lane+1's channel 1
            sq[0] = lane > 0 ? l0 : 0.f;
            sq[1] = lane > 0 ? l1 : 0.f;
            sq[CPL + 2] = lane < 31 ? r0 : 0.f;
            sq[CPL + 3] = lane < 31 ? r1 : 0.f;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const float s5 = sq[j] + sq[j + 1] + sq[j + 2] + sq[j + 3] + sq[j + 4];
                out[j] = m[j] * pow_m075(1.0f + 2e-5f * s5);
            }
        } else {
#pragma unroll
            for (int j = 0; j < CPL; ++j) out[j] = m[j];
        }
        if (active) {
            const long long o = (long long)(c0 / p.group_real) * p.group_elems +
                                (img * p.out_pos_per_img + (long long)y * p.out_grid_w + x) * p.out_ld + (c0 % p.group_real);
            uint32_t ph[CPL / 2], pl[CPL / 2];
#pragma unroll
            for (int j = 0; j < CPL / 2; ++j) {
                const __half h0 = __float2half_rn(out[2 * j]), h1 = __float2half_rn(out[2 * j + 1]);
                const __half e0 = __float2half_rn(out[2 * j] - __half2float(h0));
                const __half e1 = __float2half_rn(out[2 * j + 1] - __half2float(h1));
                ph[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                pl[j] = (uint32_t)__half_as_ushort(e0) | ((uint32_t)__half_as_ushort(e1) << 16);
            }
            if constexpr (CPL == 4) {
                *reinterpret_cast<uint2*>(p.out_hi + o) = make_uint2(ph[0], ph[1]);
                *reinterpret_cast<uint2*>(p.out_lo + o) = make_uint2(pl[0], pl[1]);
            } else {
                *reinterpret_cast<uint4*>(p.out_hi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                *reinterpret_cast<uint4*>(p.out_lo + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
            }
        }
    }
}

// Second half of the pool fused into the conv2 / conv5 epilogues (layer_tc.cu, STG = 0): one warp per
// pooled position, lane = 8 consecutive channels of 256.  Reads the window maximum (two parts where the
// window straddles a 128-row chunk of the conv's rows), applies the LRN (conv2) and writes the fp16
// hi/lo planes of the next layer's operand.  HBM bound: 1-2 KB read + 1 KB written per position.
// Lane l owns channels [4l, 4l+4) and [128+4l, 128+4l+4): every load / store instruction of the warp
// covers one contiguous 512-byte (fp32) or 256-byte (fp16) run.
template <bool LRN>
__global__ void __launch_bounds__(256) finish_pooled_kernel(const FinishParams p, long long total_pos) {
    constexpr int U = 2;                                  // positions per warp and pass
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const long long warp0 = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * warps_per_block;
    const int per_img = p.pool_h * p.pool_w;
    for (long long pos0 = warp0 * U; pos0 < total_pos; pos0 += nwarps * U) {
        float4 va[U], vb[U], wa[U], wb[U];
        long long orow[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long pos = pos0 + u;
            wa[u] = wb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pos < total_pos) {
                const long long img = pos / per_img;
                const int rem = (int)(pos - img * per_img);
                const int y = rem / p.pool_w, x = rem - y * p.pool_w;
                orow[u] = img * p.out_pos_per_img + (long long)y * p.out_grid_w + x;
                const float* src = p.pooled + pos * 256 + lane * 4;
                va[u] = __ldcs(reinterpret_cast<const float4*>(src));
                vb[u] = __ldcs(reinterpret_cast<const float4*>(src + 128));
                if (pool_crosses(img, y, x, p.in_pos_per_img, p.in_grid_w)) {
                    const float* src2 = p.pooled2 + pos * 256 + lane * 4;
                    wa[u] = __ldcs(reinterpret_cast<const float4*>(src2));
                    wb[u] = __ldcs(reinterpret_cast<const float4*>(src2 + 128));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (pos0 + u >= total_pos) break;             // warp-uniform
            // values are >= 0 (post-ReLU), and a missing second part reads as 0
            const float ma[4] = {fmaxf(va[u].x, wa[u].x), fmaxf(va[u].y, wa[u].y), fmaxf(va[u].z, wa[u].z),
                                 fmaxf(va[u].w, wa[u].w)};
            const float mb[4] = {fmaxf(vb[u].x, wb[u].x), fmaxf(vb[u].y, wb[u].y), fmaxf(vb[u].z, wb[u].z),
                                 fmaxf(vb[u].w, wb[u].w)};
            float oa[4], ob[4];
            if (LRN) {
                // squares of channels c-2 .. c+2: [0,1] from the lane below, [6,7] from the lane above;
                // channels 126..129 cross from the first block of lane 31 to the second block of lane 0
                float qa[8], qb[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) { qa[j + 2] = ma[j] * ma[j]; qb[j + 2] = mb[j] * mb[j]; }
                const float ua0 = __shfl_up_sync(0xffffffffu, qa[4], 1), ua1 = __shfl_up_sync(0xffffffffu, qa[5], 1);
                const float ub0 = __shfl_up_sync(0xffffffffu, qb[4], 1), ub1 = __shfl_up_sync(0xffffffffu, qb[5], 1);
                const float da0 = __shfl_down_sync(0xffffffffu, qa[2], 1), da1 = __shfl_down_sync(0xffffffffu, qa[3], 1);
                const float db0 = __shfl_down_sync(0xffffffffu, qb[2], 1), db1 = __shfl_down_sync(0xffffffffu, qb[3], 1);
                const float lo0 = __shfl_sync(0xffffffffu, qa[4], 31), lo1 = __shfl_sync(0xffffffffu, qa[5], 31);
                const float hi0 = __shfl_sync(0xffffffffu, qb[2], 0), hi1 = __shfl_sync(0xffffffffu, qb[3], 0);
                qa[0] = lane > 0 ? ua0 : 0.f;   qa[1] = lane > 0 ? ua1 : 0.f;
                qa[6] = lane < 31 ? da0 : hi0;  qa[7] = lane < 31 ? da1 : hi1;
                qb[0] = lane > 0 ? ub0 : lo0;   qb[1] = lane > 0 ? ub1 : lo1;
                qb[6] = lane < 31 ? db0 : 0.f;  qb[7] = lane < 31 ? db1 : 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float sa = qa[j] + qa[j + 1] + qa[j + 2] + qa[j + 3] + qa[j + 4];
                    const float sb = qb[j] + qb[j + 1] + qb[j + 2] + qb[j + 3] + qb[j + 4];
                    oa[j] = ma[j] * pow_m075(1.0f + 2e-5f * sa);
                    ob[j] = mb[j] * pow_m075(1.0f + 2e-5f * sb);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { oa[j] = ma[j]; ob[j] = mb[j]; }
            }
            auto split4 = [](const float (&v)[4], uint2& hi, uint2& lo) {
                uint32_t ph[2], pl[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const __half h0 = __float2half_rn(v[2 * j]), h1 = __float2half_rn(v[2 * j + 1]);
                    const __half e0 = __float2half_rn(v[2 * j] - __half2float(h0));
                    const __half e1 = __float2half_rn(v[2 * j + 1] - __half2float(h1));
                    ph[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                    pl[j] = (uint32_t)__half_as_ushort(e0) | ((uint32_t)__half_as_ushort(e1) << 16);
                }
                hi = make_uint2(ph[0], ph[1]);
                lo = make_uint2(pl[0], pl[1]);
            };
            uint2 ha, la, hb, lb;
            split4(oa, ha, la);
            split4(ob, hb, lb);
            const long long o = orow[u] * p.out_ld + lane * 4;
            *reinterpret_cast<uint2*>(p.out_hi + o) = ha;
            *reinterpret_cast<uint2*>(p.out_hi + o + 128) = hb;
            *reinterpret_cast<uint2*>(p.out_lo + o) = la;
            *reinterpret_cast<uint2*>(p.out_lo + o + 128) = lb;
        }
    }
}

// One warp per site, 8 sites per CTA.  Besides labels / probs / logits the kernel can emit the
// 8-byte (label, score) call of every site -- what src/network/predict.py:230,251 consumes -- into
// up to CALL_MAX_SINKS destinations: the local buffer and, in a multi-GPU exchange, the gathered
// buffer of EVERY rank (peer-mapped memory written over NVLink, 64 contiguous bytes per CTA and
// sink).  With `done` set, the last CTA to finish publishes `epoch` to each rank's flag word
// (release, system scope): fc8 + softmax + argmax + all-gather + signal are this one kernel.
__global__ void __launch_bounds__(256)
fc8_softmax_kernel(const __half* __restrict__ x_hi, const __half* __restrict__ x_lo,
                   const float* __restrict__ w8, const float* __restrict__ b8, long long n,
                   int32_t* __restrict__ labels, float* __restrict__ probs,
                   float* __restrict__ logits, const CallSinks sinks) {
    __shared__ int2 call_s[8];
    __shared__ int is_last;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const long long site0 = (long long)blockIdx.x * 8;
    const long long site = site0 + warp;
    if (site < n) {
        const __half* xh = x_hi + site * 4096;
        const __half* xl = x_lo + site * 4096;
        float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        // 8 consecutive features per lane and pass: one 16-byte load per plane, the 40 weights of those
        // features as ten 16-byte loads (w8 is [4096][5], so 8 rows are 160 contiguous bytes)
        for (int i = 0; i < 16; ++i) {
            const int k0 = 8 * lane + 256 * i;
            const uint4 h = *reinterpret_cast<const uint4*>(xh + k0);
            const uint4 l = *reinterpret_cast<const uint4*>(xl + k0);
            const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
            float x[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                x[2 * q] = __half2float(__ushort_as_half((unsigned short)(hw[q] & 0xFFFFu))) +
                           __half2float(__ushort_as_half((unsigned short)(lw[q] & 0xFFFFu)));
                x[2 * q + 1] = __half2float(__ushort_as_half((unsigned short)(hw[q] >> 16))) +
                               __half2float(__ushort_as_half((unsigned short)(lw[q] >> 16)));
            }
            float w[40];
            const float4* wp = reinterpret_cast<const float4*>(w8 + k0 * 5);
#pragma unroll
            for (int q = 0; q < 10; ++q) {
                const float4 t = __ldg(wp + q);
                w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e)
#pragma unroll
                for (int j = 0; j < 5; ++j) acc[j] = fmaf(x[e], w[5 * e + j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
        if (lane == 0) {
            float l[5], mx = -CUDART_INF_F;
            int arg = 0;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                l[j] = acc[j] + b8[j];
                if (l[j] > mx) { mx = l[j]; arg = j; }       // first maximum, like tf.argmax
            }
            float e[5], sum = 0.f;
#pragma unroll
            for (int j = 0; j < 5; ++j) { e[j] = expf(l[j] - mx); sum += e[j]; }
            if (labels) labels[site] = arg;
            float score = 0.f;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const float pj = e[j] / sum;
                if (probs) probs[site * 5 + j] = pj;
                if (logits) logits[site * 5 + j] = l[j];
                if (j == arg) score = pj;
            }
            call_s[warp] = make_int2(arg, __float_as_int(score));
        }
    }
    if (sinks.count == 0) return;                          // uniform over the grid
    __syncthreads();
    {   // thread t -> sink t / 8, site t % 8: every sink receives one 64-byte run per CTA
        const int r = threadIdx.x >> 3, i = threadIdx.x & 7;
        if (r < sinks.count && site0 + i < n) sinks.ptr[r][site0 + i] = call_s[i];
    }
    if (sinks.done == nullptr) return;
    __threadfence_system();                                // this thread's peer stores are performed
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(sinks.done, 1u);
        is_last = prev + 1 == gridDim.x;
        if (is_last) *sinks.done = 0;                      // ready for the next launch
    }
    __syncthreads();
    if (is_last) {
        __threadfence_system();
        if (threadIdx.x < sinks.count)
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(sinks.flag[threadIdx.x]), "l"(sinks.epoch)
                         : "memory");
    }
}

// One thread per rank waits until that rank's calls have landed in the local gathered buffer
// (its fc8 kernel has published `epoch`).  Bounded: after `timeout_ns` the rank is reported in
// *error (1 + rank), its region is poisoned and the kernel exits instead of hanging the device.
__global__ void exchange_wait_kernel(const unsigned long long* __restrict__ my_flags, int world,
                                     unsigned long long epoch, unsigned long long timeout_ns,
                                     unsigned int* __restrict__ error, int2* __restrict__ gathered,
                                     long long per_rank) {
    const int r = threadIdx.x;
    bool timed_out = false;
    if (r < world) {
        unsigned long long t0, now, v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(my_flags + r) : "memory");
            if (v >= epoch) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (now - t0 > timeout_ns) { timed_out = true; break; }
            __nanosleep(200);
        }
    }
    unsigned int late = __ballot_sync(0xffffffffu, timed_out);
    if (late == 0) return;
    if (threadIdx.x == 0) {
        *error = 1u + (unsigned)(__ffs(late) - 1);
        __threadfence_system();
    }
    const int2 poison = make_int2(-1, 0x7fc00000);           // label -1, score NaN
    while (late) {
        const int rr = __ffs(late) - 1;
        late &= late - 1;
        for (long long i = threadIdx.x; i < per_rank; i += 32) gathered[(long long)rr * per_rank + i] = poison;
    }
}

template <typename T>
__global__ void nhwc_to_s2d_kernel(const T* __restrict__ img, long long n, __half* __restrict__ out) {
    constexpr int S2D = 57, IMG = 227;
    const long long total = n * S2D * S2D * 8;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total;
         v += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(v & 7);
        const long long sp_all = v >> 3;
        const long long im = sp_all / (S2D * S2D);
        const int sp = (int)(sp_all - im * (S2D * S2D));
        const int Y = sp / S2D, X = sp - Y * S2D;
        uint32_t h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = 8 * q + j;
            float val = 0.f;
            if (k < 48) {
                const int d = k / 3, ch = k - 3 * d;
                const int r = 4 * Y + (d >> 2), c = 4 * X + (d & 3);
                if (r < IMG && c < IMG)
                    val = (float)img[(im * IMG * IMG + (long long)r * IMG + c) * 3 + ch];
            }
            h[j] = __half_as_ushort(__float2half_rn(val));
        }
        reinterpret_cast<uint4*>(out)[v] =
            make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
    }
}

__global__ void split_hilo_kernel(const float* __restrict__ in, long long count,
                                  __half* __restrict__ hi, __half* __restrict__ lo) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (long long)gridDim.x * blockDim.x)
        split_store(in[i], hi, lo, i);
}

}  // namespace

int launch_pool(const PoolParams& p, long long n_img, int num_sms, cudaStream_t stream) {
    const long long total = n_img * p.out_h * p.out_w;
    if (total <= 0) return 0;
    // the one user left is the dense path's pool1 (96 channels: 24 lanes x 4); pool2 / pool5 run in the
    // conv epilogues (layer_tc.cu) and are finished by finish_pooled_kernel
    constexpr int cpl = 4;
    if (p.C % cpl != 0 || p.C / cpl > 32 || p.group_real % cpl != 0 || p.group_elems % cpl != 0 ||
        p.out_ld % cpl != 0)
        return fail(-1, "pool: unsupported channel count / grouping");
    long long blocks = (long long)num_sms * 8;           // 8 CTAs x 8 warps resident per SM
    if (blocks > (total + 7) / 8) blocks = (total + 7) / 8;
    pool_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(p, total);
    SVX_LAUNCH_CHECK("pool_kernel");
    return 0;
}

int launch_finish_pooled(const FinishParams& p, long long n_img, int num_sms, cudaStream_t stream) {
    const long long total = n_img * p.pool_h * p.pool_w;
    if (total <= 0) return 0;
    if (p.out_ld % 8 != 0) return fail(-1, "finish_pooled: output rows must be 16-byte aligned");
    long long blocks = (long long)num_sms * 8;
    if (blocks > (total + 15) / 16) blocks = (total + 15) / 16;
    if (p.lrn) finish_pooled_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(p, total);
    else       finish_pooled_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(p, total);
    SVX_LAUNCH_CHECK("finish_pooled_kernel");
    return 0;
}

int launch_fc8_softmax(const __half* x_hi, const __half* x_lo, const float* w8, const float* b8,
                       long long n, int32_t* labels, float* probs, float* logits,
                       const CallSinks& sinks, cudaStream_t stream) {
    if (n <= 0) return 0;
    if (sinks.count < 0 || sinks.count > CALL_MAX_SINKS) return fail(-1, "fc8: bad sink count");
    fc8_softmax_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(x_hi, x_lo, w8, b8, n, labels,
                                                                     probs, logits, sinks);
    SVX_LAUNCH_CHECK("fc8_softmax_kernel");
    return 0;
}

int launch_exchange_wait(const unsigned long long* my_flags, int world, unsigned long long epoch,
                         unsigned long long timeout_ns, unsigned int* error, int2* gathered,
                         long long per_rank, cudaStream_t stream) {
    if (world < 1 || world > CALL_MAX_SINKS) return fail(-1, "exchange: bad world size");
    exchange_wait_kernel<<<1, 32, 0, stream>>>(my_flags, world, epoch, timeout_ns, error, gathered, per_rank);
    SVX_LAUNCH_CHECK("exchange_wait_kernel");
    return 0;
}

int launch_nhwc_to_s2d(const void* images, int dtype, long long n, __half* out, cudaStream_t stream) {
    if (n <= 0) return 0;
    const long long total = n * 57 * 57 * 8;
    const unsigned blocks = (unsigned)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    if (dtype == 0)
        nhwc_to_s2d_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(images), n, out);
    else if (dtype == 1)
        nhwc_to_s2d_kernel<__half><<<blocks, 256, 0, stream>>>(static_cast<const __half*>(images), n, out);
    else
        return fail(-1, "nhwc_to_s2d: bad dtype");
    SVX_LAUNCH_CHECK("nhwc_to_s2d_kernel");
    return 0;
}

int launch_split_hilo(const float* in, long long count, __half* hi, __half* lo, cudaStream_t stream) {
    if (count <= 0) return 0;
    const unsigned blocks = (unsigned)((count + 255) / 256 > 148 * 16 ? 148 * 16 : (count + 255) / 256);
    split_hilo_kernel<<<blocks, 256, 0, stream>>>(in, count, hi, lo);
    SVX_LAUNCH_CHECK("split_hilo_kernel");
    return 0;
}

}  // namespace svx
