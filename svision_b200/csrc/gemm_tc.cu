// Tensor-core layer kernel for the SVision CNN (conv1..conv5, fc6, fc7) on sm_100a:
// TMA-fed, tcgen05.mma with TMEM accumulators, warp-specialised, persistent.
//
// Replaces the TF-CPU kernels behind src/network/alexnet.py:100-155 (`conv`, `fc`):
// tf.nn.conv2d (+ the groups split/concat of :124-129), bias_add, relu, xw_plus_b.
//
// Every layer is the same "shifted GEMM"
//     D[m, n] = sum_{tap t} sum_{c}  A[m + row_off[t], c] * W[n, t*Cg + c]
// over activations stored as a 2-D row-major matrix [positions, channels] whose spatial zero
// padding is part of the layout (DESIGN.md "HBM layouts"): a filter tap is then just a row
// offset of the same matrix, so the A operand of every tap is one plain 2-D TMA tile and no
// im2col buffer ever exists.  Rows that fall before/after the matrix are zero-filled by TMA.
//
// Numerics: operands are fp16 hi/lo pairs (x = hi + lo exactly to ~22 bits); per
// 16-wide k-step the kernel issues  A_hi x [B_hi;B_lo]  as ONE N=2*BLOCK_N MMA (main columns
// get hi*hi, cross columns hi*lo) and  A_lo x B_hi  into the main columns ("3-pass").  conv1's
// activations are exact in fp16 so it runs 2-pass; a 1-pass mode exists for comparison.
// The tensor core's fp32 accumulation truncates, so error grows linearly with the length of an
// accumulation chain (measured: 5e-7 relative at K=64, 1.5e-5 at K=4096).  Chains are therefore
// kept short: every `chunk_kblocks` k-blocks the TMEM accumulator is handed to the epilogue
// warps, which add it into an fp32 running sum held in registers (round-to-nearest), while the
// MMA warp continues into the other TMEM buffer.
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2-5 epilogue
// (TMEM chunk -> register running sum; after the last chunk bias/ReLU -> fp32 or fp16 hi/lo ->
// global).  The two TMEM buffers alternate per chunk, also across tile boundaries, so chunk
// draining and the final store overlap the MMAs of the next chunk/tile.
#include "common.cuh"
#include "kernels.h"

#include <mutex>

namespace svx {

namespace {

constexpr int BLOCK_M = GEMM_BLOCK_M;
constexpr int BLOCK_K = GEMM_BLOCK_K;
constexpr int UMMA_K = 16;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;      // 16 KB
constexpr int GEMM_THREADS = 192;
constexpr int SMEM_BUDGET = 220 * 1024;

template <int BLOCK_N> struct Cfg {
    static constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
    static constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
    static constexpr int ACC_STRIDE = 256;      // TMEM columns per buffer: main | cross
    static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
    static_assert(BLOCK_N <= 128, "running sums live in registers: BLOCK_N <= 128");
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack
    static_assert(STAGES >= 2, "need at least a double buffer");
};

template <int BLOCK_N>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ GemmLayer L) {
    using C = Cfg<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[C::STAGES];
    __shared__ uint64_t empty_bar[C::STAGES];
    __shared__ uint64_t tmem_full_bar[2];
    __shared__ uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float bias_s[BLOCK_N];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // 1024-byte aligned operand ring (SWIZZLE_128B atoms are 8 rows x 128 B)
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

    const int num_m_tiles = (int)((L.m_rows + BLOCK_M - 1) / BLOCK_M);
    const int n_tiles = L.n_per_group / BLOCK_N;
    const int tiles_per_group = num_m_tiles * n_tiles;
    const int total_tiles = tiles_per_group * L.groups;
    const int kblocks = L.taps * L.cblocks;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&L.tm_a_hi);
        tma_prefetch_desc(&L.tm_a_lo);
        tma_prefetch_desc(&L.tm_b_hi);
        tma_prefetch_desc(&L.tm_b_lo);
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full_bar[b], 1);
            mbar_init(&tmem_empty_bar[b], 4);      // one arrive per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_smem, C::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const uint32_t stage_tx = A_TILE_BYTES * (1 + (L.use_a_lo ? 1 : 0)) +
                                      C::B_TILE_BYTES * (1 + (L.use_b_lo ? 1 : 0));
            int stage = 0;
            uint32_t phase = 0;
            long long c_prod_wait = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int g = tile / tiles_per_group;
                const int rem = tile - g * tiles_per_group;
                const int n_tile = rem / num_m_tiles;
                const int m_tile = rem - n_tile * num_m_tiles;
                const int m0 = m_tile * BLOCK_M;
                const int n0 = g * L.n_per_group + n_tile * BLOCK_N;
                const int a_col0 = g * L.a_group_cols;
                int kb = 0;
                for (int t = 0; t < L.taps; ++t) {
                    const int a_row = m0 + L.row_off[t];
                    for (int cb = 0; cb < L.cblocks; ++cb, ++kb) {
                        { const long long t0 = clock64(); mbar_wait(&empty_bar[stage], phase ^ 1u); c_prod_wait += clock64() - t0; }
                        uint8_t* st = smem + stage * C::STAGE_BYTES;
                        mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
                        tma_load_2d(&L.tm_a_hi, &full_bar[stage], st, a_col0 + cb * BLOCK_K, a_row);
                        if (L.use_a_lo)
                            tma_load_2d(&L.tm_a_lo, &full_bar[stage], st + A_TILE_BYTES,
                                        a_col0 + cb * BLOCK_K, a_row);
                        tma_load_2d(&L.tm_b_hi, &full_bar[stage], st + 2 * A_TILE_BYTES,
                                    kb * BLOCK_K, n0);
                        if (L.use_b_lo)
                            tma_load_2d(&L.tm_b_lo, &full_bar[stage],
                                        st + 2 * A_TILE_BYTES + C::B_TILE_BYTES, kb * BLOCK_K, n0);
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
            if (L.dbg) atomicAdd(&L.dbg[4], (unsigned long long)c_prod_wait);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_n = umma_idesc_f16(BLOCK_N);
            constexpr uint32_t idesc_2n = umma_idesc_f16(2 * BLOCK_N);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            long long c_wait_op = 0, c_wait_tm = 0, c_kb = 0;
            const long long c_start = clock64();
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                for (int kb0 = 0; kb0 < kblocks; kb0 += L.chunk_kblocks) {
                    const int kb1 = min(kb0 + L.chunk_kblocks, kblocks);
                    { const long long t0 = clock64(); mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u); c_wait_tm += clock64() - t0; }
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(acc * C::ACC_STRIDE);
                    for (int kb = kb0; kb < kb1; ++kb) {
                        { const long long t0 = clock64(); mbar_wait(&full_bar[stage], phase); c_wait_op += clock64() - t0; ++c_kb; }
                        tc_fence_after();
                        const uint32_t st = smem_u32(smem + stage * C::STAGE_BYTES);
                        const uint64_t da_hi = umma_desc_sw128(st);
                        const uint64_t da_lo = umma_desc_sw128(st + A_TILE_BYTES);
                        const uint64_t db = umma_desc_sw128(st + 2 * A_TILE_BYTES);   // [B_hi ; B_lo]
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            // advance the start address by k*32 B inside the 128-B swizzle atom
                            const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);
                            const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
                            if (L.use_b_lo)
                                umma_f16(tmem_d, da_hi + koff, db + koff, idesc_2n, accum);
                            else
                                umma_f16(tmem_d, da_hi + koff, db + koff, idesc_n, accum);
                            if (L.use_a_lo) umma_f16(tmem_d, da_lo + koff, db + koff, idesc_n, 1u);
                        }
                        umma_commit(&empty_bar[stage]);        // frees the smem slot when done
                        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
                    }
                    umma_commit(&tmem_full_bar[acc]);          // chunk ready for the epilogue
                    acc ^= 1;
                    if (acc == 0) acc_phase ^= 1u;
                }
            }
            if (L.dbg) {
                atomicAdd(&L.dbg[0], (unsigned long long)(clock64() - c_start));
                atomicAdd(&L.dbg[1], (unsigned long long)c_wait_op);
                atomicAdd(&L.dbg[2], (unsigned long long)c_wait_tm);
                atomicAdd(&L.dbg[3], (unsigned long long)c_kb);
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
        const int epi_tid = threadIdx.x - 64;
        int acc = 0;
        uint32_t acc_phase = 0;
        long long c_epi_wait = 0, c_epi_drain = 0, c_epi_store = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int g = tile / tiles_per_group;
            const int rem = tile - g * tiles_per_group;
            const int n_tile = rem / num_m_tiles;
            const int m_tile = rem - n_tile * num_m_tiles;
            const int n0 = g * L.n_per_group + n_tile * BLOCK_N;
            asm volatile("bar.sync 1, 128;" ::: "memory");   // previous tile's bias reads done
            for (int j = epi_tid; j < BLOCK_N; j += 128) bias_s[j] = L.bias[n0 + j];
            asm volatile("bar.sync 1, 128;" ::: "memory");

            // ---- drain the K-chunks into the fp32 running sum (registers) ----
            float sum[BLOCK_N];
#pragma unroll
            for (int j = 0; j < BLOCK_N; ++j) sum[j] = 0.f;
            for (int kb0 = 0; kb0 < kblocks; kb0 += L.chunk_kblocks) {
                long long t0 = clock64();
                mbar_wait(&tmem_full_bar[acc], acc_phase);
                const long long t1 = clock64();
                c_epi_wait += t1 - t0;
                tc_fence_after();
                const uint32_t taddr0 =
                    tmem_base + (uint32_t)(acc * C::ACC_STRIDE) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
                for (int c = 0; c < BLOCK_N / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(taddr0 + (uint32_t)(c * 32), r);
                    if (L.use_b_lo) {
                        uint32_t x[32];
                        tmem_ld_32x32b_x32(taddr0 + (uint32_t)(BLOCK_N + c * 32), x);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            sum[c * 32 + j] += __uint_as_float(r[j]) + __uint_as_float(x[j]);
                    } else {
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) sum[c * 32 + j] += __uint_as_float(r[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1u;
                c_epi_drain += clock64() - t1;
            }
            const long long t_store = clock64();

            const long long row = (long long)m_tile * BLOCK_M + quarter * 32 + lane;
            bool store = row < L.m_rows;
            if (L.pos_per_img > 0) {
                const int q = (int)(row % L.pos_per_img);
                const int y = q / L.grid_w, x = q - y * L.grid_w;
                store = store && (y < L.valid_h) && (x < L.valid_w);
            }
#pragma unroll
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float x = sum[c * 32 + j] + bias_s[c * 32 + j];
                    v[j] = L.relu ? fmaxf(x, 0.f) : x;
                }
                if (store) {
                    const long long off = row * (long long)L.ldc + n0 + c * 32;
                    if (L.out_f32) {
                        float4* o = reinterpret_cast<float4*>(L.out_f32 + off);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                    if (L.out_hi) {
                        uint32_t ph[16], pl[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const __half h0 = __float2half_rn(v[2 * j]);
                            const __half h1 = __float2half_rn(v[2 * j + 1]);
                            const __half l0 = __float2half_rn(v[2 * j] - __half2float(h0));
                            const __half l1 = __float2half_rn(v[2 * j + 1] - __half2float(h1));
                            ph[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                            pl[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                        }
                        uint4* oh = reinterpret_cast<uint4*>(L.out_hi + off);
                        uint4* ol = reinterpret_cast<uint4*>(L.out_lo + off);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            oh[j] = make_uint4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
                            ol[j] = make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
                        }
                    }
                }
            }
            c_epi_store += clock64() - t_store;
        }
        if (L.dbg && warp == 2 && lane == 0) {
            atomicAdd(&L.dbg[5], (unsigned long long)c_epi_wait);
            atomicAdd(&L.dbg[6], (unsigned long long)c_epi_drain);
            atomicAdd(&L.dbg[7], (unsigned long long)c_epi_store);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

template <int BLOCK_N>
int launch_impl(const GemmLayer& L, int num_sms, cudaStream_t stream) {
    using C = Cfg<BLOCK_N>;
    // function attributes are per device: remember which devices have been configured
    static std::mutex attr_mutex;
    static bool attr_done[64] = {};
    cudaError_t attr_err = cudaSuccess;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lock(attr_mutex);
        if (dev < 0 || dev >= 64 || !attr_done[dev]) {
            attr_err = cudaFuncSetAttribute(gemm_tc_kernel<BLOCK_N>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
            if (attr_err == cudaSuccess && dev >= 0 && dev < 64) attr_done[dev] = true;
        }
    }
    if (attr_err != cudaSuccess)
        return fail(-2, std::string("cudaFuncSetAttribute(gemm_tc_kernel): ") +
                            cudaGetErrorString(attr_err));
    const long long num_m_tiles = (L.m_rows + BLOCK_M - 1) / BLOCK_M;
    const long long total = num_m_tiles * (L.n_per_group / BLOCK_N) * L.groups;
    if (total <= 0) return 0;
    if (total > 0x7fffffffLL) return fail(-1, "gemm: too many tiles");
    const unsigned grid = (unsigned)(total < num_sms ? total : num_sms);
    gemm_tc_kernel<BLOCK_N><<<grid, GEMM_THREADS, C::SMEM_BYTES, stream>>>(L);
    SVX_LAUNCH_CHECK("gemm_tc_kernel");
    return 0;
}

}  // namespace

int launch_gemm_layer(const GemmLayer& L, int num_sms, cudaStream_t stream) {
    if (L.n_per_group % L.block_n != 0) return fail(-1, "gemm: n_per_group % block_n != 0");
    if (L.taps < 1 || L.taps > GEMM_MAX_TAPS) return fail(-1, "gemm: bad tap count");
    if ((L.out_hi == nullptr) != (L.out_lo == nullptr)) return fail(-1, "gemm: hi/lo outputs must pair");
    if (L.m_rows + BLOCK_M >= 0x7fffffffLL) return fail(-1, "gemm: too many rows for int32 TMA coordinates");
    if (L.chunk_kblocks < 1) return fail(-1, "gemm: chunk_kblocks must be >= 1");
    switch (L.block_n) {
        case 64: return launch_impl<64>(L, num_sms, stream);
        case 96: return launch_impl<96>(L, num_sms, stream);
        case 128: return launch_impl<128>(L, num_sms, stream);
        default: return fail(-1, "gemm: unsupported block_n (64, 96 or 128)");
    }
}

// ---- tensor maps ------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
                cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

int make_tensor_map_2d(CUtensorMap* tm, const void* base, long long rows, long long cols,
                       long long ld, int box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return fail(-2, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15))
        return fail(-1, "tensor map: base/stride must be 16-byte aligned");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(-2, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 0;
}

}  // namespace svx
