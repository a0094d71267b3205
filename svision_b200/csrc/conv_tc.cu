// Tensor-core layer kernel of the SVision CNN (conv1..conv5, fc6, fc7) for sm_100a:
// TMA-fed tcgen05.mma with TMEM accumulators, warp-specialised, persistent, "slab" operand reuse.
//
// Replaces the TF-CPU kernels behind src/network/alexnet.py:100-155 (`conv`, `fc`):
// tf.nn.conv2d (+ the groups split/concat of :124-129), bias_add, relu, xw_plus_b.
//
// Every layer is the "shifted GEMM"
//     D[m, n] = sum_{tap t} sum_c  A[m + row_off[t], c] * W[n, t*Cg + c]
// over activation matrices [positions, channels] whose spatial zero padding is part of the layout
// (DESIGN.md §3), so a filter tap is a row offset and no im2col buffer exists.  The taps of one
// output tile read the row ranges [m0 + row_off[t], +128) of the same matrix, which overlap almost
// completely: warp 0 loads rows [m0 + off_min, m0 + off_min + slab_rows) ONCE per (tile,
// 64-channel block) with one TMA box (the "slab") and the MMA issuer addresses tap t as the
// 128-row window that starts at slab row (row_off[t] - off_min): the UMMA shared-memory descriptor
// start address moves in 128-byte steps.  Measured on B200: the 128-byte swizzle is a function of
// the absolute shared-memory address bits, so such row-shifted views need descriptor
// base_offset = 0 (mode 1, (addr >> 7) & 7, gives wrong results).  Weights (B) stream through
// their own mbarrier ring.  Rows outside the matrix are zero-filled by TMA.
//
// Numerics: fp16 hi/lo split operands, per 16-wide k-step
//   A_hi x [B_hi;B_lo]  as ONE N = 2*BLOCK_N MMA (main columns hi*hi, cross columns hi*lo) and
//   A_lo x B_hi         into the main columns                         (PASSES = 3);
// conv1's activations are exact in fp16 (PASSES = 2); PASSES = 1 is for comparison only.
// The tensor core's fp32 accumulation truncates (error grows linearly with chain length), so
// every `chunk_kblocks` k-blocks the TMEM accumulator is handed to the epilogue warps, which add
// it into an fp32 running sum in registers while the MMA warp continues in the other TMEM buffer.
//
// Roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2-5 epilogue.  Producer and
// issuer run warp-uniformly and issue under elect.sync: with `if (lane == 0)` the compiler wraps
// every UTCHMMA/UTMALDG in an ELECT/BRA.U.ANY serialisation loop, which made the issuing thread
// (~750 cycles per k-block) the bottleneck of the first version of this kernel.
#include "common.cuh"
#include "kernels.h"

#include <mutex>

namespace svx {

namespace {

constexpr int BLOCK_M = GEMM_BLOCK_M;
constexpr int BLOCK_K = GEMM_BLOCK_K;
constexpr int UMMA_K = 16;
constexpr int CONV_THREADS = 192;
constexpr int MAX_SLAB_SLOTS = 4;
constexpr int MAX_B_STAGES = 8;
constexpr int SMEM_OPERAND_BUDGET = 224 * 1024;      // + 1 KB alignment slack + ~1.5 KB static
constexpr int ACC_STRIDE = 256;                      // TMEM columns per buffer: main | cross
constexpr int TMEM_COLS = 512;
constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO, version, SW128

__device__ __forceinline__ uint64_t make_desc(uint32_t lo) {
    return ((uint64_t)DESC_HI << 32) | (uint64_t)lo;
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) {
    return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16);
}

template <int BLOCK_N, int PASSES, bool DBG>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_tc_kernel(const __grid_constant__ GemmLayer L) {
    constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;
    constexpr bool A_LO = PASSES == 3;
    constexpr bool B_LO = PASSES >= 2;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_s[MAX_SLAB_SLOTS], empty_s[MAX_SLAB_SLOTS];
    __shared__ uint64_t full_b[MAX_B_STAGES], empty_b[MAX_B_STAGES];
    __shared__ uint64_t tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float bias_s[BLOCK_N];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

    const int slab_plane = L.slab_rows * 128;                           // bytes of one hi/lo plane
    const int slab_slot_bytes = slab_plane * (A_LO ? 2 : 1);
    constexpr int b_stage_bytes = B_TILE_BYTES * (B_LO ? 2 : 1);
    uint8_t* smem_b = smem + L.n_slab_slots * slab_slot_bytes;

    const int num_m_tiles = (int)((L.m_rows + BLOCK_M - 1) / BLOCK_M);
    const int n_tiles = L.n_per_group / BLOCK_N;
    const int tiles_per_group = num_m_tiles * n_tiles;
    const int total_tiles = tiles_per_group * L.groups;
    const int kblocks = L.taps * L.cblocks;
    const int chunk = L.chunk_kblocks;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&L.tm_a_hi);
        tma_prefetch_desc(&L.tm_a_lo);
        tma_prefetch_desc(&L.tm_b_hi);
        tma_prefetch_desc(&L.tm_b_lo);
        for (int s = 0; s < MAX_SLAB_SLOTS; ++s) { mbar_init(&full_s[s], 1); mbar_init(&empty_s[s], 1); }
        for (int s = 0; s < MAX_B_STAGES; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 4); }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(&tmem_base_smem, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform, issue under elect.sync) ==============
        int slot = 0, stage = 0;
        uint32_t slot_phase = 0, phase = 0;
        long long c_prod_wait = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int g = tile / tiles_per_group;
            const int rem = tile - g * tiles_per_group;
            const int n_tile = rem / num_m_tiles;
            const int m_tile = rem - n_tile * num_m_tiles;
            const int m0 = m_tile * BLOCK_M;
            const int n0 = g * L.n_per_group + n_tile * BLOCK_N;
            const int a_col0 = g * L.a_group_cols;
            const int a_row0 = g * L.a_group_rows + L.a_row_bias + L.off_min;
            for (int cb = 0; cb < L.cblocks; ++cb) {
                long long t0 = 0;
                if (DBG) t0 = clock64();
                mbar_wait(&empty_s[slot], slot_phase ^ 1u);
                if (DBG) c_prod_wait += clock64() - t0;
                if (elect_one()) {
                    uint8_t* sl = smem + slot * slab_slot_bytes;
                    mbar_arrive_expect_tx(&full_s[slot], (uint32_t)slab_slot_bytes);
                    tma_load_2d(&L.tm_a_hi, &full_s[slot], sl, a_col0 + cb * BLOCK_K, m0 + a_row0);
                    if (A_LO)
                        tma_load_2d(&L.tm_a_lo, &full_s[slot], sl + slab_plane, a_col0 + cb * BLOCK_K,
                                    m0 + a_row0);
                }
                __syncwarp();
                if (++slot == L.n_slab_slots) { slot = 0; slot_phase ^= 1u; }
                int kcol = cb * BLOCK_K;
                const int kstep = L.cblocks * BLOCK_K;
                for (int t = 0; t < L.taps; ++t, kcol += kstep) {
                    if (DBG) t0 = clock64();
                    mbar_wait(&empty_b[stage], phase ^ 1u);
                    if (DBG) c_prod_wait += clock64() - t0;
                    if (elect_one()) {
                        uint8_t* sb = smem_b + stage * b_stage_bytes;
                        mbar_arrive_expect_tx(&full_b[stage], (uint32_t)b_stage_bytes);
                        tma_load_2d(&L.tm_b_hi, &full_b[stage], sb, kcol, n0);
                        if (B_LO) tma_load_2d(&L.tm_b_lo, &full_b[stage], sb + B_TILE_BYTES, kcol, n0);
                    }
                    __syncwarp();
                    if (++stage == L.n_b_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
        if (DBG && lane == 0 && L.dbg) atomicAdd(&L.dbg[4], (unsigned long long)c_prod_wait);
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform, issue under elect.sync) =================
        constexpr uint32_t idesc_n = umma_idesc_f16(BLOCK_N);
        constexpr uint32_t idesc_2n = umma_idesc_f16(2 * BLOCK_N);
        int slot = 0, stage = 0, acc = 0;
        uint32_t slot_phase = 0, phase = 0, acc_phase = 0;
        long long c_wait_op = 0, c_wait_tm = 0, c_kb = 0, c_start = 0, t0 = 0;
        if (DBG) c_start = clock64();
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int in_chunk = 0, kb = 0;
            for (int cb = 0; cb < L.cblocks; ++cb) {
                if (DBG) t0 = clock64();
                mbar_wait(&full_s[slot], slot_phase);
                if (DBG) c_wait_op += clock64() - t0;
                const uint32_t slab_lo = desc_lo(smem_u32(smem + slot * slab_slot_bytes));
                for (int t = 0; t < L.taps; ++t) {
                    if (in_chunk == 0) {
                        if (DBG) t0 = clock64();
                        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
                        if (DBG) c_wait_tm += clock64() - t0;
                    }
                    if (DBG) t0 = clock64();
                    mbar_wait(&full_b[stage], phase);
                    if (DBG) { c_wait_op += clock64() - t0; ++c_kb; }
                    tc_fence_after();
                    ++kb;
                    const bool chunk_end = (in_chunk + 1 == chunk) || (kb == kblocks);
                    if (elect_one()) {
                        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_STRIDE);
                        // slab row shift of this tap: 128 B per row -> 8 descriptor units per row
                        const uint32_t a_lo32 = slab_lo + (uint32_t)((L.row_off[t] - L.off_min) * 8);
                        const uint32_t b_lo32 = desc_lo(smem_u32(smem_b + stage * b_stage_bytes));
                        const int nk = (cb + 1 == L.cblocks) ? L.last_ksteps : BLOCK_K / UMMA_K;
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            if (k >= nk) break;
                            const uint32_t ko = (uint32_t)(k * UMMA_K * 2 / 16);   // 32 B per k-step
                            const uint64_t da = make_desc(a_lo32 + ko);
                            const uint64_t db = make_desc(b_lo32 + ko);
                            const uint32_t accum = (in_chunk > 0 || k > 0) ? 1u : 0u;
                            umma_f16(tmem_d, da, db, B_LO ? idesc_2n : idesc_n, accum);
                            if (A_LO)
                                umma_f16(tmem_d, make_desc(a_lo32 + (uint32_t)(slab_plane >> 4) + ko), db,
                                         idesc_n, 1u);
                        }
                        umma_commit(&empty_b[stage]);                 // B stage free when MMAs retire
                        if (t + 1 == L.taps) umma_commit(&empty_s[slot]);   // slab free
                        if (chunk_end) umma_commit(&tmem_full_bar[acc]);    // chunk -> epilogue
                    }
                    __syncwarp();
                    if (++stage == L.n_b_stages) { stage = 0; phase ^= 1u; }
                    if (chunk_end) {
                        in_chunk = 0;
                        acc ^= 1;
                        if (acc == 0) acc_phase ^= 1u;
                    } else {
                        ++in_chunk;
                    }
                }
                if (++slot == L.n_slab_slots) { slot = 0; slot_phase ^= 1u; }
            }
        }
        if (DBG && lane == 0 && L.dbg) {
            atomicAdd(&L.dbg[0], (unsigned long long)(clock64() - c_start));
            atomicAdd(&L.dbg[1], (unsigned long long)c_wait_op);
            atomicAdd(&L.dbg[2], (unsigned long long)c_wait_tm);
            atomicAdd(&L.dbg[3], (unsigned long long)c_kb);
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
        const int epi_tid = threadIdx.x - 64;
        int acc = 0;
        uint32_t acc_phase = 0;
        long long c_epi_wait = 0, c_epi_drain = 0, c_epi_store = 0, t0 = 0, t1 = 0;
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
            for (int kb0 = 0; kb0 < kblocks; kb0 += chunk) {
                if (DBG) t0 = clock64();
                mbar_wait(&tmem_full_bar[acc], acc_phase);
                if (DBG) { t1 = clock64(); c_epi_wait += t1 - t0; }
                tc_fence_after();
                const uint32_t taddr0 =
                    tmem_base + (uint32_t)(acc * ACC_STRIDE) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
                for (int c = 0; c < BLOCK_N / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(taddr0 + (uint32_t)(c * 32), r);
                    if (B_LO) {
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
                if (DBG) c_epi_drain += clock64() - t1;
            }
            if (DBG) t0 = clock64();

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
            if (DBG) c_epi_store += clock64() - t0;
        }
        if (DBG && L.dbg && warp == 2 && lane == 0) {
            atomicAdd(&L.dbg[5], (unsigned long long)c_epi_wait);
            atomicAdd(&L.dbg[6], (unsigned long long)c_epi_drain);
            atomicAdd(&L.dbg[7], (unsigned long long)c_epi_store);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int BLOCK_N, int PASSES, bool DBG>
int launch_impl(const GemmLayer& L, int num_sms, cudaStream_t stream) {
    constexpr int smem_bytes = SMEM_OPERAND_BUDGET + 1024;
    // function attributes are per device: remember which devices have been configured
    static std::mutex attr_mutex;
    static bool attr_done[64] = {};
    cudaError_t attr_err = cudaSuccess;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lock(attr_mutex);
        if (dev < 0 || dev >= 64 || !attr_done[dev]) {
            attr_err = cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, PASSES, DBG>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
            if (attr_err == cudaSuccess && dev >= 0 && dev < 64) attr_done[dev] = true;
        }
    }
    if (attr_err != cudaSuccess)
        return fail(-2, std::string("cudaFuncSetAttribute(conv_tc_kernel): ") + cudaGetErrorString(attr_err));
    const long long num_m_tiles = (L.m_rows + BLOCK_M - 1) / BLOCK_M;
    const long long total = num_m_tiles * (L.n_per_group / BLOCK_N) * L.groups;
    if (total <= 0) return 0;
    if (total > 0x7fffffffLL) return fail(-1, "conv: too many tiles");
    const unsigned grid = (unsigned)(total < num_sms ? total : num_sms);
    conv_tc_kernel<BLOCK_N, PASSES, DBG><<<grid, CONV_THREADS, smem_bytes, stream>>>(L);
    SVX_LAUNCH_CHECK("conv_tc_kernel");
    return 0;
}

template <int BLOCK_N>
int launch_passes(const GemmLayer& L, int num_sms, cudaStream_t stream) {
    const int passes = L.use_a_lo ? 3 : (L.use_b_lo ? 2 : 1);
    if (L.dbg) {
        switch (passes) {
            case 3: return launch_impl<BLOCK_N, 3, true>(L, num_sms, stream);
            case 2: return launch_impl<BLOCK_N, 2, true>(L, num_sms, stream);
            default: return launch_impl<BLOCK_N, 1, true>(L, num_sms, stream);
        }
    }
    switch (passes) {
        case 3: return launch_impl<BLOCK_N, 3, false>(L, num_sms, stream);
        case 2: return launch_impl<BLOCK_N, 2, false>(L, num_sms, stream);
        default: return launch_impl<BLOCK_N, 1, false>(L, num_sms, stream);
    }
}

}  // namespace

int plan_slab(GemmLayer& L) {
    int lo = L.row_off[0], hi = L.row_off[0];
    for (int t = 1; t < L.taps; ++t) {
        lo = L.row_off[t] < lo ? L.row_off[t] : lo;
        hi = L.row_off[t] > hi ? L.row_off[t] : hi;
    }
    L.off_min = lo;
    L.slab_rows = ((BLOCK_M + (hi - lo)) + 7) & ~7;
    if (L.slab_rows > 256) return fail(-1, "conv: tap span too large for one TMA box (slab_rows > 256)");
    if (L.use_a_lo && !L.use_b_lo) return fail(-1, "conv: unsupported pass combination");
    const int slot = L.slab_rows * 128 * (L.use_a_lo ? 2 : 1);
    const int stage = L.block_n * BLOCK_K * 2 * (L.use_b_lo ? 2 : 1);
    L.n_slab_slots = L.taps == 1 ? 3 : 2;
    int nb = (SMEM_OPERAND_BUDGET - L.n_slab_slots * slot) / stage;
    if (nb > MAX_B_STAGES) nb = MAX_B_STAGES;
    if (nb < 2) return fail(-1, "conv: shared memory budget too small for this layer");
    L.n_b_stages = nb;
    L.use_slab = 1;
    return 0;
}

int launch_conv_layer(const GemmLayer& L, int num_sms, cudaStream_t stream) {
    if (!L.use_slab) return fail(-1, "conv: layer was not planned for slab mode");
    if (L.n_per_group % L.block_n != 0) return fail(-1, "conv: n_per_group % block_n != 0");
    if (L.taps < 1 || L.taps > GEMM_MAX_TAPS) return fail(-1, "conv: bad tap count");
    if ((L.out_hi == nullptr) != (L.out_lo == nullptr)) return fail(-1, "conv: hi/lo outputs must pair");
    if (L.m_rows + 2 * BLOCK_M >= 0x7fffffffLL) return fail(-1, "conv: too many rows for int32 TMA coordinates");
    if (L.chunk_kblocks < 1) return fail(-1, "conv: chunk_kblocks must be >= 1");
    if (L.desc_base_offset_mode != 0) return fail(-1, "conv: descriptor base_offset mode 1 is wrong on sm_100 (measured)");
    if (L.n_slab_slots < 2 || L.n_slab_slots > MAX_SLAB_SLOTS || L.n_b_stages < 2 || L.n_b_stages > MAX_B_STAGES)
        return fail(-1, "conv: bad pipeline depths");
    switch (L.block_n) {
        case 64: return launch_passes<64>(L, num_sms, stream);
        case 96: return launch_passes<96>(L, num_sms, stream);
        case 128: return launch_passes<128>(L, num_sms, stream);
        default: return fail(-1, "conv: unsupported block_n (64, 96 or 128)");
    }
}

}  // namespace svx
