// CTA-pair (cta_group::2) variant of the tensor-core layer kernel.  Same computation, operand
// layouts, numerics and epilogue contract as conv_tc.cu (read that header first); what changes
// is how operands reach the tensor cores.
//
// Why.  Per-role cycle counters (svx_debug_counters) showed the 1-CTA kernel is bound by shared-
// memory bandwidth, not by MMA issue, L2 or HBM: every tcgen05.mma with M=128 re-reads its whole
// A and B tiles from shared memory, and UMMA operand reads + TMA writes come to ~154 B/clk per
// SM against ~128 available, capping the tensor pipe at ~63 % active.  With cta_group::2 the two
// SMs of a TPC execute ONE M=256 MMA: each reads only its own 128 A rows and HALF of the weight
// tile (N/2 rows), and each TMA-loads only that half.  Shared-memory traffic per MAC drops ~40 %
// (N=256: 24 KB per k-step per CTA for twice the MACs), which puts the layer under the limit.
//
// Structure (cluster of 2 CTAs, persistent over pair-tiles of 256 rows x BLOCK_N columns):
//   * warp 0 of BOTH CTAs: TMA producer for its own A slab (rows m0 + 128*rank ...) and its own
//     half of the weight tile; completion bytes are credited to the LEADER's mbarriers
//     (cp.async.bulk.tensor ... .cta_group::2 with the peer bit of the barrier address cleared);
//   * warp 1 of the leader: issues tcgen05.mma.cta_group::2 (3 per k-step: hi*hi, hi*lo, lo*hi
//     into one accumulator); tcgen05.commit ... .multicast::cluster releases the smem slots and
//     signals the accumulator in both CTAs;
//   * warps 2-9 of BOTH CTAs: epilogue over the CTA's own TMEM (128 rows): warp = (lane quarter,
//     column half), so running sums are BLOCK_N/2 <= 128 registers per thread; both CTAs arrive
//     on the leader's TMEM-empty barrier (mapa + mbarrier.arrive.shared::cluster).  Output goes
//     through a small per-warp shared-memory staging tile (16 columns per pass) so that every
//     global store instruction writes whole 32-byte sectors of few lines (thread-per-row stores
//     touched 32 lines per instruction and made the per-tile store phase, ~7k cycles, the
//     bottleneck of the first pair version); the tile is kept small (20 KB per CTA) so the weight
//     ring keeps 4+ stages.
#include "common.cuh"
#include "kernels.h"

#include <mutex>

namespace svx {

namespace {

constexpr int BLOCK_M = GEMM_BLOCK_M;              // rows per CTA; the pair covers 256
constexpr int BLOCK_K = GEMM_BLOCK_K;
constexpr int UMMA_K = 16;
constexpr int PAIR_THREADS = 320;                  // warp 0 TMA, warp 1 MMA/alloc, warps 2-9 epilogue
constexpr int EPI_THREADS = 256;
constexpr int MAX_SLAB_SLOTS = 4;
constexpr int MAX_B_STAGES = 8;
// epilogue staging: one 32-row tile per epilogue warp, STG columns per pass.  STG = 32: rows of
// 128 B payload + 16 B pad (whole 128-byte lines per store); STG = 16: 64 B + 16 B (half the
// shared memory, chosen where it buys the weight ring a 4th stage)
constexpr int SMEM_TOTAL = 222 * 1024;
__host__ __device__ constexpr int stage_row_bytes(int stg) { return stg == 32 ? 144 : (stg == 16 ? 80 : 48); }
__host__ __device__ constexpr int stage_bytes(int stg) { return 8 * 32 * stage_row_bytes(stg); }
__host__ __device__ constexpr int operand_budget(int stg) { return SMEM_TOTAL - stage_bytes(stg); }
constexpr int TMEM_COLS = 512;
constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO, version, SW128

__device__ __forceinline__ uint64_t make_desc(uint32_t lo) {
    return ((uint64_t)DESC_HI << 32) | (uint64_t)lo;
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) {
    return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16);
}

template <int BLOCK_N, int PASSES, int STG, bool DBG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ GemmLayer L) {
    constexpr int STAGE_ROW_BYTES = stage_row_bytes(STG);
    constexpr int STAGE_WARP_BYTES = 32 * STAGE_ROW_BYTES;
    constexpr int SMEM_OPERAND_BUDGET = operand_budget(STG);
    constexpr int HALF_N = BLOCK_N / 2;                        // weight rows held by each CTA
    constexpr int B_PLANE_BYTES = HALF_N * BLOCK_K * 2;
    constexpr bool A_LO = PASSES == 3;
    constexpr bool B_LO = PASSES >= 2;
    // TMEM accumulator ring: tiles of <= 128 columns get four buffers, so the MMA warp can run four
    // K = 256 chunks ahead of the epilogue (its per-tile store phase is the longest for the
    // small-K layers: conv2 has only 20 k-blocks per 256 x 128 fp32 tile)
    // (L.acc_bufs: 2 or 4; 4 only where the tile fits 128 TMEM columns)
    const int NUM_ACC = (BLOCK_N <= 128 && L.acc_bufs == 4) ? 4 : 2;
    const int ACC_STRIDE = TMEM_COLS / NUM_ACC;                // TMEM columns per buffer
    constexpr int COLS_PER_THREAD = BLOCK_N / 2;               // epilogue: column half per warp set
    static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 256, "BLOCK_N");
    static_assert(B_PLANE_BYTES % 1024 == 0, "weight half-tile must be whole swizzle atoms");
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_s[MAX_SLAB_SLOTS], empty_s[MAX_SLAB_SLOTS];
    __shared__ uint64_t full_b[MAX_B_STAGES], empty_b[MAX_B_STAGES];
    __shared__ uint64_t tmem_full_bar[4], tmem_empty_bar[4];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float bias_s[BLOCK_N];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t raw_addr = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

    const int slab_plane = L.slab_rows * 128;
    const int slab_slot_bytes = slab_plane * (A_LO ? 2 : 1);
    constexpr int b_stage_bytes = B_PLANE_BYTES * (B_LO ? 2 : 1);
    uint8_t* smem_b = smem + L.n_slab_slots * slab_slot_bytes;
    uint8_t* smem_stage = smem + SMEM_OPERAND_BUDGET;          // 1024-aligned: budget is a multiple of 1 KB

    const int num_mp_tiles = (int)((L.m_rows + 2 * BLOCK_M - 1) / (2 * BLOCK_M));
    const int n_tiles = L.n_per_group / BLOCK_N;
    const int tiles_per_group = num_mp_tiles * n_tiles;
    const int total_tiles = tiles_per_group * L.groups;
    const int kblocks = L.taps * L.cblocks;
    const int chunk = L.chunk_kblocks;
    const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&L.tm_a_hi);
        tma_prefetch_desc(&L.tm_a_lo);
        tma_prefetch_desc(&L.tm_b_hi);
        tma_prefetch_desc(&L.tm_b_lo);
        for (int s = 0; s < MAX_SLAB_SLOTS; ++s) { mbar_init(&full_s[s], 1); mbar_init(&empty_s[s], 1); }
        for (int s = 0; s < MAX_B_STAGES; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
        for (int b = 0; b < 4; ++b) {
            mbar_init(&tmem_full_bar[b], 1);
            mbar_init(&tmem_empty_bar[b], 16);        // 8 epilogue warps x 2 CTAs (leader's copy is used)
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc_pair(&tmem_base_smem, TMEM_COLS);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                // peer barriers initialised before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int slot = 0, stage = 0;
        uint32_t slot_phase = 0, phase = 0;
        long long c_prod_wait = 0, t0 = 0;
        for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
            const int g = tile / tiles_per_group;
            const int rem = tile - g * tiles_per_group;
            const int n_tile = rem / num_mp_tiles;
            const int mp_tile = rem - n_tile * num_mp_tiles;
            const int m0 = mp_tile * 2 * BLOCK_M + (int)rank * BLOCK_M;       // this CTA's rows
            const int n0 = g * L.n_per_group + n_tile * BLOCK_N + (int)rank * HALF_N;   // its weight rows
            const int a_col0 = g * L.a_group_cols;
            const int a_row0 = g * L.a_group_rows + L.a_row_bias + L.off_min;
            for (int cb = 0; cb < L.cblocks; ++cb) {
                if (DBG) t0 = clock64();
                mbar_wait(&empty_s[slot], slot_phase ^ 1u);
                if (DBG) c_prod_wait += clock64() - t0;
                if (elect_one()) {
                    uint8_t* sl = smem + slot * slab_slot_bytes;
                    if (leader) mbar_arrive_expect_tx(&full_s[slot], 2u * (uint32_t)slab_slot_bytes);
                    tma_load_2d_pair(&L.tm_a_hi, &full_s[slot], sl, a_col0 + cb * BLOCK_K, m0 + a_row0);
                    if (A_LO)
                        tma_load_2d_pair(&L.tm_a_lo, &full_s[slot], sl + slab_plane, a_col0 + cb * BLOCK_K,
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
                        if (leader) mbar_arrive_expect_tx(&full_b[stage], 2u * (uint32_t)b_stage_bytes);
                        tma_load_2d_pair(&L.tm_b_hi, &full_b[stage], sb, kcol, n0);
                        if (B_LO) tma_load_2d_pair(&L.tm_b_lo, &full_b[stage], sb + B_PLANE_BYTES, kcol, n0);
                    }
                    __syncwarp();
                    if (++stage == L.n_b_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
        if (DBG && lane == 0 && L.dbg) atomicAdd(&L.dbg[4], (unsigned long long)c_prod_wait);
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            constexpr uint32_t idesc = umma_idesc_f16_m256(BLOCK_N);
            int slot = 0, stage = 0, acc = 0;
            uint32_t slot_phase = 0, phase = 0, acc_phase = 0;
            long long c_wait_op = 0, c_wait_tm = 0, c_kb = 0, c_start = 0, t0 = 0;
            if (DBG) c_start = clock64();
            // The time this warp spends between two k-blocks is on the critical path of the N = 128
            // layers (12 MMAs of 64 clk each per k-block), so everything loop-invariant lives in
            // registers: barrier addresses, the weight ring's descriptor base, the slab plane offset.
            // (through an opaque move: otherwise the compiler re-derives each address where it is
            // used, S2R SR_CgaCtaId included, instead of keeping it)
            auto keep = [](uint32_t v) { uint32_t r; asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v)); return r; };
            const uint32_t adr_full_b = keep(smem_u32(&full_b[0])), adr_empty_b = keep(smem_u32(&empty_b[0]));
            const uint32_t adr_full_s = keep(smem_u32(&full_s[0])), adr_empty_s = keep(smem_u32(&empty_s[0]));
            const uint32_t adr_tm_full = keep(smem_u32(&tmem_full_bar[0])), adr_tm_empty = keep(smem_u32(&tmem_empty_bar[0]));
            const uint32_t b_ring_lo = keep(desc_lo(smem_u32(smem_b)));
            const uint32_t slab_ring_lo = keep(desc_lo(smem_u32(smem)));
            const uint32_t slab_slot_lo = (uint32_t)(slab_slot_bytes >> 4);
            const uint32_t b_stage_lo = (uint32_t)(b_stage_bytes >> 4);
            const uint32_t a_lo_plane = (uint32_t)(slab_plane >> 4);
            const int off_min8 = L.off_min * 8;
            for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
                int in_chunk = 0, kb = 0;
                for (int cb = 0; cb < L.cblocks; ++cb) {
                    if (DBG) t0 = clock64();
                    mbar_wait_addr(adr_full_s + 8u * (uint32_t)slot, slot_phase);
                    if (DBG) c_wait_op += clock64() - t0;
                    const uint32_t slab_lo = slab_ring_lo + (uint32_t)slot * slab_slot_lo - (uint32_t)off_min8;
                    const int nk = (cb + 1 == L.cblocks) ? L.last_ksteps : BLOCK_K / UMMA_K;
                    for (int t = 0; t < L.taps; ++t) {
                        if (in_chunk == 0) {
                            if (DBG) t0 = clock64();
                            mbar_wait_addr(adr_tm_empty + 8u * (uint32_t)acc, acc_phase ^ 1u);
                            if (DBG) c_wait_tm += clock64() - t0;
                        }
                        const uint32_t a_hi32 = slab_lo + (uint32_t)(L.row_off[t] * 8);
                        const uint32_t a_lo32 = a_hi32 + a_lo_plane;
                        const uint32_t b_hi32 = b_ring_lo + (uint32_t)stage * b_stage_lo;
                        const uint32_t b_lo32 = b_hi32 + (uint32_t)(B_PLANE_BYTES >> 4);
                        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC_STRIDE);
                        if (DBG) t0 = clock64();
                        mbar_wait_addr(adr_full_b + 8u * (uint32_t)stage, phase);
                        if (DBG) { c_wait_op += clock64() - t0; ++c_kb; }
                        tc_fence_after();
                        ++kb;
                        const bool chunk_end = (in_chunk + 1 == chunk) || (kb == kblocks);
                        if (elect_one()) {
                            auto kstep = [&](int k) {
                                const uint32_t ko = (uint32_t)(k * UMMA_K * 2 / 16);
                                const uint32_t accum = (in_chunk > 0 || k > 0) ? 1u : 0u;
                                umma_f16_pair(tmem_d, make_desc(a_hi32 + ko), make_desc(b_hi32 + ko), idesc, accum);
                                if (B_LO)
                                    umma_f16_pair(tmem_d, make_desc(a_hi32 + ko), make_desc(b_lo32 + ko), idesc, 1u);
                                if (A_LO)
                                    umma_f16_pair(tmem_d, make_desc(a_lo32 + ko), make_desc(b_hi32 + ko), idesc, 1u);
                            };
                            // straight-line code for the common full block: no per-k-step test
                            kstep(0);
                            if (nk == BLOCK_K / UMMA_K) {
                                kstep(1); kstep(2); kstep(3);
                            } else {
                                for (int k = 1; k < nk; ++k) kstep(k);
                            }
                            umma_commit_pair_addr(adr_empty_b + 8u * (uint32_t)stage);
                            if (t + 1 == L.taps) umma_commit_pair_addr(adr_empty_s + 8u * (uint32_t)slot);
                            if (chunk_end) umma_commit_pair_addr(adr_tm_full + 8u * (uint32_t)acc);
                        }
                        __syncwarp();
                        if (++stage == L.n_b_stages) { stage = 0; phase ^= 1u; }
                        if (chunk_end) {
                            in_chunk = 0;
                            if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1u; }
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
        }
    } else {
        // ===================== epilogue (warps 2..9, both CTAs) =====================
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;              // which half of the tile's columns
        const int epi_tid = threadIdx.x - 64;
        const int col0 = half * COLS_PER_THREAD;
        int acc = 0;
        uint32_t acc_phase = 0;
        long long c_epi_wait = 0, c_epi_drain = 0, c_epi_store = 0, t0 = 0, t1 = 0;
        for (int tile = pair_id; tile < total_tiles; tile += num_pairs) {
            const int g = tile / tiles_per_group;
            const int rem = tile - g * tiles_per_group;
            const int n_tile = rem / num_mp_tiles;
            const int mp_tile = rem - n_tile * num_mp_tiles;
            const int n0 = g * L.n_per_group + n_tile * BLOCK_N;
            asm volatile("bar.sync 1, 256;" ::: "memory");   // previous tile's bias reads done
            for (int j = epi_tid; j < BLOCK_N; j += EPI_THREADS) bias_s[j] = L.bias[n0 + j];
            asm volatile("bar.sync 1, 256;" ::: "memory");

            float sum[COLS_PER_THREAD];
#pragma unroll
            for (int j = 0; j < COLS_PER_THREAD; ++j) sum[j] = 0.f;
            for (int kb0 = 0; kb0 < kblocks; kb0 += chunk) {
                if (DBG) t0 = clock64();
                mbar_wait(&tmem_full_bar[acc], acc_phase);
                if (DBG) { t1 = clock64(); c_epi_wait += t1 - t0; }
                tc_fence_after();
                const uint32_t taddr0 = tmem_base + (uint32_t)(acc * ACC_STRIDE + col0) +
                                        ((uint32_t)(quarter * 32) << 16);
#pragma unroll
                for (int c = 0; c < COLS_PER_THREAD / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(taddr0 + (uint32_t)(c * 32), r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) sum[c * 32 + j] += __uint_as_float(r[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);    // leader's barrier
                if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1u; }
                if (DBG) c_epi_drain += clock64() - t1;
            }
            if (DBG) t0 = clock64();

            // ---- bias + ReLU, then out through the warp's staging tile: whole lines per store ----
            const long long row0 = (long long)mp_tile * 2 * BLOCK_M + (long long)rank * BLOCK_M + quarter * 32;
            uint8_t* stg = smem_stage + (warp - 2) * STAGE_WARP_BYTES;
            // valid-row mask of the warp's 32 rows (bit i = row0 + i is stored)
            uint32_t row_mask;
            {
                const long long row = row0 + lane;
                bool ok = row < L.m_rows;
                if (L.pos_per_img > 0) {
                    const int q = (int)(row % L.pos_per_img);
                    const int y = q / L.grid_w, x = q - y * L.grid_w;
                    ok = ok && (y < L.valid_h) && (x < L.valid_w);
                }
                row_mask = __ballot_sync(0xffffffffu, ok);
            }
            if constexpr (STG == 32) {
#pragma unroll
            for (int c = 0; c < COLS_PER_THREAD / 32; ++c) {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float x = sum[c * 32 + j] + bias_s[col0 + c * 32 + j];
                    v[j] = L.relu ? fmaxf(x, 0.f) : x;
                }
                const long long gcol = n0 + col0 + c * 32;
                uint4* my = reinterpret_cast<uint4*>(stg + lane * STAGE_ROW_BYTES);
                if (L.out_f32) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        my[j] = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                           __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
                    __syncwarp();
                    // 8 lanes cover one row's 128 B; one instruction writes 4 whole lines
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int r = it * 4 + (lane >> 3), ch = lane & 7;
                        if ((row_mask >> r) & 1u) {
                            const uint4 val = *reinterpret_cast<const uint4*>(stg + r * STAGE_ROW_BYTES + ch * 16);
                            *reinterpret_cast<uint4*>(L.out_f32 + (row0 + r) * (long long)L.ldc + gcol + ch * 4) = val;
                        }
                    }
                    __syncwarp();
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
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        my[j] = make_uint4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);       // hi: bytes 0..63
                        my[4 + j] = make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);   // lo: bytes 64..127
                    }
                    __syncwarp();
                    // 4 lanes cover one row's 64 B of a plane; one instruction writes 8 rows x 64 B
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int r = it * 8 + (lane >> 2), ch = lane & 3;
                        if ((row_mask >> r) & 1u) {
                            const long long o = (row0 + r) * (long long)L.ldc + gcol + ch * 8;
                            *reinterpret_cast<uint4*>(L.out_hi + o) =
                                *reinterpret_cast<const uint4*>(stg + r * STAGE_ROW_BYTES + ch * 16);
                            *reinterpret_cast<uint4*>(L.out_lo + o) =
                                *reinterpret_cast<const uint4*>(stg + r * STAGE_ROW_BYTES + 64 + ch * 16);
                        }
                    }
                    __syncwarp();
                }
            }
            } else if constexpr (STG == 8) {
            // 8 columns per pass (32 B payload + 16 B pad per row): the smallest staging tile, chosen where
            // it buys the weight ring a 5th stage (conv2: two 63.5 KB slab slots leave 75-83 KB)
#pragma unroll
            for (int c = 0; c < COLS_PER_THREAD / 8; ++c) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float x = sum[c * 8 + j] + bias_s[col0 + c * 8 + j];
                    v[j] = L.relu ? fmaxf(x, 0.f) : x;
                }
                const long long gcol = n0 + col0 + c * 8;
                uint4* my = reinterpret_cast<uint4*>(stg + lane * STAGE_ROW_BYTES);
                if (L.out_f32) {
                    my[0] = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
                    my[1] = make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7]));
                    __syncwarp();
                    // 2 lanes cover one row's 32 B; one instruction writes 16 rows x 1 whole sector
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const int r = it * 16 + (lane >> 1), ch = lane & 1;
                        if ((row_mask >> r) & 1u) {
                            const uint4 val = *reinterpret_cast<const uint4*>(stg + r * STAGE_ROW_BYTES + ch * 16);
                            *reinterpret_cast<uint4*>(L.out_f32 + (row0 + r) * (long long)L.ldc + gcol + ch * 4) = val;
                        }
                    }
                    __syncwarp();
                }
                if (L.out_hi) {                                  // one 16-byte store per row and plane
                    uint32_t ph[4], pl[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __half h0 = __float2half_rn(v[2 * j]);
                        const __half h1 = __float2half_rn(v[2 * j + 1]);
                        const __half l0 = __float2half_rn(v[2 * j] - __half2float(h0));
                        const __half l1 = __float2half_rn(v[2 * j + 1] - __half2float(h1));
                        ph[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                        pl[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                    }
                    if ((row_mask >> lane) & 1u) {
                        const long long o = (row0 + lane) * (long long)L.ldc + gcol;
                        *reinterpret_cast<uint4*>(L.out_hi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                        *reinterpret_cast<uint4*>(L.out_lo + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    }
                }
            }
            } else {
#pragma unroll
            for (int c = 0; c < COLS_PER_THREAD / 16; ++c) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float x = sum[c * 16 + j] + bias_s[col0 + c * 16 + j];
                    v[j] = L.relu ? fmaxf(x, 0.f) : x;
                }
                const long long gcol = n0 + col0 + c * 16;
                uint4* my = reinterpret_cast<uint4*>(stg + lane * STAGE_ROW_BYTES);
                if (L.out_f32) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        my[j] = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                           __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
                    __syncwarp();
                    // 4 lanes cover one row's 64 B; one instruction writes 8 rows x 2 whole sectors
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int r = it * 8 + (lane >> 2), ch = lane & 3;
                        if ((row_mask >> r) & 1u) {
                            const uint4 val = *reinterpret_cast<const uint4*>(stg + r * STAGE_ROW_BYTES + ch * 16);
                            *reinterpret_cast<uint4*>(L.out_f32 + (row0 + r) * (long long)L.ldc + gcol + ch * 4) = val;
                        }
                    }
                    __syncwarp();
                }
                if (L.out_hi) {
                    uint32_t ph[8], pl[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const __half h0 = __float2half_rn(v[2 * j]);
                        const __half h1 = __float2half_rn(v[2 * j + 1]);
                        const __half l0 = __float2half_rn(v[2 * j] - __half2float(h0));
                        const __half l1 = __float2half_rn(v[2 * j + 1] - __half2float(h1));
                        ph[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                        pl[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                    }
                    my[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]);       // hi: bytes 0..31
                    my[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
                    my[2] = make_uint4(pl[0], pl[1], pl[2], pl[3]);       // lo: bytes 32..63
                    my[3] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
                    __syncwarp();
                    // 2 lanes cover one row's 32 B of a plane; one instruction writes 16 rows x 1 sector
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const int r = it * 16 + (lane >> 1), ch = lane & 1;
                        if ((row_mask >> r) & 1u) {
                            const long long o = (row0 + r) * (long long)L.ldc + gcol + ch * 8;
                            *reinterpret_cast<uint4*>(L.out_hi + o) =
                                *reinterpret_cast<const uint4*>(stg + r * STAGE_ROW_BYTES + ch * 16);
                            *reinterpret_cast<uint4*>(L.out_lo + o) =
                                *reinterpret_cast<const uint4*>(stg + r * STAGE_ROW_BYTES + 32 + ch * 16);
                        }
                    }
                    __syncwarp();
                }
            }
            }
            if (DBG) c_epi_store += clock64() - t0;
        }
        if (DBG && L.dbg && leader && warp == 2 && lane == 0) {
            atomicAdd(&L.dbg[5], (unsigned long long)c_epi_wait);
            atomicAdd(&L.dbg[6], (unsigned long long)c_epi_drain);
            atomicAdd(&L.dbg[7], (unsigned long long)c_epi_store);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                // both CTAs done with TMEM / remote barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, TMEM_COLS);
    }
}

template <int BLOCK_N, int PASSES, int STG, bool DBG>
int launch_impl(const GemmLayer& L, int num_sms, cudaStream_t stream) {
    constexpr int smem_bytes = SMEM_TOTAL + 1024;
    // function attributes are per device: remember which devices have been configured
    static std::mutex attr_mutex;
    static bool attr_done[64] = {};
    cudaError_t attr_err = cudaSuccess;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lock(attr_mutex);
        if (dev < 0 || dev >= 64 || !attr_done[dev]) {
            attr_err = cudaFuncSetAttribute(conv_tc2_kernel<BLOCK_N, PASSES, STG, DBG>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
            if (attr_err == cudaSuccess && dev >= 0 && dev < 64) attr_done[dev] = true;
        }
    }
    if (attr_err != cudaSuccess)
        return fail(-2, std::string("cudaFuncSetAttribute(conv_tc2_kernel): ") + cudaGetErrorString(attr_err));
    const long long num_mp_tiles = (L.m_rows + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    const long long total = num_mp_tiles * (L.n_per_group / BLOCK_N) * L.groups;
    if (total <= 0) return 0;
    if (total > 0x7fffffffLL) return fail(-1, "conv2: too many tiles");
    long long pairs = num_sms / 2;
    if (pairs > total) pairs = total;
    conv_tc2_kernel<BLOCK_N, PASSES, STG, DBG><<<(unsigned)(2 * pairs), PAIR_THREADS, smem_bytes, stream>>>(L);
    SVX_LAUNCH_CHECK("conv_tc2_kernel");
    return 0;
}

template <int BLOCK_N, int STG>
int launch_passes_stg(const GemmLayer& L, int num_sms, cudaStream_t stream) {
    const int passes = L.use_a_lo ? 3 : (L.use_b_lo ? 2 : 1);
    if (L.dbg) {
        switch (passes) {
            case 3: return launch_impl<BLOCK_N, 3, STG, true>(L, num_sms, stream);
            case 2: return launch_impl<BLOCK_N, 2, STG, true>(L, num_sms, stream);
            default: return launch_impl<BLOCK_N, 1, STG, true>(L, num_sms, stream);
        }
    }
    switch (passes) {
        case 3: return launch_impl<BLOCK_N, 3, STG, false>(L, num_sms, stream);
        case 2: return launch_impl<BLOCK_N, 2, STG, false>(L, num_sms, stream);
        default: return launch_impl<BLOCK_N, 1, STG, false>(L, num_sms, stream);
    }
}

template <int BLOCK_N>
int launch_passes(const GemmLayer& L, int num_sms, cudaStream_t stream) {
    if constexpr (BLOCK_N == 128) {
        if (L.stage_cols == 8) return launch_passes_stg<BLOCK_N, 8>(L, num_sms, stream);
    }
    return L.stage_cols == 16 ? launch_passes_stg<BLOCK_N, 16>(L, num_sms, stream)
                              : launch_passes_stg<BLOCK_N, 32>(L, num_sms, stream);
}

}  // namespace

// Pipeline depths for the pair kernel: each CTA holds its own slab and HALF of the weight tile.
int plan_slab_pair(GemmLayer& L) {
    int lo = L.row_off[0], hi = L.row_off[0];
    for (int t = 1; t < L.taps; ++t) {
        lo = L.row_off[t] < lo ? L.row_off[t] : lo;
        hi = L.row_off[t] > hi ? L.row_off[t] : hi;
    }
    L.off_min = lo;
    L.slab_rows = ((BLOCK_M + (hi - lo)) + 7) & ~7;
    if (L.slab_rows > 256) return fail(-1, "conv2: tap span too large for one TMA box (slab_rows > 256)");
    if (L.use_a_lo && !L.use_b_lo) return fail(-1, "conv2: unsupported pass combination");
    const int slot = L.slab_rows * 128 * (L.use_a_lo ? 2 : 1);
    const int stage = (L.block_n / 2) * BLOCK_K * 2 * (L.use_b_lo ? 2 : 1);
    L.n_slab_slots = L.taps == 1 ? 3 : 2;
    // wide (32-column) epilogue staging unless the narrow one buys the weight ring a 4th stage
    L.stage_cols = 32;
    int nb = (operand_budget(32) - L.n_slab_slots * slot) / stage;
    if (nb < 4) {
        const int nb16 = (operand_budget(16) - L.n_slab_slots * slot) / stage;
        if (nb16 > nb) { nb = nb16; L.stage_cols = 16; }
        // the 8-column tile only where it buys yet another stage (128-column tiles only)
        const int nb8 = (operand_budget(8) - L.n_slab_slots * slot) / stage;
        if (L.block_n == 128 && nb8 > nb && L.allow_stg8) { nb = nb8; L.stage_cols = 8; }
    }
    if (nb > MAX_B_STAGES) nb = MAX_B_STAGES;
    if (nb < 2) return fail(-1, "conv2: shared memory budget too small for this layer");
    L.n_b_stages = nb;
    L.acc_bufs = L.block_n <= 128 ? 4 : 2;
    L.use_slab = 2;
    return 0;
}

int launch_conv_layer_pair(const GemmLayer& L, int num_sms, cudaStream_t stream) {
    if (L.use_slab != 2) return fail(-1, "conv2: layer was not planned for the pair kernel");
    if (L.n_per_group % L.block_n != 0) return fail(-1, "conv2: n_per_group % block_n != 0");
    if (L.taps < 1 || L.taps > GEMM_MAX_TAPS) return fail(-1, "conv2: bad tap count");
    if ((L.out_hi == nullptr) != (L.out_lo == nullptr)) return fail(-1, "conv2: hi/lo outputs must pair");
    if (L.m_rows + 4 * BLOCK_M >= 0x7fffffffLL) return fail(-1, "conv2: too many rows for int32 TMA coordinates");
    if (L.chunk_kblocks < 1) return fail(-1, "conv2: chunk_kblocks must be >= 1");
    if (L.desc_base_offset_mode != 0) return fail(-1, "conv2: descriptor base_offset mode must be 0");
    if (num_sms < 2) return fail(-1, "conv2: needs at least one SM pair");
    switch (L.block_n) {
        case 128: return launch_passes<128>(L, num_sms, stream);
        case 192: return launch_passes<192>(L, num_sms, stream);
        case 256: return launch_passes<256>(L, num_sms, stream);
        default: return fail(-1, "conv2: unsupported block_n (128, 192 or 256)");
    }
}

}  // namespace svx
