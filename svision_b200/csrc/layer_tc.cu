// The tensor-core layer kernel of the SVision CNN (conv1..conv5, fc6, fc7) for sm_100a: TMA-fed
// tcgen05.mma.cta_group::2 with TMEM accumulators, warp-specialised, persistent, "slab" operand
// reuse.  It is the ONLY tensor-core kernel of the library.
//
// Replaces the TF-CPU kernels behind src/network/alexnet.py:100-155 (`conv`, `fc`): tf.nn.conv2d
// (+ the groups split/concat of :124-129), bias_add, relu, xw_plus_b -- and, with the pooled
// epilogue, the max_pool of alexnet.py:158-161 that follows conv2 and conv5.
//
// Every layer is the "shifted GEMM"
//     D[m, n] = sum_{tap t} sum_c  A[m + row_off[t], c] * W[n, t*Cg + c]
// over activation matrices [positions, channels] whose spatial zero padding is part of the layout
// (DESIGN.md 3), so a filter tap is a row offset and no im2col buffer exists.  The taps of one
// output tile read the row ranges [m0 + row_off[t], +128) of the same matrix, which overlap almost
// completely: the producer loads rows [m0 + off_min, m0 + off_min + slab_rows) ONCE per (tile,
// 64-channel block) with one TMA box (the "slab") and the MMA issuer addresses tap t as the
// 128-row window that starts at slab row (row_off[t] - off_min): the UMMA shared-memory descriptor
// start address moves in 128-byte steps.  Measured on B200: the 128-byte swizzle is a function of
// the absolute shared-memory address bits, so such row-shifted views use descriptor
// base_offset = 0.  Weights (B) stream through their own mbarrier ring.  Rows outside the matrix
// are zero-filled by TMA.
//
// Why CTA pairs.  Per-role cycle counters (svx_debug_counters) showed a 1-CTA version (M = 128)
// bound by shared-memory bandwidth, not by MMA issue, L2 or HBM: every tcgen05.mma with M = 128
// re-reads its whole A and B tiles from shared memory, and UMMA operand reads + TMA writes came to
// ~154 B/clk per SM against ~128 available, capping the tensor pipe at ~63 % active.  With
// cta_group::2 the two SMs of a TPC execute ONE M = 256 MMA: each reads only its own 128 A rows
// and HALF of the weight tile (N/2 rows), and each TMA-loads only that half.
//
// Numerics: fp16 hi/lo split operands, per 16-wide k-step
// A_hi*B_hi + A_hi*B_lo + A_lo*B_hi into one fp32 TMEM accumulator (PASSES = 3); activations that
// are exact in fp16 (the dense conv1 of svx_forward) need PASSES = 2; PASSES = 1 exists for
// comparison only.  The tensor core's fp32 accumulation truncates (error grows linearly with the
// chain length), so every `chunk_kblocks` k-blocks the TMEM accumulator is handed to the epilogue
// warps, which add it into an fp32 running sum in registers while the MMA warp continues in
// another TMEM buffer.
//
// Structure (cluster of 2 CTAs, persistent over pair-tiles of 256 rows x BLOCK_N columns):
//   * warp 0 of BOTH CTAs: TMA producer for its own A slab (rows m0 + 128*rank ...) and its own
//     half of the weight tile; completion bytes are credited to the LEADER's mbarriers
//     (cp.async.bulk.tensor ... .cta_group::2 with the peer bit of the barrier address cleared);
//   * warp 1 of the leader: issues tcgen05.mma.cta_group::2 under elect.sync (with `if (lane == 0)`
//     ptxas wraps every UTCHMMA in an ELECT / BRA.U.ANY serialisation loop); tcgen05.commit ...
//     .multicast::cluster releases the smem slots and signals the accumulator in both CTAs;
//   * warps 2-9 of BOTH CTAs: epilogue over the CTA's own TMEM (128 rows): warp = (lane quarter,
//     column half), so running sums are BLOCK_N/2 <= 128 registers per thread; both CTAs arrive
//     on the leader's TMEM-empty barrier (mapa + mbarrier.arrive.shared::cluster).
//
// Epilogues (template STG):
//   * STG = 32 / 16 / 8: bias + ReLU, output through a small per-warp shared-memory staging tile
//     (STG columns per pass) so that every global store instruction writes whole 32-byte sectors
//     of few lines (thread-per-row stores touched 32 lines per instruction and made the per-tile
//     store phase the bottleneck of the first pair version); fp32 or fp16 hi/lo planes.  The
//     narrowest tile that still buys the weight ring another stage is chosen per layer.
//   * STG = 0 ("pooled"): bias + ReLU + the 3x3/2 VALID max-pool that follows the layer.  A CTA's
//     128 rows are 128 consecutive grid positions (4.4 grid rows of conv2, 9.1 of conv5): the four
//     warps of a column half share a staging tile, and every pooling window takes the max over its
//     members there.  A window cut by the chunk boundary is written in two parts (pool_out by the
//     chunk of its first member, pool_out2 by the next chunk) which the finishing pass combines.
//     Measured dead end: accumulating window rows with red.global.max.s32 into one zero-initialised
//     buffer (exact, since post-ReLU values are >= 0) costs ~1.3 clk per lane and SM -- 12k clk per
//     tile against 14k of MMA time, and still 4k with the in-chunk reduction.  The full-resolution
//     output never reaches HBM (conv2: 861 KB/site of fp32 before, <= 346 KB/site pooled now).
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
constexpr int SMEM_TOTAL = 222 * 1024;
// Epilogue staging, one tile per epilogue warp.  STG = 32: rows of 128 B payload + 16 B pad (whole
// 128-byte lines per store); 16: 64 B + 16 B; 8: 32 B + 16 B (chosen where the smaller tile buys the
// weight ring another stage).  STG = 0 (pooled): per column half (4 warps) two tiles of 128 rows x 32 B
// (double-buffered over the column passes) and the list of the chunk's pooling windows.
constexpr int POOL_MAX_WINDOWS = 96;               // per 128-row chunk: <= ~40 own + ~20 cut by the boundary
constexpr int POOL_TILE_BYTES = 128 * 8 * 4;        // T[128][8] fp32; two of them per column half
constexpr int POOL_HALF_BYTES = 2 * POOL_TILE_BYTES + POOL_MAX_WINDOWS * 8 + 16;
__host__ __device__ constexpr int stage_row_bytes(int stg) { return stg == 32 ? 144 : (stg == 16 ? 80 : 48); }
__host__ __device__ constexpr int stage_warp_bytes(int stg) {
    return 32 * stage_row_bytes(stg);
}
// rounded to 1 KB so that the staging area starts on a swizzle-atom boundary after the operands
__host__ __device__ constexpr int stage_bytes(int stg) {
    return ((stg == 0 ? 2 * POOL_HALF_BYTES : 8 * stage_warp_bytes(stg)) + 1023) & ~1023;
}
__host__ __device__ constexpr int operand_budget(int stg) { return SMEM_TOTAL - stage_bytes(stg); }
constexpr int TMEM_COLS = 512;
constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO, version, SW128

__device__ __forceinline__ uint64_t make_desc(uint32_t lo) {
    return ((uint64_t)DESC_HI << 32) | (uint64_t)lo;
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) {
    return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16);
}
__device__ __forceinline__ uint32_t pack_half2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

template <int BLOCK_N, int PASSES, int STG, bool DBG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
layer_tc_kernel(const __grid_constant__ GemmLayer L) {
    constexpr int STAGE_ROW_BYTES = stage_row_bytes(STG);
    constexpr int STAGE_WARP_BYTES = stage_warp_bytes(STG == 0 ? 8 : STG);
    constexpr int SMEM_OPERAND_BUDGET = operand_budget(STG);
    constexpr int HALF_N = BLOCK_N / 2;                        // weight rows held by each CTA
    constexpr int B_PLANE_BYTES = HALF_N * BLOCK_K * 2;
    constexpr bool A_LO = PASSES == 3;
    constexpr bool B_LO = PASSES >= 2;
    // TMEM accumulator ring: tiles of <= 128 columns get four buffers, so the MMA warp can run four
    // K = 256 chunks ahead of the epilogue (its per-tile store phase is the longest for the
    // small-K layers: conv2 has only 20 k-blocks per 256 x 128 tile)
    const int NUM_ACC = (BLOCK_N <= 128 && L.acc_bufs == 4) ? 4 : 2;
    const int ACC_STRIDE = TMEM_COLS / NUM_ACC;                // TMEM columns per buffer
    constexpr int COLS_PER_THREAD = BLOCK_N / 2;               // epilogue: column half per warp set
    constexpr int LDW = (COLS_PER_THREAD % 32 == 0) ? 32 : 16; // TMEM load width
    static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 256, "BLOCK_N");
    static_assert(B_PLANE_BYTES % 1024 == 0, "weight half-tile must be whole swizzle atoms");
    static_assert(COLS_PER_THREAD % (STG == 0 ? 8 : STG) == 0, "staging width must divide the column half");
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_s[MAX_SLAB_SLOTS], empty_s[MAX_SLAB_SLOTS];
    __shared__ uint64_t full_b[MAX_B_STAGES], empty_b[MAX_B_STAGES];
    __shared__ uint64_t tmem_full_bar[4], tmem_empty_bar[4];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float bias_s[BLOCK_N];

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
                for (int c = 0; c < COLS_PER_THREAD / LDW; ++c) {
                    uint32_t r[LDW];
                    if constexpr (LDW == 32) tmem_ld_32x32b_x32(taddr0 + (uint32_t)(c * 32), r);
                    else tmem_ld_32x32b_x16(taddr0 + (uint32_t)(c * 16), r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < LDW; ++j) sum[c * LDW + j] += __uint_as_float(r[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);    // leader's barrier
                if (++acc == NUM_ACC) { acc = 0; acc_phase ^= 1u; }
                if (DBG) c_epi_drain += clock64() - t1;
            }
            if (DBG) t0 = clock64();

            const long long row0 = (long long)mp_tile * 2 * BLOCK_M + (long long)rank * BLOCK_M + quarter * 32;
            uint8_t* stg = smem_stage + (warp - 2) * STAGE_WARP_BYTES;
            // this lane's row as a grid position
            bool ok;
            int gy = 0, gx = 0;
            uint32_t img = 0;
            {
                const long long row = row0 + lane;
                ok = row < L.m_rows;
                if (L.pos_per_img > 0) {
                    img = (uint32_t)row / (uint32_t)L.pos_per_img;               // rows < 2^31 (launch check)
                    const int q = (int)((uint32_t)row - img * (uint32_t)L.pos_per_img);
                    gy = q / L.grid_w;
                    gx = q - gy * L.grid_w;
                    ok = ok && (gy < L.valid_h) && (gx < L.valid_w);
                }
            }
            if constexpr (STG == 0) {
                // ---- bias + ReLU + 3x3/2 max-pool over this CTA's 128 rows (one column half = 4 warps) ----
                // valid extent = 2 * pooled extent + 1 in both directions (27 -> 13, 13 -> 6).
                // The 128 rows are 128 consecutive grid positions.  Per pass of 8 columns the four warps
                // put their values into a shared tile T[128][8]; every pooling window with a member in
                // the chunk takes the max over its members there.  A window spans 2*grid_w + 3 < 128
                // positions, so it has members in at most TWO chunks: the chunk that holds its first
                // member writes its (complete or partial) maximum to pool_out, the following chunk
                // writes the maximum of the remaining members to pool_out2.  Chunk starts are multiples
                // of 128 rows, so the pass that finishes the pool (finish_pooled_kernel) knows from the
                // window's row alone whether a second part exists.  Plain stores only.
                const int pw = L.pool_w, ph = L.pool_h, gw = L.grid_w;
                const int t = quarter * 32 + lane;                        // row within the chunk
                uint8_t* const half_base = smem_stage + half * POOL_HALF_BYTES;
                float* const T = reinterpret_cast<float*>(half_base);     // two buffers of [128][8]
                int2* const wlist = reinterpret_cast<int2*>(half_base + 2 * POOL_TILE_BYTES);
                int* const wcount = reinterpret_cast<int*>(half_base + 2 * POOL_TILE_BYTES + POOL_MAX_WINDOWS * 8);
                const int bar_id = 2 + half;
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // previous tile's list / tile reads done
                if (t == 0) *wcount = 0;
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                if (ok) {
                    // the (up to 2 x 2) windows this position belongs to
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        const int py = (gy >> 1) - a;
                        if (a == 0 ? py >= ph : ((gy & 1) || gy < 2)) continue;
#pragma unroll
                        for (int b = 0; b < 2; ++b) {
                            const int px = (gx >> 1) - b;
                            if (b == 0 ? px >= pw : ((gx & 1) || gx < 2)) continue;
                            const int dy = gy - 2 * py, dx = gx - 2 * px;
                            const int r0 = t - dy * gw - dx;              // chunk row of the window's first member
                            bool mine;
                            if (dy == 0 && dx == 0) {
                                mine = true;
                            } else if (r0 >= 0) {
                                mine = false;                             // listed by its first member
                            } else {
                                // second part of a window that starts in the previous chunk: listed by the
                                // first member inside this chunk
                                int first = -1;
#pragma unroll
                                for (int yy = 0; yy < 3; ++yy) {
                                    const int rr = r0 + yy * gw;
                                    if (first < 0 && rr + 2 >= 0) first = rr < 0 ? 0 : rr;
                                }
                                mine = first == t;
                            }
                            if (mine) {
                                const int slot = atomicAdd(wcount, 1);
                                if (slot < POOL_MAX_WINDOWS)
                                    wlist[slot] = make_int2(r0, (int)(((img * (uint32_t)(ph * pw) + (uint32_t)(py * pw + px)) << 1) |
                                                                      (r0 < 0 ? 1u : 0u)));
                            }
                        }
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                const int nwin = min(*wcount, POOL_MAX_WINDOWS);
                const int c4 = (t & 1) * 4;
#pragma unroll
                for (int c = 0; c < COLS_PER_THREAD / 8; ++c) {
                    float v[8];
                    {
                        const float4 b0 = *reinterpret_cast<const float4*>(&bias_s[col0 + c * 8]);
                        const float4 b1 = *reinterpret_cast<const float4*>(&bias_s[col0 + c * 8 + 4]);
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = ok ? fmaxf(sum[c * 8 + j] + bb[j], 0.f) : 0.f;
                    }
                    float* const Tc = T + (c & 1) * (POOL_TILE_BYTES / 4);
                    float4* my = reinterpret_cast<float4*>(Tc + t * 8);
                    my[0] = make_float4(v[0], v[1], v[2], v[3]);
                    my[1] = make_float4(v[4], v[5], v[6], v[7]);
                    // one barrier per pass: the buffer written now was last read two passes ago, and every
                    // thread has passed the barrier of the pass in between since
                    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
                    // 2 threads cover the 32 B of one window (16-byte loads and stores): 64 windows per
                    // round, i.e. one round
                    const long long gcol = n0 + col0 + c * 8 + c4;
                    for (int w = t >> 1; w < nwin; w += 64) {
                        const int2 e = wlist[w];
                        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int yy = 0; yy < 3; ++yy)
#pragma unroll
                            for (int xx = 0; xx < 3; ++xx) {
                                const int r = e.x + yy * gw + xx;
                                if ((unsigned)r < (unsigned)BLOCK_M) {
                                    const float4 q = *reinterpret_cast<const float4*>(Tc + r * 8 + c4);
                                    m.x = fmaxf(m.x, q.x); m.y = fmaxf(m.y, q.y);
                                    m.z = fmaxf(m.z, q.z); m.w = fmaxf(m.w, q.w);
                                }
                            }
                        float* const dst = (e.y & 1) ? L.pool_out2 : L.pool_out;
                        *reinterpret_cast<float4*>(dst + (long long)((uint32_t)e.y >> 1) * L.ldc + gcol) = m;
                    }
                }
            } else {
            // ---- bias + ReLU, then out through the warp's staging tile: whole sectors per store ----
            const uint32_t row_mask = __ballot_sync(0xffffffffu, ok);     // bit i = row0 + i is stored
            constexpr int LANES_PER_ROW = STG / 4;                         // 16-byte pieces per fp32 row
            constexpr int ROWS_PER_IT = 32 / LANES_PER_ROW;
#pragma unroll
            for (int c = 0; c < COLS_PER_THREAD / STG; ++c) {
                float v[STG];
#pragma unroll
                for (int j = 0; j < STG; ++j) {
                    const float x = sum[c * STG + j] + bias_s[col0 + c * STG + j];
                    v[j] = L.relu ? fmaxf(x, 0.f) : x;
                }
                const long long gcol = n0 + col0 + c * STG;
                uint4* my = reinterpret_cast<uint4*>(stg + lane * STAGE_ROW_BYTES);
                if (L.out_f32) {
#pragma unroll
                    for (int j = 0; j < STG / 4; ++j)
                        my[j] = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                           __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
                    __syncwarp();
                    // LANES_PER_ROW lanes cover one row's STG*4 bytes; one instruction writes
                    // ROWS_PER_IT rows of whole sectors
#pragma unroll
                    for (int it = 0; it < LANES_PER_ROW; ++it) {
                        const int r = it * ROWS_PER_IT + lane / LANES_PER_ROW, ch = lane % LANES_PER_ROW;
                        if ((row_mask >> r) & 1u) {
                            const uint4 val = *reinterpret_cast<const uint4*>(stg + r * STAGE_ROW_BYTES + ch * 16);
                            *reinterpret_cast<uint4*>(L.out_f32 + (row0 + r) * (long long)L.ldc + gcol + ch * 4) = val;
                        }
                    }
                    __syncwarp();
                }
                if (L.out_hi) {
                    uint32_t ph[STG / 2], pl[STG / 2];
#pragma unroll
                    for (int j = 0; j < STG / 2; ++j) {
                        const __half h0 = __float2half_rn(v[2 * j]);
                        const __half h1 = __float2half_rn(v[2 * j + 1]);
                        const __half l0 = __float2half_rn(v[2 * j] - __half2float(h0));
                        const __half l1 = __float2half_rn(v[2 * j + 1] - __half2float(h1));
                        ph[j] = pack_half2(h0, h1);
                        pl[j] = pack_half2(l0, l1);
                    }
                    if constexpr (STG == 8) {                    // one 16-byte store per row and plane
                        if (ok) {
                            const long long o = (row0 + lane) * (long long)L.ldc + gcol;
                            *reinterpret_cast<uint4*>(L.out_hi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                            *reinterpret_cast<uint4*>(L.out_lo + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                        }
                    } else {
                        // hi plane in the first STG*2 bytes of the staging row, lo plane in the next
                        constexpr int PIECES = STG / 8;          // 16-byte pieces per row and plane
#pragma unroll
                        for (int j = 0; j < PIECES; ++j) {
                            my[j] = make_uint4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
                            my[PIECES + j] = make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
                        }
                        __syncwarp();
                        constexpr int ROWS_H = 32 / PIECES;
#pragma unroll
                        for (int it = 0; it < PIECES; ++it) {
                            const int r = it * ROWS_H + lane / PIECES, ch = lane % PIECES;
                            if ((row_mask >> r) & 1u) {
                                const long long o = (row0 + r) * (long long)L.ldc + gcol + ch * 8;
                                *reinterpret_cast<uint4*>(L.out_hi + o) =
                                    *reinterpret_cast<const uint4*>(stg + r * STAGE_ROW_BYTES + ch * 16);
                                *reinterpret_cast<uint4*>(L.out_lo + o) =
                                    *reinterpret_cast<const uint4*>(stg + r * STAGE_ROW_BYTES + STG * 2 + ch * 16);
                            }
                        }
                        __syncwarp();
                    }
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
            attr_err = cudaFuncSetAttribute(layer_tc_kernel<BLOCK_N, PASSES, STG, DBG>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
            if (attr_err == cudaSuccess && dev >= 0 && dev < 64) attr_done[dev] = true;
        }
    }
    if (attr_err != cudaSuccess)
        return fail(-2, std::string("cudaFuncSetAttribute(layer_tc_kernel): ") + cudaGetErrorString(attr_err));
    const long long num_mp_tiles = (L.m_rows + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
    const long long total = num_mp_tiles * (L.n_per_group / BLOCK_N) * L.groups;
    if (total <= 0) return 0;
    if (total > 0x7fffffffLL) return fail(-1, "layer_tc: too many tiles");
    long long pairs = num_sms / 2;
    if (pairs > total) pairs = total;
    layer_tc_kernel<BLOCK_N, PASSES, STG, DBG><<<(unsigned)(2 * pairs), PAIR_THREADS, smem_bytes, stream>>>(L);
    SVX_LAUNCH_CHECK("layer_tc_kernel");
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
        if (L.stage_cols == 0) return launch_passes_stg<BLOCK_N, 0>(L, num_sms, stream);
        if (L.stage_cols == 8) return launch_passes_stg<BLOCK_N, 8>(L, num_sms, stream);
    }
    if constexpr ((BLOCK_N / 2) % 32 == 0) {
        if (L.stage_cols == 32) return launch_passes_stg<BLOCK_N, 32>(L, num_sms, stream);
    }
    if (L.stage_cols == 16) return launch_passes_stg<BLOCK_N, 16>(L, num_sms, stream);
    return fail(-1, "layer_tc: staging width not available for this tile width");
}

}  // namespace

// Pipeline depths and epilogue staging: each CTA holds its own slab and HALF of the weight tile.
int plan_layer(GemmLayer& L) {
    int lo = L.row_off[0], hi = L.row_off[0];
    for (int t = 1; t < L.taps; ++t) {
        lo = L.row_off[t] < lo ? L.row_off[t] : lo;
        hi = L.row_off[t] > hi ? L.row_off[t] : hi;
    }
    L.off_min = lo;
    L.slab_rows = ((BLOCK_M + (hi - lo)) + 7) & ~7;
    if (L.slab_rows > 256) return fail(-1, "layer_tc: tap span too large for one TMA box (slab_rows > 256)");
    if (L.use_a_lo && !L.use_b_lo) return fail(-1, "layer_tc: unsupported pass combination");
    const int slot = L.slab_rows * 128 * (L.use_a_lo ? 2 : 1);
    const int stage = (L.block_n / 2) * BLOCK_K * 2 * (L.use_b_lo ? 2 : 1);
    L.n_slab_slots = L.taps == 1 ? 3 : 2;
    int nb;
    if (L.pool_out) {
        if (L.block_n != 128) return fail(-1, "layer_tc: the pooled epilogue needs 128-column tiles");
        L.stage_cols = 0;
        nb = (operand_budget(0) - L.n_slab_slots * slot) / stage;
    } else if ((L.block_n / 2) % 32 != 0) {
        L.stage_cols = 16;                        // 48-column halves (the dense conv1's 96 channels)
        nb = (operand_budget(16) - L.n_slab_slots * slot) / stage;
    } else {
        // wide (32-column) epilogue staging unless a narrower one buys the weight ring another stage
        L.stage_cols = 32;
        nb = (operand_budget(32) - L.n_slab_slots * slot) / stage;
        if (nb < 4) {
            const int nb16 = (operand_budget(16) - L.n_slab_slots * slot) / stage;
            if (nb16 > nb) { nb = nb16; L.stage_cols = 16; }
            const int nb8 = (operand_budget(8) - L.n_slab_slots * slot) / stage;     // 128-column tiles only
            if (L.block_n == 128 && nb8 > nb) { nb = nb8; L.stage_cols = 8; }
        }
    }
    if (nb > MAX_B_STAGES) nb = MAX_B_STAGES;
    if (nb < 2) return fail(-1, "layer_tc: shared memory budget too small for this layer");
    L.n_b_stages = nb;
    L.acc_bufs = L.block_n <= 128 ? 4 : 2;
    L.planned = 1;
    return 0;
}

int launch_layer(const GemmLayer& L, int num_sms, cudaStream_t stream) {
    if (!L.planned) return fail(-1, "layer_tc: layer was not planned (plan_layer)");
    if (L.n_per_group % L.block_n != 0) return fail(-1, "layer_tc: n_per_group % block_n != 0");
    if (L.taps < 1 || L.taps > GEMM_MAX_TAPS) return fail(-1, "layer_tc: bad tap count");
    if ((L.out_hi == nullptr) != (L.out_lo == nullptr)) return fail(-1, "layer_tc: hi/lo outputs must pair");
    if (L.m_rows + 4 * BLOCK_M >= 0x7fffffffLL) return fail(-1, "layer_tc: too many rows for int32 TMA coordinates");
    if (L.chunk_kblocks < 1) return fail(-1, "layer_tc: chunk_kblocks must be >= 1");
    if (num_sms < 2) return fail(-1, "layer_tc: needs at least one SM pair");
    if (L.pool_out) {
        if (!L.pool_out2) return fail(-1, "layer_tc: the pooled epilogue needs both pooled buffers");
        if (!L.relu || L.pos_per_img <= 0 || L.valid_w != 2 * L.pool_w + 1 || L.valid_h != 2 * L.pool_h + 1)
            return fail(-1, "layer_tc: pooled epilogue needs ReLU and valid extent = 2 * pooled extent + 1");
        if ((L.m_rows / L.pos_per_img + 1) * (long long)(L.pool_h * L.pool_w) >= (1LL << 30))
            return fail(-1, "layer_tc: too many pooled rows for the epilogue's 30-bit row index");
        if (2 * L.grid_w + 3 >= BLOCK_M) return fail(-1, "layer_tc: pooling window spans more than one chunk boundary");
        // windows per 128-row chunk: those that start in it (every other grid row, pool_w each) plus those
        // cut by its upper boundary (started within the 2 * grid_w + 2 positions before it)
        const int own = ((BLOCK_M / L.grid_w) / 2 + 2) * L.pool_w;
        const int cut = (((2 * L.grid_w + 2) / L.grid_w) / 2 + 1) * L.pool_w;
        if (own + cut > POOL_MAX_WINDOWS) return fail(-1, "layer_tc: too many pooling windows per chunk for the epilogue's list");
    }
    switch (L.block_n) {
        case 96: return launch_passes<96>(L, num_sms, stream);
        case 128: return launch_passes<128>(L, num_sms, stream);
        case 192: return launch_passes<192>(L, num_sms, stream);
        case 256: return launch_passes<256>(L, num_sms, stream);
        default: return fail(-1, "layer_tc: unsupported block_n (96, 128, 192 or 256)");
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
