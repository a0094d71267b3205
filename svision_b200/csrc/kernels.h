// Internal launcher interface between the C-ABI (svx_api.cu) and the kernels.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace svx {

// ---- encoder.cu -------------------------------------------------------------------------------
// mode 0: NHWC fp32, 1: NHWC fp16, 2: conv1 operand layout [n][57*57][64] fp16
int launch_encode(const int32_t* rows_dev, long long n, void* out, int mode, int num_sms,
                  cudaStream_t stream);

// ---- front.cu ---------------------------------------------------------------------------------
// Fused encode + conv1 + ReLU + pool1 + LRN1 for encoder-generated images (sparse update form).
struct FrontParams {
    const float* w255;         // [11][11][3][96] = 255 * conv1 weights (TF layout)
    const float* base;         // [96] = bias + sum_k lo_ch * W[k]  (conv1 of the all-background image)
    __half* x2_hi;             // conv2 operand, see DESIGN.md §3
    __half* x2_lo;
    float* scratch;            // [scratch_blocks][640][96] conv1 values of dirty positions (L2-resident)
    int scratch_blocks;        // upper bound for the grid
    int ld;                    // elements per position row (128 padded layout, 48 packed layout)
    long long group_elems;     // element offset of channel group 1 (64 padded, plane size packed)
};
int launch_front(const int32_t* rows_dev, long long n, const FrontParams& P, int num_sms,
                 cudaStream_t stream);

// ---- layer_tc.cu ------------------------------------------------------------------------------
constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_MAX_TAPS = 25;

// One tensor-core layer expressed as a "shifted GEMM":
//   D[m, g*Ng + n] = sum_t sum_c  A[m + row_off[t], g*a_group_cols + c] * W[g*Ng + n, t*Cg + c]
// A: [rows_a][lda] fp16 (hi and lo planes), W: [n_total][k_total] fp16 K-major (hi, lo planes).
// The A tensor maps have box rows = slab_rows, the weight maps box rows = block_n / 2 (plan_layer).
struct GemmLayer {
    CUtensorMap tm_a_hi, tm_a_lo, tm_b_hi, tm_b_lo;
    int block_n;               // 96 / 128 / 192 / 256: columns per pair tile
    int chunk_kblocks;         // k-blocks accumulated in TMEM before promotion to fp32 registers
    int groups;                // 1 or 2
    int n_per_group;           // output channels per group (multiple of block_n)
    int taps;                  // filter taps (1 for fc)
    int cblocks;               // 64-wide channel blocks per tap
    int a_group_cols;          // column offset of group g in A
    int a_group_rows;          // row offset of group g in A (group-major planes); usually 0
    int a_row_bias;            // constant added to every A row coordinate (leading zero rows of the
                               // buffer, so overlapping-row maps never see a negative row)
    int last_ksteps;           // 16-wide k-steps issued in the LAST channel block of a tap (1..4):
                               // lets a tap's K be any multiple of 16 (zero tail never multiplied)
    int row_off[GEMM_MAX_TAPS];
    int use_a_lo, use_b_lo;    // which hi/lo cross terms are issued (3-pass / 2-pass / 1-pass)
    long long m_rows;          // rows of D actually computed (and of A addressable)
    // epilogue: bias + optional ReLU, then fp32, fp16 hi/lo planes, or the pooled buffer
    const float* bias;         // [groups * n_per_group]
    int relu;
    float* out_f32;            // [m_rows][ldc] or nullptr
    __half* out_hi;            // [m_rows][ldc] or nullptr
    __half* out_lo;
    int ldc;
    // valid-row mask for padded spatial layouts: row m is stored iff
    //   (m % pos_per_img) / grid_w < valid_h  &&  (m % pos_per_img) % grid_w < valid_w
    // pos_per_img == 0 disables the mask.
    int pos_per_img, grid_w, valid_h, valid_w;
    // pooled epilogue (set before plan_layer): ReLU, then the 3x3/2 VALID max-pool of the valid
    // positions into [img * pool_h*pool_w + py*pool_w + px][ldc] fp32; valid_h = 2*pool_h + 1,
    // valid_w = 2*pool_w + 1.  A window whose positions straddle a 128-row chunk boundary is
    // written in two parts: pool_out by the chunk of its first position, pool_out2 by the next one
    // (pool_crosses tells which windows have a second part).
    float* pool_out;
    float* pool_out2;
    int pool_h, pool_w;
    // filled by plan_layer: ONE slab [m0 + off_min, m0 + off_min + slab_rows) per (tile, channel
    // block) feeds every tap
    int planned;
    int slab_rows;             // multiple of 8, >= 128 + max(row_off) - min(row_off), <= 256
    int off_min;               // min(row_off)
    int n_slab_slots, n_b_stages;
    int acc_bufs;              // TMEM accumulator buffers (2, or 4 when block_n <= 128)
    int stage_cols;            // epilogue staging width (32 / 16 / 8; 0 = pooled epilogue)
    // optional cycle counters (development): 8 x unsigned long long, atomically accumulated per CTA
    //  0 MMA-role total  1 MMA wait operands  2 MMA wait TMEM-empty  3 k-blocks
    //  4 producer wait smem-empty  5 epilogue wait TMEM-full  6 epilogue drain  7 epilogue store
    unsigned long long* dbg;
};

// fills slab_rows / off_min / n_slab_slots / n_b_stages / stage_cols from taps, row_off, block_n,
// use_*_lo and pool_out
int plan_layer(GemmLayer& L);
int launch_layer(const GemmLayer& L, int num_sms, cudaStream_t stream);

// Builds a 2-D tiled fp16 tensor map (SWIZZLE_128B, box = {64, box_rows}) over a row-major
// [rows][cols] matrix with leading dimension ld (elements).
int make_tensor_map_2d(CUtensorMap* tm, const void* base, long long rows, long long cols,
                       long long ld, int box_rows);

// ---- cnn_aux.cu -------------------------------------------------------------------------------
struct PoolParams {
    const float* in;           // [n*in_pos_per_img][C] fp32 (post-ReLU conv output)
    int in_grid_w, in_pos_per_img;
    int C;
    int out_h, out_w;          // pooled size
    int lrn;                   // apply LRN(radius 2, alpha 2e-5, beta .75, bias 1) after pooling
    __half* out_hi;
    __half* out_lo;
    int out_ld;                // leading dimension of the output rows (elements)
    int out_grid_w, out_pos_per_img;   // output row = img*out_pos_per_img + y*out_grid_w + x
    int group_real;            // channels per output group
    long long group_elems;     // element offset between groups: channel c of row r lives at
                               //   (c / group_real) * group_elems + r * out_ld + c % group_real
};
int launch_pool(const PoolParams& p, long long n_img, int num_sms, cudaStream_t stream);

// Second half of a fused pool: window maxima (fp32, written by the conv epilogue in one or two
// parts) -> optional LRN -> fp16 hi/lo planes in the next layer's layout.  C = 256 channels.
// does the window of pooled position (img, py, px) have a second part in pool_out2?
__host__ __device__ inline bool pool_crosses(long long img, int py, int px, int pos_per_img, int grid_w) {
    const long long r0 = img * pos_per_img + (long long)(2 * py) * grid_w + 2 * px;   // first member
    return (int)(r0 & (GEMM_BLOCK_M - 1)) + 2 * grid_w + 2 >= GEMM_BLOCK_M;
}

struct FinishParams {
    const float* pooled;       // [n * per_img][256] fp32, compact (py * pool_w + px): first parts
    const float* pooled2;      // second parts of the windows that cross a chunk boundary
    int in_pos_per_img, in_grid_w;     // grid of the layer that was pooled (841 / 29, 196 / 14)
    int pool_h, pool_w;
    int lrn;                   // LRN(radius 2, alpha 2e-5, beta .75, bias 1): alexnet.py:164-166
    __half* out_hi;
    __half* out_lo;
    int out_ld;                // elements per output row
    int out_grid_w, out_pos_per_img;   // output row = img*out_pos_per_img + py*out_grid_w + px
};
int launch_finish_pooled(const FinishParams& p, long long n_img, int num_sms, cudaStream_t stream);

// Destinations of the per-site (label, score) calls written by the fc8 kernel: the local buffer
// and/or every rank's gathered buffer (peer-mapped), plus the completion signal of an exchange.
constexpr int CALL_MAX_SINKS = 16;
struct CallSinks {
    int2* ptr[CALL_MAX_SINKS];                 // where site 0 of THIS launch goes in each sink
    int count;                                 // 0: no calls are written
    unsigned long long* flag[CALL_MAX_SINKS];  // word to publish `epoch` to, per sink (with `done`)
    unsigned long long epoch;
    unsigned int* done;                        // CTA completion counter (local); nullptr: no signal
};

// logits = (x_hi + x_lo) @ W8 + b8 ; labels = argmax ; probs = softmax   (fp32, CUDA cores);
// labels / probs / logits may each be nullptr
int launch_fc8_softmax(const __half* x_hi, const __half* x_lo, const float* w8 /*[4096][5]*/,
                       const float* b8, long long n, int32_t* labels, float* probs, float* logits,
                       const CallSinks& sinks, cudaStream_t stream);
// waits (bounded) until every rank has published `epoch` in my_flags[0..world).  A rank that does not
// show up within timeout_ns is reported in *error (1 + rank; mapped host memory) and its `per_rank`
// calls in the local gathered buffer are poisoned (label -1, score NaN), so stale results of an
// earlier epoch cannot be mistaken for this one's.
int launch_exchange_wait(const unsigned long long* my_flags, int world, unsigned long long epoch,
                         unsigned long long timeout_ns, unsigned int* error, int2* gathered,
                         long long per_rank, cudaStream_t stream);

// NHWC [n][227][227][3] (fp32 or fp16) -> conv1 operand layout [n][57*57][64] fp16
int launch_nhwc_to_s2d(const void* images, int dtype, long long n, __half* out, cudaStream_t stream);

// float32 -> fp16 hi/lo planes (self-test helper)
int launch_split_hilo(const float* in, long long count, __half* hi, __half* lo, cudaStream_t stream);

}  // namespace svx
