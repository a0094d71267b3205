/*
 * svx.h -- C-ABI of the B200-native SVision encode-and-classify path (libsvx.so).
 *
 * The reference has no plugin/FFI interface: the seam is
 *   Predict(chrom, segments_out_file).run(out_path_prefix, options)
 *                                         (reference: src/network/predict.py:15,148)
 * and inside it exactly two calls are replaced by this library:
 *   batch_generator.next_batch(batch_size)          src/network/predict.py:207
 *       -> src/network/create_batch.py:88-155  (-> src/segmentplot/plot_segment.py:33-73)
 *   sess.run([score, argmax, softmax], feed_dict)   src/network/predict.py:209-210
 *       -> src/network/alexnet.py:26-58
 * The ctypes binding a maintainer would add on the reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types;
 *   - every function returns 0 on success or a negative svx_status; the message of the last
 *     failure on the calling thread is returned by svx_last_error();
 *   - "rows" are packed candidate sites, int32[n][12]:
 *        xS1 xE1 yS1 yE1 f1  xS2 xE2 yS2 yE2 f2  len_a len_b
 *     (the 12 '_'-joined tokens of create_batch.py:45; f = 1 for the token 'True', 0 otherwise;
 *     xE is carried but ignored, as create_batch.py:106,121 ignore it);
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream);
 *   - device pointers must belong to the handle's device; the caller owns every buffer it
 *     passes; the library owns weights and workspaces; nothing is allocated on the hot path
 *     after svx_create;
 *   - a handle is not thread-safe and not fork-safe (create it in the process that uses it);
 *     calls on one handle are ordered by the library even when they use different streams
 *     (they share the workspaces), so they never overlap.
 */
#ifndef SVX_H_
#define SVX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVX_IMG 227            /* src/network/predict.py:167, create_batch.py:13 */
#define SVX_ROW_FIELDS 12
#define SVX_NUM_CLASSES 5      /* 0 DEL, 1 INS, 2 INV, 3 DUP, 4 tDUP: predict.py:133-142 */

typedef enum {
    SVX_OK = 0,
    SVX_ERR_INVALID = -1,      /* bad argument */
    SVX_ERR_CUDA = -2,         /* CUDA runtime / driver failure (message has the detail) */
    SVX_ERR_NOMEM = -3,
    SVX_ERR_UNSUPPORTED = -4   /* e.g. device is not sm_100 */
} svx_status;

/* element type of images written by svx_encode / read by svx_forward */
typedef enum {
    SVX_IMAGE_F32 = 0,         /* float32 NHWC [n][227][227][3]: what next_batch() yields and TF
                                  receives (create_batch.py:147-152, predict.py:167) */
    SVX_IMAGE_F16 = 1          /* IEEE half NHWC, same values (lossless: 2 levels per channel) */
} svx_image_dtype;

/* numeric recipe of the tensor-core layers */
typedef enum {
    SVX_PRECISION_3PASS = 0,   /* fp16 hi/lo split operands, 3 MMAs per product: parity default */
    SVX_PRECISION_1PASS = 1    /* single fp16 pass: reported for comparison only */
} svx_precision;

/* The reference's 16 TensorFlow variables (src/network/alexnet.py:113-116,141-145), host
 * float32, TF layouts: conv weights [kh][kw][Cin/groups][Cout], fc weights [in][out]. */
typedef struct {
    const float *conv1_w, *conv1_b;   /* [11][11][3][96],    [96]   */
    const float *conv2_w, *conv2_b;   /* [5][5][48][256],    [256]  groups=2 */
    const float *conv3_w, *conv3_b;   /* [3][3][256][384],   [384]  */
    const float *conv4_w, *conv4_b;   /* [3][3][192][384],   [384]  groups=2 */
    const float *conv5_w, *conv5_b;   /* [3][3][192][256],   [256]  groups=2 */
    const float *fc6_w, *fc6_b;       /* [9216][4096],       [4096] */
    const float *fc7_w, *fc7_b;       /* [4096][4096],       [4096] */
    const float *fc8_w, *fc8_b;       /* [4096][5],          [5]    */
} svx_weights;

typedef struct svx_handle svx_handle;

/* Replaces Predict.run's model setup (src/network/predict.py:155-189: graph build +
 * Saver.restore).  `weights` may be NULL for an encoder-only handle.  `max_batch` is the
 * micro-batch (sites resident on the device at once; workspaces are sized for it: ~1.4 MB of HBM per
 * site, 1..65536).  The buffers of the dense conv1 path are allocated by the first svx_forward. */
int svx_create(const svx_weights *weights, int device, int64_t max_batch, int precision,
               svx_handle **out);
void svx_destroy(svx_handle *h);

/* Replaces BatchGenerator.next_batch's per-image work (create_batch.py:103-152 ->
 * plot_segment.py:9-73): rows_dev int32[n][12] -> images_dev NHWC [n][227][227][3] of `dtype`.
 * Any n; asynchronous on `stream`. */
int svx_encode(svx_handle *h, const int32_t *rows_dev, int64_t n, void *images_dev, int dtype,
               void *stream);

/* Replaces sess.run(score) (predict.py:209 -> alexnet.py:26-58) on caller-provided images:
 * images_dev NHWC [n][227][227][3] of `dtype` -> logits_dev float32[n][5].  Arbitrary images take
 * the dense tcgen05 conv1 (1024 sites per pass); the classify entries below use the fused sparse
 * front end instead, which is exact only for images this library's encoder produces. */
int svx_forward(svx_handle *h, const void *images_dev, int dtype, int64_t n, float *logits_dev,
                void *stream);

/* The fused hot path on device buffers: rows_dev -> labels_dev int32[n] (argmax, predict.py:209),
 * probs_dev float32[n][5] (softmax), logits_dev float32[n][5] (may be NULL).  Images never
 * leave the device: the first tensor that reaches HBM is the conv2 operand.  Asynchronous. */
int svx_classify_device(svx_handle *h, const int32_t *rows_dev, int64_t n, int32_t *labels_dev,
                        float *probs_dev, float *logits_dev, void *stream);

/* The same path on HOST buffers (what a reference-side caller holds): copies rows to the
 * device, runs, copies labels/probs back, synchronises.  rows_host should be pinned for speed. */
int svx_classify(svx_handle *h, const int32_t *rows_host, int64_t n, int32_t *labels_host,
                 float *probs_host);

/* What the reference consumes per row downstream of the path: predict_value[i] (argmax) and
 * softmax_value[i][predict_value[i]] (src/network/predict.py:230,251). */
typedef struct {
    int32_t label;             /* 0 DEL, 1 INS, 2 INV, 3 DUP, 4 tDUP */
    float score;               /* softmax of that class */
} svx_call;

/* svx_classify_device, returning only the 8-byte call per site (written by the fc8 kernel). */
int svx_classify_device_calls(svx_handle *h, const int32_t *rows_dev, int64_t n, svx_call *calls_dev,
                              void *stream);

/* ---- several GPUs behind one handle, in ONE process -----------------------------------------------
 * The reference drives Step 2 from a single `SVision` process (SVision:296-341).  svx_multi gives
 * that caller every GPU of the box without torchrun: one svx_handle and one host thread per device;
 * the rows of a call are cut into chunks which the devices take from a shared counter (a slower GPU
 * takes fewer), each result is written straight to its place in the caller's arrays (file order is
 * preserved, no gather).  Same results as svx_classify, bit for bit: sites are independent.
 *   devices[ndev]   CUDA device ordinals, distinct
 *   max_batch       micro-batch per device (as svx_create)
 * svx_multi_last_split reports how many sites each device processed in the last call.  Like a single
 * handle, an svx_multi is externally synchronised: one svx_multi_classify at a time. */
typedef struct svx_multi svx_multi;
int svx_multi_create(const svx_weights *weights, const int *devices, int ndev, int64_t max_batch,
                     int precision, svx_multi **out);
int svx_multi_classify(svx_multi *m, const int32_t *rows_host, int64_t n, int32_t *labels_host,
                       float *probs_host);
int svx_multi_device_count(const svx_multi *m);
int svx_multi_last_split(const svx_multi *m, int64_t *sites_per_device /* [ndev] */);
void svx_multi_destroy(svx_multi *m);

/* ---- multi-GPU result exchange ---------------------------------------------------
 * The reference parallelises Step 2 with one process per chromosome and temp files
 * (SVision:311-323); here sites shard contiguously over one process per GPU, and the only exchange
 * is the per-site call.  svx_classify_exchange classifies this rank's shard and its fc8 kernel stores
 * every call straight into the gathered buffer of EVERY rank (peer-mapped over NVLink via CUDA IPC)
 * and publishes an epoch flag; a one-warp kernel then waits for the other ranks' flags.  No separate
 * collective runs.  Setup: every rank creates an exchange, exports its IPC handle blob, the
 * handles are all-gathered by the caller (torch.distributed in svision_b200/sharded.py) and attached.
 *   gathered_dev  device pointer to svx_call[world][sites_per_rank] (rank-major = file order for
 *                 contiguous shards), valid once the work queued on `stream` has finished and until
 *                 the next-but-one svx_classify_exchange (two buffers alternate).  Consume it on
 *                 `stream` (or after synchronising) before calling again.
 * All ranks must call svx_classify_exchange the same number of times.  A rank that does not show up
 * within SVX_EXCHANGE_TIMEOUT_MS (default 30000; covers start skew such as one rank still parsing its
 * BED) cannot hang the GPU: the wait kernel gives up, POISONS that rank's calls in the local gathered
 * buffer (label -1, score NaN -- stale results of an earlier epoch can not be mistaken for this one's)
 * and raises a sticky error word in mapped host memory.  The error is returned by svx_exchange_status
 * (which synchronises the device and clears it) and by every later svx_classify_exchange until then. */
#define SVX_IPC_HANDLE_BYTES 72   /* CUDA IPC handle (64) + offset of the buffer inside the exported allocation */
typedef struct svx_exchange svx_exchange;
int svx_exchange_create(svx_handle *h, int rank, int world, int64_t sites_per_rank, svx_exchange **out);
int svx_exchange_export(svx_exchange *x, void *ipc_handle_out /* SVX_IPC_HANDLE_BYTES */);
int svx_exchange_attach(svx_exchange *x, const void *ipc_handles /* [world][SVX_IPC_HANDLE_BYTES], rank order */);
int svx_classify_exchange(svx_handle *h, svx_exchange *x, const int32_t *rows_dev, int64_t n,
                          const svx_call **gathered_dev, void *stream);
int svx_exchange_status(svx_exchange *x);
void svx_exchange_destroy(svx_exchange *x);

/* Parity/debug: copy the activation named `name` of the LAST micro-batch (first `n` sites) to
 * host as float32, NHWC, valid positions only.  Names: "conv1" [55][55][96], "norm1"
 * [27][27][96], "norm2" [13][13][256], "conv3"/"conv4" [13][13][384], "pool5" [6][6][256],
 * "fc6"/"fc7" [4096].  ("conv1" exists only after svx_forward; the full-resolution outputs of conv2 and
 * conv5 are never materialised: their max-pool runs in the layer's epilogue.) */
int svx_debug_activation(svx_handle *h, const char *name, int64_t n, float *out_host);

/* Standalone tcgen05 GEMM self-test entry (used by tests): C[M][N] = A[M][K] * B[N][K]^T with
 * fp16 hi/lo operands given as float32 on the device; returns float32 C on the device.  block_n is
 * the tile width of the layer kernel: 96, 128, 192 or 256. */
int svx_gemm_selftest(int device, const float *a_dev, const float *b_dev, float *c_dev,
                      int64_t m, int64_t n, int64_t k, int block_n, int precision, void *stream);

/* Shifted-GEMM self-test: C[m][n] = sum_t sum_c A[m + row_off[t]][c] * B[n][t*k_per_tap + c]
 * (rows outside A read as zero), through the same layer kernel (layer_tc.cu). */
int svx_conv_selftest(int device, const float *a_dev, const float *b_dev, float *c_dev, int64_t m,
                      int64_t n, int64_t k_per_tap, int taps, const int *row_off, int block_n,
                      int precision, void *stream);

/* Milliseconds the layer kernel of the calling thread's last svx_gemm_selftest / svx_conv_selftest took
 * (CUDA events on its stream; the operand conversion around it is not included). */
float svx_selftest_last_ms(void);

/* Development aid: per-role cycle counters of the 7 tensor-core layers, uint64 out[7][8]
 * (handle created with SVX_DBG=1 in the environment -- besides SVX_EXCHANGE_TIMEOUT_MS the only
 * environment variable the library reads): 0 MMA-role total, 1 MMA wait operands,
 * 2 MMA wait TMEM-empty, 3 k-blocks, 4 producer wait smem-empty, 5 epilogue wait TMEM-full,
 * 6 epilogue drain, 7 epilogue store; summed over CTAs. */
int svx_debug_counters(svx_handle *h, uint64_t *out, int reset);

/* Per-kernel device timing with CUDA events recorded on the launching stream (bench.py's live
 * roofline numbers).  Slots: 0 encode, 1 conv1, 2 pool1+lrn1, 3 conv2, 4 pool2+lrn2, 5 conv3,
 * 6 conv4, 7 conv5, 8 pool5, 9 fc6, 10 fc7, 11 fc8+softmax.  svx_profile_read synchronises,
 * adds the elapsed milliseconds and launch counts since the last reset into ms_out[12] /
 * launches_out[12] (either may be NULL) and optionally resets.  Slot 3 includes pool2 and slot 7
 * pool5 (fused into the layer's epilogue); slots 4 and 8 are the passes that finish them (LRN2 +
 * fp16 split; fp16 split). */
#define SVX_PROFILE_SLOTS 12
int svx_set_profiling(svx_handle *h, int enable);
int svx_profile_read(svx_handle *h, float *ms_out, int64_t *launches_out, int reset);

/* Kernels launched by this library on the calling thread since the last reset (bench.py's
 * `gpu_launches`). */
int64_t svx_launch_count(void);
void svx_launch_count_reset(void);

/* ---- host side: <chrom>.segments.all.bed -> packed rows (no GPU involved) -------------------------
 * Replaces BatchGenerator.read_class_list (src/network/create_batch.py:29-61) and the token parsing
 * of next_batch (create_batch.py:103-137).  `text` is the whole file (23 tab-separated columns per
 * line, writer: src/collection/output_clusters.py:180-182,207-209); blank lines are skipped.
 *   rows  [n][12] int32  columns 1-12 packed as above (strand tokens -> 0/1)
 *   bkp   [n][3]  int64  columns 17, 18, 22 (breakpoint start, end, length: predict.py:222-224)
 *   spans [n][7][2] int64 byte offset and length, within `text`, of columns 0 (region), 13 (read id),
 *                        15 (qname), 16 (signature type), 19 (score), 20 (forward), 21 (mechanism)
 *   flags [n]     int32  SVX_BED_FLAG_* below
 * A malformed line fails the whole call (SVX_ERR_INVALID; the message names the line). */
#define SVX_BED_SPANS 7
#define SVX_BED_FLAG_MAIN 1         /* read id contains 'm': a main segment pair (predict.py:279) */
#define SVX_BED_FLAG_FORWARD 2      /* column 20 == "True" (predict.py:229) */
#define SVX_BED_FLAG_UNCOVERED 4    /* column 16 == "sigUncovered" (output.py:526) */
#define SVX_BED_FLAG_SAME_REGION 8  /* column 0 equals the previous row's (predict.py:235) */
#define SVX_BED_FLAG_COMPLEMENT 16  /* one of the label columns (0, 13, 15, 16, 19, 20, 21) contains the
                                       substring "complement": the reference skips the row whatever else
                                       it holds (predict.py:214 tests the joined label; its own pad rows
                                       are labelled 'complement-complement', create_batch.py:56) */
int svx_bed_count_rows(const char *text, int64_t len, int64_t *n_rows);
int svx_bed_parse(const char *text, int64_t len, int64_t n_rows, int32_t *rows, int64_t *bkp,
                  int64_t *spans, int32_t *flags);

/* ---- host side: signatures -> packed rows, skipping the text BED (no GPU involved) ---------------
 * Replaces Signature.get_segs_cords (src/collection/classes.py:72-117) + proc_one_sig / linearOrNot /
 * cal_non_linear (src/collection/output_clusters.py:11-27,124-251) for all signatures of a chromosome.
 *   sig_aln_off [n_sig+1]      alignments of signature s are aln[sig_aln_off[s] .. sig_aln_off[s+1])
 *   aln         [n_aln][5]     ref_start, ref_end, q_start, q_end, is_reverse: the fields of
 *                              `sorted_aligns` that get_segs_cords reads (absolute coordinates)
 *   sig_bkp_off [n_sig+1]      breakpoints of signature s (pair k of a main x inner combination uses
 *                              breakpoint k+1, a main pair breakpoint 0: output_clusters.py:173,199)
 *   rows        [capacity][12] packed rows as above
 *   meta        [capacity][5]  signature index, sub id (BED column 14), bit0 main pair | bit1 forward
 *                              (column 20), index into the breakpoint array, non-linear score (column 19)
 *   n_rows                     rows produced; with capacity 0 nothing is written and this is the count.
 * A signature whose segments span no reference base is skipped, as the reference skips it. */
#define SVX_ALN_FIELDS 5
#define SVX_PAIR_META 5
int svx_pairs_generate(int64_t n_sig, const int64_t *sig_aln_off, const int64_t *aln,
                       const int64_t *sig_bkp_off, int64_t capacity, int32_t *rows, int64_t *meta,
                       int64_t *n_rows);

/* ---- host side: labels -> candidate SVs per region (no GPU involved) -----------------------------
 * Replaces, for one chromosome, the per-row loop of Predict.run (src/network/predict.py:213-300),
 * Predict.get_region_potential_svtypes (predict.py:29-145) and the numeric part of
 * write_results_to_vcf (src/network/output.py:473-474,495-496,525-529,551-552).  Inputs are what
 * svx_bed_parse produced (text, spans, flags, the three breakpoint columns) plus the classifier's
 * labels and `win[i]` = round(softmax of the label, 2) as float32 (predict.py:251).
 *   cand   [capacity n][25] int64  per emitted candidate (support >= min_support), in file order:
 *            0 first kept row of its region, 1 support, 2 number of classes, 3 1 = filter "Uncovered",
 *            4 offset into `reads`, 5..9 class ids ascending, 10..24 breakpoints [class][start,end,len]
 *   qual   [capacity n] double     std(signature scores) / support + (1 - round(mean(win), 2)) * 100,
 *                                  before the min with 100 (output.py:474,551-552)
 *   reads  [capacity n] int64      per supporting read the LAST row that named it (its qname and
 *                                  signature score are that row's: predict.py:249,255)
 * numpy.mean over float32 and numpy.std over ints are restated operation by operation, because the
 * values are printed; svx_np_mean_f32 / svx_np_std_i64 expose the two for the parity tests. */
#define SVX_CAND_FIELDS 25
int svx_calls_aggregate(const char *text, int64_t len, int64_t n, const int64_t *spans, const int32_t *flags,
                        const int64_t *bkp_start, const int64_t *bkp_end, const int64_t *bkp_len,
                        const int32_t *labels, const float *win, int64_t min_support, int64_t *cand,
                        double *qual, int64_t *reads, int64_t *n_cand, int64_t *n_reads);
int svx_np_mean_f32(const float *values, int64_t n, float *out);
int svx_np_std_i64(const int64_t *values, int64_t n, double *out);

int64_t svx_max_batch(const svx_handle *h);
int svx_device(const svx_handle *h);
const char *svx_last_error(void);
const char *svx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SVX_H_ */
