// C-ABI of libsvx (see include/svx.h): handle, one-time weight repack, HBM workspaces, the
// per-micro-batch kernel sequence.  Host code here is plumbing; all arithmetic on the path runs in
// the kernels of encoder.cu / front.cu / layer_tc.cu / cnn_aux.cu.
#include "../../include/svx.h"

#include "common.cuh"
#include "kernels.h"

#include <cstdlib>
#include <cstring>
#include <memory>
#include <utility>
#include <vector>

namespace svx {

static thread_local std::string g_error;
static thread_local long long g_launches = 0;

void set_error(const std::string& msg) { g_error = msg; }
int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}
void count_launch(int n) { g_launches += n; }

// ---- geometry of the padded activation layouts (DESIGN.md "HBM layouts") ----------------------
constexpr int S2D = 57;                       // conv1 input grid (228/4), outputs valid on 55x55
constexpr int P1 = S2D * S2D;                 // 3249 positions per site in x1 / y1
constexpr int G2 = 29, P2 = G2 * G2;          // conv2 grid: 27 + 2 shared pad rows/cols -> 841
constexpr int G3 = 14, P3 = G3 * G3;          // conv3-5 grid: 13 + 1 -> 196

// K = 256 per TMEM accumulation chain: the tensor core's fp32 accumulation truncates, so longer
// chains cost precision.  conv3 / conv4 use 6: their 192-column tiles have only two
// TMEM buffers and a per-tile store phase (~9.3k clk) longer than two 4-block chunks of MMA work, so
// the MMA warp waited 140-170 clk per k-block for a free buffer; with K = 384 chunks it runs 13.8k clk
// ahead: 1242 -> 1155 clk per k-block (99.7 % of the issue bound) for max |dsoftmax| 1.08e-4 -> 1.33e-4
constexpr int kChunkKBlocks = 4;
constexpr int kChunkKBlocksWide = 6;
// sites per pass of the dense conv1 path (svx_forward only): its x1 / y1 buffers take 1.66 MB per
// site and are allocated on first use
constexpr long long kDenseBatch = 1024;
// workspaces are ~1.4 MB per site; beyond this the request cannot fit a 180 GB device anyway
constexpr long long kMaxBatchLimit = 65536;

enum { L_CONV1 = 0, L_CONV2, L_CONV3, L_CONV4, L_CONV5, L_FC6, L_FC7, L_COUNT };

struct LayerSpec {
    int taps, cg_real, cg_pad, groups, n_total, block_n, kh, kw, chunk;
};
static const LayerSpec kSpec[L_COUNT] = {
    /* conv1 (s2d) */ {9, 48, 64, 1, 96, 96, 3, 3, kChunkKBlocks},
    /* conv2       */ {25, 48, 64, 2, 256, 128, 5, 5, kChunkKBlocks},
    /* conv3       */ {9, 256, 256, 1, 384, 192, 3, 3, kChunkKBlocksWide},
    /* conv4       */ {9, 192, 192, 2, 384, 192, 3, 3, kChunkKBlocksWide},
    /* conv5       */ {9, 192, 192, 2, 256, 128, 3, 3, kChunkKBlocks},
    /* fc6         */ {1, 9216, 9216, 1, 4096, 256, 1, 1, kChunkKBlocks},
    /* fc7         */ {1, 4096, 4096, 1, 4096, 256, 1, 1, kChunkKBlocks},
};

}  // namespace svx

using namespace svx;

struct svx_handle {
    int device = 0;
    int num_sms = 148;
    long long max_batch = 0;
    int precision = 0;
    bool has_model = false;
    float* front_w255 = nullptr;
    float* front_base = nullptr;
    float* front_scratch = nullptr;      // [front_blocks][3025][96]
    int front_blocks = 0;
    // conv2 operand, K-packed: two group planes of [B*841 + 64][48]; the 5 kw-taps of a kernel row
    // are 240 contiguous values (5 row-taps, K = 1200)
    long long x2_group_elems = 0;        // element offset of channel group 1 in x2
    int x2_ld = 48;                      // elements per x2 position row
    cudaStream_t stream = nullptr;       // used by svx_classify (host entry)
    cudaEvent_t busy = nullptr;          // recorded at the end of every entry that uses the workspaces:
                                         // the next entry (possibly on another stream) waits on it
    long long last_n = 0;

    std::vector<void*> allocs;
    __half* w_hi[L_COUNT] = {};
    __half* w_lo[L_COUNT] = {};
    float* bias[L_COUNT] = {};
    float* w8 = nullptr;
    float* b8 = nullptr;

    // dense conv1 path (svx_forward only), allocated on first use for kDenseBatch sites
    long long dense_batch = 0;
    __half* x1 = nullptr;                                   // [dense_batch*3249][64]
    float* y1 = nullptr;                                    // [dense_batch*3249][96]
    __half *x2_hi = nullptr, *x2_lo = nullptr;              // [B*841][128]
    float *p2 = nullptr, *p2b = nullptr;                    // [B*169][256] pooled conv2: window maxima, two parts
    __half *x3_hi = nullptr, *x3_lo = nullptr;              // [B*196][256]
    __half *x4_hi = nullptr, *x4_lo = nullptr;              // [B*196][384]
    __half *x5_hi = nullptr, *x5_lo = nullptr;              // [B*196][384]
    float *p5 = nullptr, *p5b = nullptr;                    // [B*36][256] pooled conv5
    __half *x6_hi = nullptr, *x6_lo = nullptr;              // [B][9216]
    __half *x7_hi = nullptr, *x7_lo = nullptr;              // [B][4096]
    __half *x8_hi = nullptr, *x8_lo = nullptr;              // [B][4096]

    int32_t* rows_dev = nullptr;                            // staging for the host entry
    int32_t* labels_dev = nullptr;
    float* probs_dev = nullptr;

    GemmLayer layer[L_COUNT];
    unsigned long long* dbg = nullptr;   // [L_COUNT][8] cycle counters when SVX_DBG=1

    // optional per-kernel timing (svx_set_profiling)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;                       // reusable events
    std::vector<std::pair<int, cudaEvent_t>> ev_marks;      // (slot or -1 = end of sequence, event)
    size_t ev_used = 0;
    double prof_ms[SVX_PROFILE_SLOTS] = {};
    long long prof_launches[SVX_PROFILE_SLOTS] = {};
};

// Multi-GPU result exchange (include/svx.h): the local gathered buffer, peer mappings of every
// other rank's buffer (CUDA IPC), epoch-parity double buffering.
struct svx_exchange {
    svx_handle* h = nullptr;
    int device = 0;                                 // copy: destroy must not touch a freed handle
    int rank = 0, world = 1;
    long long per_rank = 0;
    void* local = nullptr;                          // cudaMalloc: [flags 256 B][region 0][region 1]
    void* base[CALL_MAX_SINKS] = {};                // base[r]: rank r's buffer as mapped here
    void* mapped[CALL_MAX_SINKS] = {};              // what cudaIpcOpenMemHandle returned (to close)
    bool opened[CALL_MAX_SINKS] = {};
    bool attached = false;
    unsigned int* done = nullptr;                   // fc8 CTA counter
    unsigned int* error_host = nullptr;             // wait-kernel timeout report: mapped pinned host word,
    unsigned int* error = nullptr;                  //   and its device alias (readable without a sync)
    unsigned long long epoch = 0;
    unsigned long long timeout_ns = 30000000000ull; // SVX_EXCHANGE_TIMEOUT_MS overrides
    size_t region_bytes() const { return (size_t)world * (size_t)per_rank * sizeof(svx_call); }
    char* region(int r, int parity) const {
        return static_cast<char*>(base[r]) + 256 + (size_t)parity * region_bytes();
    }
    unsigned long long* flags(int r) const { return static_cast<unsigned long long*>(base[r]); }
};

namespace {

// Is the primary context of `dev` alive in this process?  (driver entry point, resolved once)
bool primary_context_active(int dev) {
    typedef CUresult (*PFN_state)(CUdevice, unsigned int*, int*);
    static PFN_state fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuDevicePrimaryCtxGetState", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<PFN_state>(p);
    }();
    if (!fn) return true;                       // cannot tell: behave as before
    unsigned int flags = 0;
    int active = 0;
    return fn((CUdevice)dev, &flags, &active) == CUDA_SUCCESS ? active != 0 : true;
}

// Makes the handle's device current for the duration of a call and gives the caller's device back
// afterwards -- unless that would CREATE a context there: since CUDA 12 cudaSetDevice initialises the
// device's primary context, and a thread that never touched CUDA has device 0 as its "current" one.
// Restoring it blindly made every process that drives GPU 1..7 from a helper thread build a context on
// GPU 0 (measured: 1.85 s on the first call of the thread, and ~0.5 GB of GPU 0's memory per process).
struct DeviceGuard {
    int prev = -1, cur = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) : cur(dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != cur && primary_context_active(prev)) cudaSetDevice(prev);
    }
};

template <typename T>
int dev_alloc(svx_handle* h, T** p, size_t count, bool zero = true) {
    void* q = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess)
        return fail(SVX_ERR_NOMEM, std::string("cudaMalloc(") + std::to_string(bytes) +
                                       " B): " + cudaGetErrorString(e));
    h->allocs.push_back(q);
    if (zero) SVX_CUDA_CHECK(cudaMemset(q, 0, bytes));
    *p = static_cast<T*>(q);
    return 0;
}

// TF-layout weights -> K-major [n_total][taps*cg_pad] fp16 hi/lo planes on the device.
// K layout of the packed conv2: [kh][kw*48 + c] with each kernel row padded 240 -> 256
constexpr int kPackTaps = 5, kPackK = 256;
// zero rows in front of the packed x2: a virtual (overlapping) row with a NEGATIVE index would be
// zero-filled by TMA as a whole although most of its 256 elements are valid data
constexpr int kPackLeadRows = 64;

int upload_layer_weights(svx_handle* h, int li, const float* w_tf, const float* b_tf) {
    const LayerSpec& s = kSpec[li];
    const bool packed = li == L_CONV2;
    const size_t K = packed ? (size_t)kPackTaps * kPackK : (size_t)s.taps * s.cg_pad;
    const size_t count = (size_t)s.n_total * K;
    std::vector<__half> hi(count), lo(count);
    std::memset(hi.data(), 0, count * sizeof(__half));
    std::memset(lo.data(), 0, count * sizeof(__half));
    auto put = [&](size_t n, size_t k, float w) {
        const __half hh = __float2half_rn(w);
        hi[n * K + k] = hh;
        lo[n * K + k] = __float2half_rn(w - __half2float(hh));
    };
    if (li == L_CONV1) {
        // 11x11/4 conv as a 3x3/1 conv over the 4x4 space-to-depth image:
        // tap (a,b), channel (dy*4+dx)*3+c  <-  W[4a+dy][4b+dx][c][n]  (zero beyond 10)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                for (int dy = 0; dy < 4; ++dy)
                    for (int dx = 0; dx < 4; ++dx) {
                        const int kh = 4 * a + dy, kw = 4 * b + dx;
                        if (kh > 10 || kw > 10) continue;
                        for (int c = 0; c < 3; ++c)
                            for (int n = 0; n < 96; ++n)
                                put(n, (size_t)(a * 3 + b) * 64 + (dy * 4 + dx) * 3 + c,
                                    w_tf[((size_t)(kh * 11 + kw) * 3 + c) * 96 + n]);
                    }
    } else if (packed) {
        for (int kh = 0; kh < 5; ++kh)
            for (int kw = 0; kw < 5; ++kw)
                for (int c = 0; c < 48; ++c) {
                    const float* src = w_tf + ((size_t)(kh * 5 + kw) * 48 + c) * s.n_total;
                    for (int n = 0; n < s.n_total; ++n) put(n, (size_t)kh * kPackK + kw * 48 + c, src[n]);
                }
    } else {
        for (int t = 0; t < s.taps; ++t)
            for (int c = 0; c < s.cg_real; ++c) {
                const float* src = w_tf + ((size_t)t * s.cg_real + c) * s.n_total;
                for (int n = 0; n < s.n_total; ++n) put(n, (size_t)t * s.cg_pad + c, src[n]);
            }
    }
    int rc;
    if ((rc = dev_alloc(h, &h->w_hi[li], count, false))) return rc;
    if ((rc = dev_alloc(h, &h->w_lo[li], count, false))) return rc;
    if ((rc = dev_alloc(h, &h->bias[li], (size_t)s.n_total, false))) return rc;
    SVX_CUDA_CHECK(cudaMemcpy(h->w_hi[li], hi.data(), count * sizeof(__half), cudaMemcpyHostToDevice));
    SVX_CUDA_CHECK(cudaMemcpy(h->w_lo[li], lo.data(), count * sizeof(__half), cudaMemcpyHostToDevice));
    SVX_CUDA_CHECK(cudaMemcpy(h->bias[li], b_tf, (size_t)s.n_total * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

// Geometry and epilogue of one layer.  `pool` (conv2, conv5): the layer's 3x3/2 max-pool runs in its
// epilogue and the window maxima go to that buffer instead of a full-resolution output.
struct LayerIO {
    const __half* a_hi; const __half* a_lo; long long a_rows; int lda;   // operand planes [a_rows][lda]
    int grid_w, center;                                                  // padded grid width, kernel centre
    float* out_f32; __half* out_hi; __half* out_lo; int ldc;             // full-resolution outputs
    int pos_per_img, valid_h, valid_w;                                   // valid-row mask (0: none)
    float* pool; float* pool2; int pool_h, pool_w;
};

int setup_layer(svx_handle* h, int li, const LayerIO& io) {
    const LayerSpec& s = kSpec[li];
    GemmLayer& L = h->layer[li];
    std::memset(&L, 0, sizeof(L));
    const bool packed = li == L_CONV2;
    const long long K = packed ? (long long)kPackTaps * kPackK : (long long)s.taps * s.cg_pad;
    int rc;
    L.block_n = s.block_n;
    L.chunk_kblocks = s.chunk;
    L.groups = s.groups;
    L.n_per_group = s.n_total / s.groups;
    L.taps = s.taps;
    L.cblocks = s.cg_pad / GEMM_BLOCK_K;
    L.a_group_cols = s.cg_pad;
    L.a_group_rows = 0;
    L.last_ksteps = GEMM_BLOCK_K / 16;
    for (int kh = 0; kh < s.kh; ++kh)
        for (int kw = 0; kw < s.kw; ++kw)
            L.row_off[kh * s.kw + kw] = (kh - io.center) * io.grid_w + (kw - io.center);
    const __half* a_hi = io.a_hi;
    const __half* a_lo = io.a_lo;
    long long a_rows = io.a_rows;
    long long a_cols = io.lda;                    // logical row length seen by the A tensor maps
    if (packed) {
        // one tap per kernel ROW: its 5 x 48 = 240 operand values are contiguous in the packed x2
        // (row stride 48 elements), read as 4 K-blocks through overlapping-row tensor maps; the
        // last block issues 3 of its 4 k-steps (240 = 15 x 16)
        L.taps = kPackTaps;
        L.cblocks = kPackK / GEMM_BLOCK_K;
        L.last_ksteps = 3;
        L.a_group_cols = 0;
        L.a_group_rows = (int)(h->x2_group_elems / 48);
        for (int kh = 0; kh < 5; ++kh) L.row_off[kh] = (kh - 2) * io.grid_w - 2;
        a_cols = kPackK;
        a_rows = kPackLeadRows + 2 * (h->x2_group_elems / 48);
        L.a_row_bias = kPackLeadRows;
        a_hi -= (size_t)kPackLeadRows * 48;       // tensor maps start at the leading zero rows
        a_lo -= (size_t)kPackLeadRows * 48;
    }
    const bool three = h->precision == SVX_PRECISION_3PASS;
    L.use_b_lo = three ? 1 : 0;
    L.use_a_lo = (three && a_lo != nullptr) ? 1 : 0;
    L.m_rows = 0;
    L.bias = h->bias[li];
    L.relu = 1;
    L.out_f32 = io.out_f32;
    L.out_hi = io.out_hi;
    L.out_lo = io.out_lo;
    L.ldc = io.ldc;
    L.pos_per_img = io.pos_per_img;
    L.grid_w = io.grid_w;
    L.valid_h = io.valid_h;
    L.valid_w = io.valid_w;
    L.pool_out = io.pool;
    L.pool_out2 = io.pool2;
    L.pool_h = io.pool_h;
    L.pool_w = io.pool_w;
    if ((rc = plan_layer(L))) return rc;
    if ((rc = make_tensor_map_2d(&L.tm_a_hi, a_hi, a_rows, a_cols, io.lda, L.slab_rows))) return rc;
    if ((rc = make_tensor_map_2d(&L.tm_a_lo, a_lo ? a_lo : a_hi, a_rows, a_cols, io.lda, L.slab_rows))) return rc;
    if ((rc = make_tensor_map_2d(&L.tm_b_hi, h->w_hi[li], s.n_total, K, K, L.block_n / 2))) return rc;
    if ((rc = make_tensor_map_2d(&L.tm_b_lo, h->w_lo[li], s.n_total, K, K, L.block_n / 2))) return rc;
    if (h->dbg) L.dbg = h->dbg + li * 8;
    return 0;
}

int run_layer(svx_handle* h, int li, cudaStream_t st) { return launch_layer(h->layer[li], h->num_sms, st); }

int build_model(svx_handle* h, const svx_weights* w) {
    const long long B = h->max_batch;
    int rc;
    const float* wt[L_COUNT] = {w->conv1_w, w->conv2_w, w->conv3_w, w->conv4_w, w->conv5_w, w->fc6_w, w->fc7_w};
    const float* bt[L_COUNT] = {w->conv1_b, w->conv2_b, w->conv3_b, w->conv4_b, w->conv5_b, w->fc6_b, w->fc7_b};
    for (int li = 0; li < L_COUNT; ++li) {
        if (!wt[li] || !bt[li]) return fail(SVX_ERR_INVALID, "svx_create: NULL weight pointer");
        if ((rc = upload_layer_weights(h, li, wt[li], bt[li]))) return rc;
    }
    if (!w->fc8_w || !w->fc8_b) return fail(SVX_ERR_INVALID, "svx_create: NULL fc8 pointer");
    if ((rc = dev_alloc(h, &h->w8, (size_t)4096 * 5, false))) return rc;
    if ((rc = dev_alloc(h, &h->b8, (size_t)5, false))) return rc;
    SVX_CUDA_CHECK(cudaMemcpy(h->w8, w->fc8_w, 4096 * 5 * sizeof(float), cudaMemcpyHostToDevice));
    SVX_CUDA_CHECK(cudaMemcpy(h->b8, w->fc8_b, 5 * sizeof(float), cudaMemcpyHostToDevice));

    {   // fused front end: 255*W1 and the conv1 response to the all-background image
        const float lo[3] = {-104.f, -117.f, -124.f};               // create_batch.py:13,147-150
        std::vector<float> w255((size_t)11 * 11 * 3 * 96), base(96);
        for (int nn = 0; nn < 96; ++nn) {
            double acc = w->conv1_b[nn];
            for (int k = 0; k < 121; ++k)
                for (int c = 0; c < 3; ++c) acc += (double)lo[c] * (double)w->conv1_w[((size_t)k * 3 + c) * 96 + nn];
            base[nn] = (float)acc;
        }
        for (size_t i = 0; i < w255.size(); ++i) w255[i] = 255.f * w->conv1_w[i];
        if ((rc = dev_alloc(h, &h->front_w255, w255.size(), false))) return rc;
        if ((rc = dev_alloc(h, &h->front_base, base.size(), false))) return rc;
        SVX_CUDA_CHECK(cudaMemcpy(h->front_w255, w255.data(), w255.size() * sizeof(float), cudaMemcpyHostToDevice));
        SVX_CUDA_CHECK(cudaMemcpy(h->front_base, base.data(), base.size() * sizeof(float), cudaMemcpyHostToDevice));
        // 640 scratch slots x 96 floats per CTA of the front kernel (see front.cu SCRATCH_SLOTS)
        h->front_blocks = h->num_sms * 32;
        if (h->front_blocks > B) h->front_blocks = (int)B;
        if ((rc = dev_alloc(h, &h->front_scratch, (size_t)h->front_blocks * 640 * 96, false))) return rc;
    }

    // activations: zero once; pad positions/channels are never written afterwards.
    // x2: two group planes of [B*841 + 64 zero rows][48]; the zero rows are the top padding of the
    // next plane and absorb the 256-wide over-read of the last positions
    h->x2_ld = 48;
    h->x2_group_elems = ((long long)B * P2 + 64) * 48;
    const size_t x2_elems = (size_t)(kPackLeadRows * 48 + 2 * h->x2_group_elems + 256);
    if ((rc = dev_alloc(h, &h->x2_hi, x2_elems))) return rc;
    if ((rc = dev_alloc(h, &h->x2_lo, x2_elems))) return rc;
    h->x2_hi += (size_t)kPackLeadRows * 48;       // data pointers start after the leading zero rows
    h->x2_lo += (size_t)kPackLeadRows * 48;
    if ((rc = dev_alloc(h, &h->p2, (size_t)B * 169 * 256))) return rc;
    if ((rc = dev_alloc(h, &h->p2b, (size_t)B * 169 * 256))) return rc;
    if ((rc = dev_alloc(h, &h->x3_hi, (size_t)B * P3 * 256))) return rc;
    if ((rc = dev_alloc(h, &h->x3_lo, (size_t)B * P3 * 256))) return rc;
    if ((rc = dev_alloc(h, &h->x4_hi, (size_t)B * P3 * 384))) return rc;
    if ((rc = dev_alloc(h, &h->x4_lo, (size_t)B * P3 * 384))) return rc;
    if ((rc = dev_alloc(h, &h->x5_hi, (size_t)B * P3 * 384))) return rc;
    if ((rc = dev_alloc(h, &h->x5_lo, (size_t)B * P3 * 384))) return rc;
    if ((rc = dev_alloc(h, &h->p5, (size_t)B * 36 * 256))) return rc;
    if ((rc = dev_alloc(h, &h->p5b, (size_t)B * 36 * 256))) return rc;
    if ((rc = dev_alloc(h, &h->x6_hi, (size_t)B * 9216))) return rc;
    if ((rc = dev_alloc(h, &h->x6_lo, (size_t)B * 9216))) return rc;
    if ((rc = dev_alloc(h, &h->x7_hi, (size_t)B * 4096))) return rc;
    if ((rc = dev_alloc(h, &h->x7_lo, (size_t)B * 4096))) return rc;
    if ((rc = dev_alloc(h, &h->x8_hi, (size_t)B * 4096))) return rc;
    if ((rc = dev_alloc(h, &h->x8_lo, (size_t)B * 4096))) return rc;
    if (std::getenv("SVX_DBG")) {                 // per-role cycle counters (svx_debug_counters)
        if ((rc = dev_alloc(h, &h->dbg, (size_t)L_COUNT * 8))) return rc;
    }

    //                                   A hi      A lo      A rows  lda       grid c  f32      hi        lo        ldc  pos vh  vw  pool            ph  pw
    if ((rc = setup_layer(h, L_CONV2, {h->x2_hi, h->x2_lo, B * P2, h->x2_ld, G2, 2, nullptr, nullptr, nullptr, 256, P2, 27, 27, h->p2, h->p2b, 13, 13}))) return rc;
    if ((rc = setup_layer(h, L_CONV3, {h->x3_hi, h->x3_lo, B * P3, 256, G3, 1, nullptr, h->x4_hi, h->x4_lo, 384, P3, 13, 13, nullptr, nullptr, 0, 0}))) return rc;
    if ((rc = setup_layer(h, L_CONV4, {h->x4_hi, h->x4_lo, B * P3, 384, G3, 1, nullptr, h->x5_hi, h->x5_lo, 384, P3, 13, 13, nullptr, nullptr, 0, 0}))) return rc;
    if ((rc = setup_layer(h, L_CONV5, {h->x5_hi, h->x5_lo, B * P3, 384, G3, 1, nullptr, nullptr, nullptr, 256, P3, 13, 13, h->p5, h->p5b, 6, 6}))) return rc;
    if ((rc = setup_layer(h, L_FC6, {h->x6_hi, h->x6_lo, B, 9216, 1, 0, nullptr, h->x7_hi, h->x7_lo, 4096, 0, 0, 0, nullptr, nullptr, 0, 0}))) return rc;
    if ((rc = setup_layer(h, L_FC7, {h->x7_hi, h->x7_lo, B, 4096, 1, 0, nullptr, h->x8_hi, h->x8_lo, 4096, 0, 0, 0, nullptr, nullptr, 0, 0}))) return rc;
    h->has_model = true;
    return 0;
}

// The dense conv1 path of svx_forward (arbitrary images): its buffers exist only once it is used.
int ensure_dense_path(svx_handle* h) {
    if (h->x1) return 0;
    const long long D = h->max_batch < kDenseBatch ? h->max_batch : kDenseBatch;
    int rc;
    if ((rc = dev_alloc(h, &h->x1, (size_t)D * P1 * 64))) return rc;
    if ((rc = dev_alloc(h, &h->y1, (size_t)D * P1 * 96))) return rc;
    // conv1's activations (pixel values) are exact in fp16: no lo plane, two passes
    if ((rc = setup_layer(h, L_CONV1, {h->x1, nullptr, D * P1, 64, S2D, 0, h->y1, nullptr, nullptr, 96, 0, 0, 0, nullptr, nullptr, 0, 0}))) return rc;
    SVX_CUDA_CHECK(cudaDeviceSynchronize());
    h->dense_batch = D;
    return 0;
}

// Record "kernel `slot` starts here" (slot -1 closes a sequence) on the launching stream.
void mark(svx_handle* h, int slot, cudaStream_t st) {
    if (!h->profiling) return;
    if (h->ev_used >= (size_t)1 << 16) return;               // bounded; later marks are dropped
    if (h->ev_used == h->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        h->ev_pool.push_back(e);
    }
    cudaEvent_t e = h->ev_pool[h->ev_used++];
    cudaEventRecord(e, st);
    h->ev_marks.emplace_back(slot, e);
}

int profile_collect(svx_handle* h) {
    if (h->ev_marks.empty()) return 0;
    SVX_CUDA_CHECK(cudaEventSynchronize(h->ev_marks.back().second));
    for (size_t i = 0; i + 1 < h->ev_marks.size(); ++i) {
        const int slot = h->ev_marks[i].first;
        if (slot < 0) continue;
        float ms = 0.f;
        SVX_CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev_marks[i].second, h->ev_marks[i + 1].second));
        h->prof_ms[slot] += ms;
        h->prof_launches[slot] += 1;
    }
    h->ev_marks.clear();
    h->ev_used = 0;
    return 0;
}

// conv2 operand (x2) of `n` sites is resident -> labels / probs / logits / calls
int run_cnn(svx_handle* h, long long n, int32_t* labels, float* probs, float* logits,
            cudaStream_t st, const CallSinks* sinks = nullptr) {
    int rc;
    const long long rows[L_COUNT] = {0, n * P2, n * P3, n * P3, n * P3, n, n};
    for (int li = L_CONV2; li < L_COUNT; ++li) h->layer[li].m_rows = rows[li];

    mark(h, 3, st);
    if ((rc = run_layer(h, L_CONV2, st))) return rc;             // + ReLU + pool2 (window maxima -> p2)
    mark(h, 4, st);
    const FinishParams f2{h->p2, h->p2b, P2, G2, 13, 13, 1, h->x3_hi, h->x3_lo, 256, G3, P3};
    if ((rc = launch_finish_pooled(f2, n, h->num_sms, st))) return rc;   // LRN2 + hi/lo split
    mark(h, 5, st);
    if ((rc = run_layer(h, L_CONV3, st))) return rc;
    mark(h, 6, st);
    if ((rc = run_layer(h, L_CONV4, st))) return rc;
    mark(h, 7, st);
    if ((rc = run_layer(h, L_CONV5, st))) return rc;             // + ReLU + pool5 (-> p5)
    mark(h, 8, st);
    const FinishParams f5{h->p5, h->p5b, P3, G3, 6, 6, 0, h->x6_hi, h->x6_lo, 256, 6, 36};   // NHWC flatten (alexnet.py:49)
    if ((rc = launch_finish_pooled(f5, n, h->num_sms, st))) return rc;
    mark(h, 9, st);
    if ((rc = run_layer(h, L_FC6, st))) return rc;
    mark(h, 10, st);
    if ((rc = run_layer(h, L_FC7, st))) return rc;
    mark(h, 11, st);
    static const CallSinks no_sinks = {};
    if ((rc = launch_fc8_softmax(h->x8_hi, h->x8_lo, h->w8, h->b8, n, labels, probs, logits,
                                 sinks ? *sinks : no_sinks, st))) return rc;
    mark(h, -1, st);
    h->last_n = n;
    return 0;
}

// Micro-batch size for a call of n sites: the fewest passes that fit the workspaces, of equal size
// (12 500 sites with room for 10 000 run as 2 x 6 250, not 10 000 + 2 500: the fc layers' tile waves
// and every kernel's tail are paid per pass)
long long balanced_batch(long long n, long long max_batch) {
    if (n <= max_batch) return n > 0 ? n : 1;
    const long long passes = (n + max_batch - 1) / max_batch;
    return (n + passes - 1) / passes;
}

// rows -> conv2 operand: the fused sparse front end (encode + conv1 + ReLU + pool1 + LRN1)
int encode_front(svx_handle* h, const int32_t* rows_dev, long long m, cudaStream_t st) {
    FrontParams fp{h->front_w255, h->front_base, h->x2_hi, h->x2_lo, h->front_scratch, h->front_blocks,
                   h->x2_ld, h->x2_group_elems};
    mark(h, 0, st);
    return launch_front(rows_dev, m, fp, h->num_sms, st);
}

// images -> conv2 operand: space-to-depth, dense tcgen05 conv1, pool1 + LRN1 (svx_forward)
int dense_front(svx_handle* h, const void* images, int dtype, long long m, cudaStream_t st) {
    int rc;
    if ((rc = launch_nhwc_to_s2d(images, dtype, m, h->x1, st))) return rc;
    h->layer[L_CONV1].m_rows = m * P1;
    mark(h, 1, st);
    if ((rc = run_layer(h, L_CONV1, st))) return rc;
    PoolParams p1{};
    p1.in = h->y1; p1.in_grid_w = S2D; p1.in_pos_per_img = P1; p1.C = 96; p1.out_h = 27; p1.out_w = 27;
    p1.lrn = 1; p1.out_hi = h->x2_hi; p1.out_lo = h->x2_lo; p1.out_ld = h->x2_ld; p1.out_grid_w = G2;
    p1.out_pos_per_img = P2; p1.group_real = 48; p1.group_elems = h->x2_group_elems;
    mark(h, 2, st);
    return launch_pool(p1, m, h->num_sms, st);
}

}  // namespace

extern "C" {

const char* svx_last_error(void) { return g_error.c_str(); }
const char* svx_version(void) { return "svx 0.1 (sm_100a; tcgen05+TMA)"; }
int64_t svx_launch_count(void) { return g_launches; }
void svx_launch_count_reset(void) { g_launches = 0; }
int64_t svx_max_batch(const svx_handle* h) { return h ? h->max_batch : 0; }
int svx_device(const svx_handle* h) { return h ? h->device : -1; }

int svx_create(const svx_weights* weights, int device, int64_t max_batch, int precision,
               svx_handle** out) {
    if (!out) return fail(SVX_ERR_INVALID, "svx_create: out is NULL");
    *out = nullptr;
    if (max_batch <= 0 || max_batch > kMaxBatchLimit)
        return fail(SVX_ERR_INVALID, "svx_create: max_batch must be in [1, " + std::to_string(kMaxBatchLimit) +
                                         "] (the workspaces take ~1.4 MB of HBM per site)");
    if (precision != SVX_PRECISION_3PASS && precision != SVX_PRECISION_1PASS)
        return fail(SVX_ERR_INVALID, "svx_create: bad precision");
    int ndev = 0;
    SVX_CUDA_CHECK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(SVX_ERR_INVALID, "svx_create: no such device");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(SVX_ERR_CUDA, "svx_create: cannot select device");
    cudaDeviceProp prop;
    SVX_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(SVX_ERR_UNSUPPORTED, std::string("svx_create: device '") + prop.name +
                                             "' is not sm_100 (this library has no other code path)");
    std::unique_ptr<svx_handle> h(new svx_handle());
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    h->max_batch = max_batch;
    h->precision = precision;
    auto cleanup = [&](int rc) {
        for (void* p : h->allocs) cudaFree(p);
        if (h->busy) cudaEventDestroy(h->busy);
        if (h->stream) cudaStreamDestroy(h->stream);
        return rc;
    };
    int rc;
    SVX_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    SVX_CUDA_CHECK(cudaEventCreateWithFlags(&h->busy, cudaEventDisableTiming));
    if ((rc = dev_alloc(h.get(), &h->rows_dev, (size_t)max_batch * SVX_ROW_FIELDS))) return cleanup(rc);
    if ((rc = dev_alloc(h.get(), &h->labels_dev, (size_t)max_batch))) return cleanup(rc);
    if ((rc = dev_alloc(h.get(), &h->probs_dev, (size_t)max_batch * SVX_NUM_CLASSES))) return cleanup(rc);
    if (weights && (rc = build_model(h.get(), weights))) return cleanup(rc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return cleanup(fail(SVX_ERR_CUDA, std::string("svx_create: ") + cudaGetErrorString(e)));
    *out = h.release();
    return SVX_OK;
}

void svx_destroy(svx_handle* h) {
    if (!h) return;
    DeviceGuard guard(h->device);
    cudaDeviceSynchronize();
    for (void* p : h->allocs) cudaFree(p);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    if (h->busy) cudaEventDestroy(h->busy);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int svx_debug_counters(svx_handle* h, uint64_t* out, int reset) {
    if (!h || !out) return fail(SVX_ERR_INVALID, "svx_debug_counters: bad arguments");
    if (!h->dbg) return fail(SVX_ERR_INVALID, "svx_debug_counters: create the handle with SVX_DBG=1");
    DeviceGuard guard(h->device);
    SVX_CUDA_CHECK(cudaDeviceSynchronize());
    SVX_CUDA_CHECK(cudaMemcpy(out, h->dbg, sizeof(uint64_t) * L_COUNT * 8, cudaMemcpyDeviceToHost));
    if (reset) SVX_CUDA_CHECK(cudaMemset(h->dbg, 0, sizeof(uint64_t) * L_COUNT * 8));
    return SVX_OK;
}

int svx_set_profiling(svx_handle* h, int enable) {
    if (!h) return fail(SVX_ERR_INVALID, "svx_set_profiling: NULL handle");
    DeviceGuard guard(h->device);
    int rc = profile_collect(h);
    h->profiling = enable != 0;
    return rc;
}

int svx_profile_read(svx_handle* h, float* ms_out, int64_t* launches_out, int reset) {
    if (!h) return fail(SVX_ERR_INVALID, "svx_profile_read: NULL handle");
    DeviceGuard guard(h->device);
    int rc = profile_collect(h);
    if (rc) return rc;
    for (int i = 0; i < SVX_PROFILE_SLOTS; ++i) {
        if (ms_out) ms_out[i] = (float)h->prof_ms[i];
        if (launches_out) launches_out[i] = h->prof_launches[i];
        if (reset) { h->prof_ms[i] = 0.0; h->prof_launches[i] = 0; }
    }
    return SVX_OK;
}

int svx_encode(svx_handle* h, const int32_t* rows_dev, int64_t n, void* images_dev, int dtype,
               void* stream) {
    if (!h) return fail(SVX_ERR_INVALID, "svx_encode: NULL handle");
    if (n < 0 || (n > 0 && (!rows_dev || !images_dev))) return fail(SVX_ERR_INVALID, "svx_encode: bad arguments");
    if (dtype != SVX_IMAGE_F32 && dtype != SVX_IMAGE_F16) return fail(SVX_ERR_INVALID, "svx_encode: bad dtype");
    if (reinterpret_cast<uintptr_t>(images_dev) & 15) return fail(SVX_ERR_INVALID, "svx_encode: images_dev must be 16-byte aligned");
    DeviceGuard guard(h->device);
    return launch_encode(rows_dev, n, images_dev, dtype == SVX_IMAGE_F32 ? 0 : 1, h->num_sms,
                         static_cast<cudaStream_t>(stream));
}

int svx_forward(svx_handle* h, const void* images_dev, int dtype, int64_t n, float* logits_dev,
                void* stream) {
    if (!h || !h->has_model) return fail(SVX_ERR_INVALID, "svx_forward: handle has no model");
    if (n < 0 || (n > 0 && (!images_dev || !logits_dev))) return fail(SVX_ERR_INVALID, "svx_forward: bad arguments");
    if (dtype != SVX_IMAGE_F32 && dtype != SVX_IMAGE_F16) return fail(SVX_ERR_INVALID, "svx_forward: bad dtype");
    DeviceGuard guard(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t esz = dtype == SVX_IMAGE_F32 ? 4 : 2;
    int rc;
    if ((rc = ensure_dense_path(h))) return rc;
    SVX_CUDA_CHECK(cudaStreamWaitEvent(st, h->busy, 0));        // workspaces free of earlier entries
    for (int64_t s = 0; s < n; s += h->dense_batch) {
        const int64_t m = n - s < h->dense_batch ? n - s : h->dense_batch;
        const char* src = static_cast<const char*>(images_dev) + (size_t)s * SVX_IMG * SVX_IMG * 3 * esz;
        if ((rc = dense_front(h, src, dtype, m, st))) return rc;
        if ((rc = run_cnn(h, m, h->labels_dev, h->probs_dev, logits_dev + s * SVX_NUM_CLASSES, st))) return rc;
    }
    SVX_CUDA_CHECK(cudaEventRecord(h->busy, st));
    return SVX_OK;
}

int svx_classify_device(svx_handle* h, const int32_t* rows_dev, int64_t n, int32_t* labels_dev,
                        float* probs_dev, float* logits_dev, void* stream) {
    if (!h || !h->has_model) return fail(SVX_ERR_INVALID, "svx_classify_device: handle has no model");
    if (n < 0 || (n > 0 && (!rows_dev || !labels_dev || !probs_dev)))
        return fail(SVX_ERR_INVALID, "svx_classify_device: bad arguments");
    DeviceGuard guard(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SVX_CUDA_CHECK(cudaStreamWaitEvent(st, h->busy, 0));
    const int64_t mb = balanced_batch(n, h->max_batch);
    for (int64_t s = 0; s < n; s += mb) {
        const int64_t m = n - s < mb ? n - s : mb;
        int rc;
        if ((rc = encode_front(h, rows_dev + s * SVX_ROW_FIELDS, m, st))) return rc;
        if ((rc = run_cnn(h, m, labels_dev + s, probs_dev + s * SVX_NUM_CLASSES,
                          logits_dev ? logits_dev + s * SVX_NUM_CLASSES : nullptr, st)))
            return rc;
    }
    SVX_CUDA_CHECK(cudaEventRecord(h->busy, st));
    return SVX_OK;
}

int svx_classify(svx_handle* h, const int32_t* rows_host, int64_t n, int32_t* labels_host,
                 float* probs_host) {
    if (!h || !h->has_model) return fail(SVX_ERR_INVALID, "svx_classify: handle has no model");
    if (n < 0 || (n > 0 && (!rows_host || !labels_host || !probs_host)))
        return fail(SVX_ERR_INVALID, "svx_classify: bad arguments");
    DeviceGuard guard(h->device);
    cudaStream_t st = h->stream;
    SVX_CUDA_CHECK(cudaStreamWaitEvent(st, h->busy, 0));
    const int64_t mb = balanced_batch(n, h->max_batch);
    for (int64_t s = 0; s < n; s += mb) {
        const int64_t m = n - s < mb ? n - s : mb;
        int rc;
        SVX_CUDA_CHECK(cudaMemcpyAsync(h->rows_dev, rows_host + s * SVX_ROW_FIELDS,
                                       (size_t)m * SVX_ROW_FIELDS * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        if ((rc = encode_front(h, h->rows_dev, m, st))) return rc;
        if ((rc = run_cnn(h, m, h->labels_dev, h->probs_dev, nullptr, st))) return rc;
        SVX_CUDA_CHECK(cudaMemcpyAsync(labels_host + s, h->labels_dev, (size_t)m * sizeof(int32_t),
                                       cudaMemcpyDeviceToHost, st));
        SVX_CUDA_CHECK(cudaMemcpyAsync(probs_host + s * SVX_NUM_CLASSES, h->probs_dev,
                                       (size_t)m * SVX_NUM_CLASSES * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    SVX_CUDA_CHECK(cudaEventRecord(h->busy, st));
    SVX_CUDA_CHECK(cudaStreamSynchronize(st));
    return SVX_OK;
}

int svx_classify_device_calls(svx_handle* h, const int32_t* rows_dev, int64_t n, svx_call* calls_dev,
                              void* stream) {
    if (!h || !h->has_model) return fail(SVX_ERR_INVALID, "svx_classify_device_calls: handle has no model");
    if (n < 0 || (n > 0 && (!rows_dev || !calls_dev)))
        return fail(SVX_ERR_INVALID, "svx_classify_device_calls: bad arguments");
    DeviceGuard guard(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SVX_CUDA_CHECK(cudaStreamWaitEvent(st, h->busy, 0));
    const int64_t mb = balanced_batch(n, h->max_batch);
    for (int64_t s = 0; s < n; s += mb) {
        const int64_t m = n - s < mb ? n - s : mb;
        int rc;
        CallSinks sinks = {};
        sinks.count = 1;
        sinks.ptr[0] = reinterpret_cast<int2*>(calls_dev + s);
        if ((rc = encode_front(h, rows_dev + s * SVX_ROW_FIELDS, m, st))) return rc;
        if ((rc = run_cnn(h, m, nullptr, nullptr, nullptr, st, &sinks))) return rc;
    }
    SVX_CUDA_CHECK(cudaEventRecord(h->busy, st));
    return SVX_OK;
}

int svx_exchange_create(svx_handle* h, int rank, int world, int64_t sites_per_rank, svx_exchange** out) {
    if (!out) return fail(SVX_ERR_INVALID, "svx_exchange_create: out is NULL");
    *out = nullptr;
    if (!h || !h->has_model) return fail(SVX_ERR_INVALID, "svx_exchange_create: handle has no model");
    if (world < 1 || world > CALL_MAX_SINKS || rank < 0 || rank >= world || sites_per_rank <= 0)
        return fail(SVX_ERR_INVALID, "svx_exchange_create: bad rank / world (<= 16) / sites_per_rank");
    DeviceGuard guard(h->device);
    std::unique_ptr<svx_exchange> x(new svx_exchange());
    x->h = h; x->device = h->device; x->rank = rank; x->world = world; x->per_rank = sites_per_rank;
    if (const char* e = std::getenv("SVX_EXCHANGE_TIMEOUT_MS")) {
        const long long ms = std::atoll(e);
        if (ms > 0) x->timeout_ns = (unsigned long long)ms * 1000000ull;
    }
    const size_t bytes = 256 + 2 * x->region_bytes();
    // plain cudaMalloc (not a pool): the allocation is exported with cudaIpcGetMemHandle
    cudaError_t e = cudaMalloc(&x->local, bytes);
    if (e != cudaSuccess) return fail(SVX_ERR_NOMEM, std::string("svx_exchange_create: ") + cudaGetErrorString(e));
    auto cleanup = [&](int rc) {
        cudaFree(x->local); cudaFree(x->done);
        if (x->error_host) cudaFreeHost(x->error_host);
        return rc;
    };
    if ((e = cudaMemset(x->local, 0, bytes)) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void**>(&x->done), sizeof(unsigned int))) != cudaSuccess ||
        (e = cudaMemset(x->done, 0, sizeof(unsigned int))) != cudaSuccess ||
        (e = cudaHostAlloc(reinterpret_cast<void**>(&x->error_host), sizeof(unsigned int), cudaHostAllocMapped)) != cudaSuccess ||
        (e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&x->error), x->error_host, 0)) != cudaSuccess ||
        (e = cudaDeviceSynchronize()) != cudaSuccess)
        return cleanup(fail(SVX_ERR_CUDA, std::string("svx_exchange_create: ") + cudaGetErrorString(e)));
    *x->error_host = 0;
    x->base[rank] = x->local;
    x->attached = world == 1;
    *out = x.release();
    return SVX_OK;
}

// A CUDA IPC handle names the whole driver allocation a pointer lives in (cudaMalloc sub-allocates
// small requests from larger blocks), and opening it yields that allocation's BASE: the exported blob
// therefore carries the buffer's offset from the base as well.
int svx_exchange_export(svx_exchange* x, void* ipc_handle_out) {
    if (!x || !ipc_handle_out) return fail(SVX_ERR_INVALID, "svx_exchange_export: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) + sizeof(uint64_t) == SVX_IPC_HANDLE_BYTES, "IPC blob size");
    DeviceGuard guard(x->device);
    cudaIpcMemHandle_t hd;
    SVX_CUDA_CHECK(cudaIpcGetMemHandle(&hd, x->local));
    typedef CUresult (*PFN_range)(CUdeviceptr*, size_t*, CUdeviceptr);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !fn)
        return fail(SVX_ERR_CUDA, "svx_exchange_export: cuMemGetAddressRange entry point not available");
    CUdeviceptr base = 0;
    size_t size = 0;
    if (reinterpret_cast<PFN_range>(fn)(&base, &size, reinterpret_cast<CUdeviceptr>(x->local)) != CUDA_SUCCESS)
        return fail(SVX_ERR_CUDA, "svx_exchange_export: cuMemGetAddressRange failed");
    const uint64_t offset = reinterpret_cast<CUdeviceptr>(x->local) - base;
    std::memcpy(ipc_handle_out, &hd, sizeof(hd));
    std::memcpy(static_cast<char*>(ipc_handle_out) + sizeof(hd), &offset, sizeof(offset));
    return SVX_OK;
}

int svx_exchange_attach(svx_exchange* x, const void* ipc_handles) {
    if (!x || !ipc_handles) return fail(SVX_ERR_INVALID, "svx_exchange_attach: bad arguments");
    if (x->attached) return fail(SVX_ERR_INVALID, "svx_exchange_attach: already attached");
    DeviceGuard guard(x->device);
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t hd;
        uint64_t offset = 0;
        const char* blob = static_cast<const char*>(ipc_handles) + (size_t)r * SVX_IPC_HANDLE_BYTES;
        std::memcpy(&hd, blob, sizeof(hd));
        std::memcpy(&offset, blob + sizeof(hd), sizeof(offset));
        cudaError_t e = cudaIpcOpenMemHandle(&x->mapped[r], hd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return fail(SVX_ERR_CUDA, "svx_exchange_attach: cudaIpcOpenMemHandle(rank " + std::to_string(r) +
                                          "): " + cudaGetErrorString(e));
        x->base[r] = static_cast<char*>(x->mapped[r]) + offset;
        x->opened[r] = true;
    }
    x->attached = true;
    return SVX_OK;
}

int svx_classify_exchange(svx_handle* h, svx_exchange* x, const int32_t* rows_dev, int64_t n,
                          const svx_call** gathered_dev, void* stream) {
    if (!h || !h->has_model || !x || x->h != h) return fail(SVX_ERR_INVALID, "svx_classify_exchange: bad handle");
    if (!x->attached) return fail(SVX_ERR_INVALID, "svx_classify_exchange: exchange is not attached");
    if (n <= 0 || n > x->per_rank || !rows_dev)
        return fail(SVX_ERR_INVALID, "svx_classify_exchange: need 0 < n <= sites_per_rank (pad the shard)");
    // a timeout of an earlier call is sticky: the gathered buffers of that call (and possibly of this
    // one) hold poisoned calls (label -1) for the rank that did not show up
    if (const unsigned int err = *static_cast<volatile unsigned int*>(x->error_host))
        return fail(SVX_ERR_CUDA, "svx_classify_exchange: an earlier exchange timed out waiting for rank " +
                                      std::to_string(err - 1) + " (svx_exchange_status clears the error)");
    DeviceGuard guard(h->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SVX_CUDA_CHECK(cudaStreamWaitEvent(st, h->busy, 0));
    const unsigned long long epoch = ++x->epoch;
    const int parity = (int)(epoch & 1ull);
    const int64_t mb = balanced_batch(n, h->max_batch);
    for (int64_t s = 0; s < n; s += mb) {
        const int64_t m = n - s < mb ? n - s : mb;
        int rc;
        CallSinks sinks = {};
        sinks.count = x->world;
        for (int r = 0; r < x->world; ++r) {
            sinks.ptr[r] = reinterpret_cast<int2*>(x->region(r, parity)) + (size_t)x->rank * x->per_rank + s;
            sinks.flag[r] = x->flags(r) + x->rank;
        }
        if (s + m == n) {                      // last micro-batch: its fc8 kernel publishes the epoch
            sinks.epoch = epoch;
            sinks.done = x->done;
        }
        if ((rc = encode_front(h, rows_dev + s * SVX_ROW_FIELDS, m, st))) return rc;
        if ((rc = run_cnn(h, m, nullptr, nullptr, nullptr, st, &sinks))) return rc;
    }
    int rc;
    if ((rc = launch_exchange_wait(x->flags(x->rank), x->world, epoch, x->timeout_ns, x->error,
                                   reinterpret_cast<int2*>(x->region(x->rank, parity)), x->per_rank, st))) return rc;
    SVX_CUDA_CHECK(cudaEventRecord(h->busy, st));
    if (gathered_dev) *gathered_dev = reinterpret_cast<const svx_call*>(x->region(x->rank, parity));
    return SVX_OK;
}

int svx_exchange_status(svx_exchange* x) {
    if (!x) return fail(SVX_ERR_INVALID, "svx_exchange_status: NULL exchange");
    DeviceGuard guard(x->device);
    SVX_CUDA_CHECK(cudaDeviceSynchronize());
    const unsigned int err = *static_cast<volatile unsigned int*>(x->error_host);
    if (err != 0) {
        *x->error_host = 0;
        return fail(SVX_ERR_CUDA, "svx_classify_exchange: timed out waiting for rank " + std::to_string(err - 1) +
                                      "; its calls in the gathered buffer are poisoned (label -1)");
    }
    return SVX_OK;
}

void svx_exchange_destroy(svx_exchange* x) {
    if (!x) return;
    DeviceGuard guard(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; ++r)
        if (x->opened[r]) cudaIpcCloseMemHandle(x->mapped[r]);
    cudaFree(x->local);
    cudaFree(x->done);
    if (x->error_host) cudaFreeHost(x->error_host);
    delete x;
}

int svx_debug_activation(svx_handle* h, const char* name, int64_t n, float* out_host) {
    if (!h || !h->has_model || !name || !out_host) return fail(SVX_ERR_INVALID, "svx_debug_activation: bad arguments");
    if (n <= 0 || n > h->last_n) return fail(SVX_ERR_INVALID, "svx_debug_activation: n exceeds the last micro-batch");
    DeviceGuard guard(h->device);
    SVX_CUDA_CHECK(cudaDeviceSynchronize());
    struct Src { const char* name; const float* f32; const __half* hi; const __half* lo; int pos, grid_w, H, W, ld, C, greal; long long gelems; };
    const Src table[] = {
        {"conv1", h->y1, nullptr, nullptr, P1, S2D, 55, 55, 96, 96, 96, 96},      // svx_forward only
        {"norm1", nullptr, h->x2_hi, h->x2_lo, P2, G2, 27, 27, h->x2_ld, 96, 48, h->x2_group_elems},
        {"norm2", nullptr, h->x3_hi, h->x3_lo, P3, G3, 13, 13, 256, 256, 256, 256},
        {"conv3", nullptr, h->x4_hi, h->x4_lo, P3, G3, 13, 13, 384, 384, 384, 384},
        {"conv4", nullptr, h->x5_hi, h->x5_lo, P3, G3, 13, 13, 384, 384, 384, 384},
        {"pool5", nullptr, h->x6_hi, h->x6_lo, 36, 6, 6, 6, 256, 256, 256, 256},
        {"fc6", nullptr, h->x7_hi, h->x7_lo, 1, 1, 1, 1, 4096, 4096, 4096, 4096},
        {"fc7", nullptr, h->x8_hi, h->x8_lo, 1, 1, 1, 1, 4096, 4096, 4096, 4096},
    };
    for (const Src& s : table) {
        if (std::strcmp(s.name, name) != 0) continue;
        if (!s.f32 && !s.hi) return fail(SVX_ERR_INVALID, "svx_debug_activation: 'conv1' exists only after svx_forward");
        if (s.f32 && n > h->dense_batch) return fail(SVX_ERR_INVALID, "svx_debug_activation: n exceeds the dense-path batch");
        const int ngroups = s.C / s.greal;
        const size_t count = (size_t)(ngroups - 1) * (size_t)s.gelems + ((size_t)n * s.pos - 1) * s.ld + s.greal;
        std::vector<float> buf(count);
        if (s.f32) {
            SVX_CUDA_CHECK(cudaMemcpy(buf.data(), s.f32, count * sizeof(float), cudaMemcpyDeviceToHost));
        } else {
            std::vector<__half> hi(count), lo(count);
            SVX_CUDA_CHECK(cudaMemcpy(hi.data(), s.hi, count * sizeof(__half), cudaMemcpyDeviceToHost));
            SVX_CUDA_CHECK(cudaMemcpy(lo.data(), s.lo, count * sizeof(__half), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < count; ++i) buf[i] = __half2float(hi[i]) + __half2float(lo[i]);
        }
        size_t o = 0;
        for (int64_t im = 0; im < n; ++im)
            for (int y = 0; y < s.H; ++y)
                for (int x = 0; x < s.W; ++x)
                    for (int c = 0; c < s.C; ++c) {
                        out_host[o++] = buf[(size_t)(c / s.greal) * (size_t)s.gelems +
                                            ((size_t)im * s.pos + (size_t)y * s.grid_w + x) * s.ld + (c % s.greal)];
                    }
        return SVX_OK;
    }
    return fail(SVX_ERR_INVALID, std::string("svx_debug_activation: unknown activation '") + name + "'");
}

static thread_local float g_selftest_ms = 0.f;

// Shared by the two self-test entries: fp32 operands on the device -> hi/lo planes -> one launch of
// the layer kernel with `taps` row offsets -> fp32 result.
static int layer_selftest(int device, const float* a_dev, const float* b_dev, float* c_dev, int64_t m,
                          int64_t n, int64_t k_per_tap, int taps, const int* row_off, int block_n,
                          int precision, cudaStream_t st) {
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    SVX_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(SVX_ERR_UNSUPPORTED, "selftest: device is not sm_100");
    const int64_t k = k_per_tap * taps;
    __half *a_hi = nullptr, *a_lo = nullptr, *b_hi = nullptr, *b_lo = nullptr;
    float* bias = nullptr;
    auto free_all = [&](int rc) {
        cudaStreamSynchronize(st);
        cudaFree(a_hi); cudaFree(a_lo); cudaFree(b_hi); cudaFree(b_lo); cudaFree(bias);
        return rc;
    };
    SVX_CUDA_CHECK(cudaMalloc(&a_hi, (size_t)m * k_per_tap * 2));
    SVX_CUDA_CHECK(cudaMalloc(&a_lo, (size_t)m * k_per_tap * 2));
    SVX_CUDA_CHECK(cudaMalloc(&b_hi, (size_t)n * k * 2));
    SVX_CUDA_CHECK(cudaMalloc(&b_lo, (size_t)n * k * 2));
    SVX_CUDA_CHECK(cudaMalloc(&bias, (size_t)n * 4));
    SVX_CUDA_CHECK(cudaMemsetAsync(bias, 0, (size_t)n * 4, st));
    int rc;
    if ((rc = launch_split_hilo(a_dev, m * k_per_tap, a_hi, a_lo, st))) return free_all(rc);
    if ((rc = launch_split_hilo(b_dev, n * k, b_hi, b_lo, st))) return free_all(rc);
    GemmLayer L;
    std::memset(&L, 0, sizeof(L));
    L.block_n = block_n; L.chunk_kblocks = kChunkKBlocks; L.groups = 1; L.n_per_group = (int)n;
    L.taps = taps; L.cblocks = (int)(k_per_tap / GEMM_BLOCK_K); L.a_group_cols = 0;
    L.last_ksteps = GEMM_BLOCK_K / 16;
    for (int t = 0; t < taps; ++t) L.row_off[t] = row_off[t];
    L.use_a_lo = L.use_b_lo = precision == SVX_PRECISION_3PASS ? 1 : 0;
    L.m_rows = m; L.bias = bias; L.relu = 0; L.out_f32 = c_dev; L.ldc = (int)n;
    if ((rc = plan_layer(L))) return free_all(rc);
    if ((rc = make_tensor_map_2d(&L.tm_a_hi, a_hi, m, k_per_tap, k_per_tap, L.slab_rows))) return free_all(rc);
    if ((rc = make_tensor_map_2d(&L.tm_a_lo, a_lo, m, k_per_tap, k_per_tap, L.slab_rows))) return free_all(rc);
    if ((rc = make_tensor_map_2d(&L.tm_b_hi, b_hi, n, k, k, block_n / 2))) return free_all(rc);
    if ((rc = make_tensor_map_2d(&L.tm_b_lo, b_lo, n, k, k, block_n / 2))) return free_all(rc);
    // the layer kernel alone, timed on its stream (svx_selftest_last_ms)
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    rc = launch_layer(L, prop.multiProcessorCount, st);
    cudaEventRecord(e1, st);
    if (rc == 0 && cudaEventSynchronize(e1) == cudaSuccess) cudaEventElapsedTime(&g_selftest_ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return free_all(rc);
}

float svx_selftest_last_ms(void) { return g_selftest_ms; }

int svx_gemm_selftest(int device, const float* a_dev, const float* b_dev, float* c_dev, int64_t m,
                      int64_t n, int64_t k, int block_n, int precision, void* stream) {
    if (!a_dev || !b_dev || !c_dev || m <= 0 || n <= 0 || k <= 0)
        return fail(SVX_ERR_INVALID, "svx_gemm_selftest: bad arguments");
    if (k % GEMM_BLOCK_K != 0 || block_n <= 0 || n % block_n != 0)
        return fail(SVX_ERR_INVALID, "svx_gemm_selftest: k must be a multiple of 64 and n of block_n");
    const int zero = 0;
    return layer_selftest(device, a_dev, b_dev, c_dev, m, n, k, 1, &zero, block_n, precision,
                          static_cast<cudaStream_t>(stream));
}

int svx_conv_selftest(int device, const float* a_dev, const float* b_dev, float* c_dev, int64_t m,
                      int64_t n, int64_t k_per_tap, int taps, const int* row_off, int block_n,
                      int precision, void* stream) {
    if (!a_dev || !b_dev || !c_dev || !row_off || m <= 0 || n <= 0 || k_per_tap <= 0 || taps < 1 ||
        taps > GEMM_MAX_TAPS)
        return fail(SVX_ERR_INVALID, "svx_conv_selftest: bad arguments");
    if (k_per_tap % GEMM_BLOCK_K != 0 || block_n <= 0 || n % block_n != 0)
        return fail(SVX_ERR_INVALID, "svx_conv_selftest: k_per_tap % 64 or n % block_n != 0");
    return layer_selftest(device, a_dev, b_dev, c_dev, m, n, k_per_tap, taps, row_off, block_n, precision,
                          static_cast<cudaStream_t>(stream));
}

}  // extern "C"
