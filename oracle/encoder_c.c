/*
 * ORACLE (test infrastructure, not product code) -- plain-C restatement of SVision's
 * segment-pair image encoder, used as the checker for the CUDA encoder at sizes where the
 * Python restatement (oracle/encoder.py) is too slow, and as the "port" CPU baseline that
 * bench.py times.  Nothing in svision_b200/ links or loads this file.
 *
 * Parity status: PINNED.  tests/test_oracle_encoder.py checks this file bit-for-bit against
 * tests/golden/encoder_golden.npz, which holds outputs of the reference's own
 * BatchGenerator.next_batch (see oracle/make_golden.py).
 *
 * Follows (paths relative to the reference root):
 *   src/network/create_batch.py:103-140   row -> two Segments (length = y_end - y_start)
 *   src/segmentplot/classes.py:44-54       xEnd = xStart +/- (L-1), yEnd = yStart + L - 1
 *   src/segmentplot/plot_segment.py:12-15  ratio = max(len)/227.0, clamped to >= 1 (double)
 *   src/segmentplot/plot_segment.py:43-52  end points int(v/ratio); reverse drawn end->start
 *   OpenCV cv::line (clipLine + LineIterator 8-connected, leftToRight) -- third-party, pinned
 *     to opencv-python-headless 4.13.0 through the golden vectors
 *   src/segmentplot/plot_segment.py:57-65  channel 1 = channel 0 on columns with >= 2 pixels
 *   src/network/create_batch.py:146-150    float32, minus [104,117,124]
 *
 * Build: see oracle/Makefile (gcc -O2 -pthread -shared; no -ffast-math: the scaling must be
 * IEEE double division followed by truncation).
 */
#include <stdint.h>
#include <string.h>

#define IMG 227

typedef struct { int64_t x, y; } pt_t;

static int clip_line(pt_t *p1, pt_t *p2)
{
    const int64_t right = IMG - 1, bottom = IMG - 1;
    int64_t x1 = p1->x, y1 = p1->y, x2 = p2->x, y2 = p2->y;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        int64_t a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (int64_t)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (int64_t)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (int64_t)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (int64_t)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    p1->x = x1; p1->y = y1; p2->x = x2; p2->y = y2;
    return (c1 | c2) == 0;
}

/* bits: uint8[3][227][227], 0/1 */
static void draw_line(uint8_t *bits, pt_t p1, pt_t p2, int reverse)
{
    if (p1.x < 0 || p1.x >= IMG || p1.y < 0 || p1.y >= IMG ||
        p2.x < 0 || p2.x >= IMG || p2.y < 0 || p2.y >= IMG) {
        if (!clip_line(&p1, &p2)) return;
    }
    int x1 = (int)p1.x, y1 = (int)p1.y;
    int dx = (int)(p2.x - p1.x), dy = (int)(p2.y - p1.y), sy = 1;
    if (dx < 0) { dx = -dx; dy = -dy; x1 = (int)p2.x; y1 = (int)p2.y; }
    if (dy < 0) { dy = -dy; sy = -1; }
    int vert = dy > dx;
    if (vert) { int t = dx; dx = dy; dy = t; }
    for (int i = 0; i <= dx; ++i) {
        int s = dx > 0 ? (2 * dy * i + dx - 1) / (2 * dx) : 0;
        int c = vert ? x1 + s : x1 + i;
        int r = vert ? y1 + sy * i : y1 + sy * s;
        bits[r * IMG + c] = 1;
        if (reverse) bits[2 * IMG * IMG + r * IMG + c] = 1;
    }
}

void svo_encode_bits_one(const int32_t *row, uint8_t *bits)
{
    memset(bits, 0, 3 * IMG * IMG);
    int32_t la = row[10], lb = row[11];
    double ratio = (double)(la > lb ? la : lb) / 227.0;
    if (ratio < 1.0) ratio = 1.0;
    for (int s = 0; s < 2; ++s) {
        const int32_t *g = row + 5 * s;
        int64_t xs = g[0], ys = g[2], ye = g[3];
        int fwd = g[4] == 1;
        int64_t len = ye - ys;
        int64_t xe = fwd ? xs + (len - 1) : xs - (len - 1);
        int64_t ye2 = ys + (len - 1);
        pt_t ps = { (int64_t)((double)ys / ratio), (int64_t)((double)xs / ratio) };
        pt_t pe = { (int64_t)((double)ye2 / ratio), (int64_t)((double)xe / ratio) };
        if (fwd) draw_line(bits, ps, pe, 0);
        else     draw_line(bits, pe, ps, 1);
    }
    uint8_t *c0 = bits, *c1 = bits + IMG * IMG;
    for (int c = 0; c < IMG; ++c) {
        int cnt = 0;
        for (int r = 0; r < IMG; ++r) cnt += c0[r * IMG + c];
        if (cnt >= 2)
            for (int r = 0; r < IMG; ++r) c1[r * IMG + c] = c0[r * IMG + c];
    }
}

/* ---- tiny pthread parallel-for (no libgomp in the image) ---- */
#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

typedef void (*range_fn)(const int32_t *rows, int64_t lo, int64_t hi, void *out);
typedef struct { range_fn fn; const int32_t *rows; int64_t lo, hi; void *out; } job_t;

static void *job_main(void *p) { job_t *j = (job_t *)p; j->fn(j->rows, j->lo, j->hi, j->out); return 0; }

static int g_threads = 0;
void svo_set_threads(int t) { g_threads = t; }
int svo_get_threads(void)
{
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static void parallel_for(range_fn fn, const int32_t *rows, int64_t n, void *out)
{
    int t = svo_get_threads();
    if (t > n) t = (int)(n > 0 ? n : 1);
    if (t <= 1) { fn(rows, 0, n, out); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * t);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * t);
    int64_t per = (n + t - 1) / t;
    for (int i = 0; i < t; ++i) {
        int64_t lo = i * per, hi = lo + per > n ? n : lo + per;
        if (lo > n) lo = n;
        jobs[i] = (job_t){ fn, rows, lo, hi, out };
        pthread_create(&th[i], 0, job_main, &jobs[i]);
    }
    for (int i = 0; i < t; ++i) pthread_join(th[i], 0);
    free(th); free(jobs);
}

static void range_bits(const int32_t *rows, int64_t lo, int64_t hi, void *out)
{
    uint8_t *bits = (uint8_t *)out;
    for (int64_t i = lo; i < hi; ++i)
        svo_encode_bits_one(rows + 12 * i, bits + (size_t)i * 3 * IMG * IMG);
}

static void range_f32(const int32_t *rows, int64_t lo_i, int64_t hi_i, void *out)
{
    static const float lo[3] = { -104.f, -117.f, -124.f };
    static const float hi[3] = { 151.f, 138.f, 131.f };
    uint8_t *bits = (uint8_t *)malloc(3 * IMG * IMG);
    for (int64_t i = lo_i; i < hi_i; ++i) {
        svo_encode_bits_one(rows + 12 * i, bits);
        float *o = (float *)out + (size_t)i * IMG * IMG * 3;
        for (int p = 0; p < IMG * IMG; ++p)
            for (int c = 0; c < 3; ++c)
                o[p * 3 + c] = bits[c * IMG * IMG + p] ? hi[c] : lo[c];
    }
    free(bits);
}

/* 64-bit FNV-1a digest per image over the 3-bit pixel codes; lets full-size runs be compared
 * without materialising every image on both sides. */
static void range_digest(const int32_t *rows, int64_t lo, int64_t hi, void *out)
{
    uint8_t *bits = (uint8_t *)malloc(3 * IMG * IMG);
    for (int64_t i = lo; i < hi; ++i) {
        svo_encode_bits_one(rows + 12 * i, bits);
        uint64_t h = 1469598103934665603ULL;
        for (int p = 0; p < IMG * IMG; ++p) {
            uint8_t v = (uint8_t)(bits[p] | (bits[IMG * IMG + p] << 1) | (bits[2 * IMG * IMG + p] << 2));
            h = (h ^ v) * 1099511628211ULL;
        }
        ((uint64_t *)out)[i] = h;
    }
    free(bits);
}

/* rows int32[n][12] -> bits uint8[n][3][227][227] */
void svo_encode_bits(const int32_t *rows, int64_t n, uint8_t *bits) { parallel_for(range_bits, rows, n, bits); }
/* rows int32[n][12] -> NHWC float32[n][227][227][3], mean-subtracted as the reference yields */
void svo_encode_f32(const int32_t *rows, int64_t n, float *out) { parallel_for(range_f32, rows, n, out); }
/* rows int32[n][12] -> uint64[n] digests */
void svo_encode_digest(const int32_t *rows, int64_t n, uint64_t *digest) { parallel_for(range_digest, rows, n, digest); }
