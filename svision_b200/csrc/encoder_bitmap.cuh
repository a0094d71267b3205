// Shared by encoder.cu and front.cu: builds the three 227x256-bit lit-pixel planes of one site in
// shared memory, bit-exact with the reference rasteriser (see encoder.cu for the citations).
#pragma once

#include "common.cuh"

namespace svx {
namespace bitmap {

constexpr int IMG = 227;
constexpr int NPIX = IMG * IMG;          // 51529
constexpr int NEL = NPIX * 3;            // 154587 elements per image
constexpr int BMW = 8;                   // 32-bit words per bitmap row (256 >= 227 columns)
constexpr int BMROWS = 228;              // +1 all-zero row (space-to-depth pad row / straddle reads)
constexpr int PLANE = BMROWS * BMW;      // words per channel plane
constexpr int FRONT_THREADS = 256;       // fused front end: more warps per site for the lit-pixel work

struct LineParams {
    int x1, y1, dx, dy, sy, vert, count, rev;
};

static __device__ __forceinline__ bool clip_line(long long& x1, long long& y1, long long& x2,
                                          long long& y2) {
    const long long right = IMG - 1, bottom = IMG - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// Segment s of a packed row -> draw parameters (count == 0 when the line is rejected).
static __device__ __forceinline__ LineParams setup_line(const int32_t* __restrict__ row, int s) {
    const int la = row[10], lb = row[11];
    double ratio = (double)(la > lb ? la : lb) / 227.0;          // plot_segment.py:12
    if (ratio < 1.0) ratio = 1.0;                                // plot_segment.py:14-15
    const int32_t* g = row + 5 * s;
    const long long xs = g[0], ys = g[2], ye = g[3];
    const bool fwd = g[4] == 1;
    const long long len = ye - ys;                               // create_batch.py:118,132
    const long long xe = fwd ? xs + (len - 1) : xs - (len - 1);  // segmentplot/classes.py:50-53
    const long long ye2 = ys + (len - 1);                        // segmentplot/classes.py:54
    // (col, row) = (ref, read); int(v / ratio): fp64 division, truncation toward zero
    long long sx = (long long)((double)ys / ratio), sy_ = (long long)((double)xs / ratio);
    long long ex = (long long)((double)ye2 / ratio), ey = (long long)((double)xe / ratio);
    long long x1, y1, x2, y2;
    if (fwd) { x1 = sx; y1 = sy_; x2 = ex; y2 = ey; }            // plot_segment.py:46-47
    else     { x1 = ex; y1 = ey; x2 = sx; y2 = sy_; }            // plot_segment.py:49-52
    LineParams L;
    L.rev = fwd ? 0 : 1;
    L.count = 0;
    L.x1 = L.y1 = L.dx = L.dy = L.vert = 0;
    L.sy = 1;
    if (x1 < 0 || x1 >= IMG || y1 < 0 || y1 >= IMG || x2 < 0 || x2 >= IMG || y2 < 0 || y2 >= IMG) {
        if (!clip_line(x1, y1, x2, y2)) return L;
    }
    int dx = (int)(x2 - x1), dy = (int)(y2 - y1);
    int px = (int)x1, py = (int)y1;
    if (dx < 0) { dx = -dx; dy = -dy; px = (int)x2; py = (int)y2; }   // leftToRight
    if (dy < 0) { dy = -dy; L.sy = -1; }
    L.vert = dy > dx;
    if (L.vert) { int t = dx; dx = dy; dy = t; }
    L.x1 = px; L.y1 = py; L.dx = dx; L.dy = dy;
    L.count = dx + 1;
    return L;
}

// Builds the three bit planes of one image in shared memory.  All threads of the CTA call it.
template <int NT>
static __device__ __forceinline__ void build_bitmap(const int32_t* __restrict__ row, uint32_t* bm, LineParams* lines,
                             uint32_t* red /* [NT/32 warps][8 words][2] */, uint32_t* colmask) {
    const int tid = threadIdx.x;
    // zero the planes (3*228*8 words = 1368 uint4)
    uint4* bz = reinterpret_cast<uint4*>(bm);
    for (int i = tid; i < 3 * PLANE / 4; i += NT) bz[i] = make_uint4(0, 0, 0, 0);
    if (tid < 2) lines[tid] = setup_line(row, tid);
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const LineParams L = lines[s];
        for (int i = tid; i < L.count; i += NT) {
            const int st = L.dx > 0 ? (2 * L.dy * i + L.dx - 1) / (2 * L.dx) : 0;
            const int c = L.vert ? L.x1 + st : L.x1 + i;
            const int r = L.vert ? L.y1 + L.sy * i : L.y1 + L.sy * st;
            const uint32_t bit = 1u << (c & 31);
            atomicOr(&bm[r * BMW + (c >> 5)], bit);
            if (L.rev) atomicOr(&bm[2 * PLANE + r * BMW + (c >> 5)], bit);
        }
    }
    __syncthreads();
    // columns with >= 2 lit pixels: per 32-column word, (ones, twos) carry-save over rows
    {
        constexpr int NCHUNK = NT / 8;            // row chunks; 4 of them per warp
        constexpr int CROWS = (IMG + NCHUNK - 1) / NCHUNK;
        const int w = tid & 7, chunk = tid >> 3;
        uint32_t ones = 0, twos = 0;
        const int r0 = chunk * CROWS;
#pragma unroll
        for (int k = 0; k < CROWS; ++k) {
            const int r = r0 + k;
            const uint32_t v = r < IMG ? bm[r * BMW + w] : 0u;
            twos |= ones & v;
            ones |= v;
        }
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
            const uint32_t o2 = __shfl_xor_sync(0xffffffffu, ones, off);
            const uint32_t t2 = __shfl_xor_sync(0xffffffffu, twos, off);
            twos |= t2 | (ones & o2);
            ones |= o2;
        }
        if ((tid & 31) < 8) {
            red[((tid >> 5) * 8 + w) * 2 + 0] = ones;
            red[((tid >> 5) * 8 + w) * 2 + 1] = twos;
        }
    }
    __syncthreads();
    if (tid < 8) {
        uint32_t ones = 0, twos = 0;
#pragma unroll
        for (int q = 0; q < NT / 32; ++q) {
            const uint32_t o2 = red[(q * 8 + tid) * 2 + 0], t2 = red[(q * 8 + tid) * 2 + 1];
            twos |= t2 | (ones & o2);
            ones |= o2;
        }
        colmask[tid] = twos;
    }
    __syncthreads();
    for (int i = tid; i < IMG * BMW; i += NT)
        bm[PLANE + i] = bm[i] & colmask[i & 7];             // plot_segment.py:59-65
    __syncthreads();
}


}  // namespace bitmap
}  // namespace svx
