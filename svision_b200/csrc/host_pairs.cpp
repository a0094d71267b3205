// Host-side segment-pair generation: signatures -> packed int32[12] rows, without the text BED in
// between.  Restates, for a whole chromosome at once,
//   Signature.get_segs_cords        src/collection/classes.py:72-117   (coordinates relative to the first alignment;
//                                                                       first + last alignment = main segments)
//   cord_to_segments / Segment      src/segmentplot/run_hash_lineplot.py:35-49, src/segmentplot/classes.py:44-54
//   cal_non_linear                  src/collection/output_clusters.py:213-251
//   linearOrNot                     src/collection/output_clusters.py:11-27
//   proc_one_sig (pair enumeration) src/collection/output_clusters.py:124-210
// The floating-point steps (midpoints, the non-linear score, the gap ratio) are done in IEEE double in
// the reference's order of operations, so the integers that reach the BED columns are the same.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/svx.h"

namespace svx {
void set_error(const std::string& msg);   // svx_api.cu
}

namespace {

struct Seg {
    int64_t xs, xe, ys, ye;
    bool fwd;
};

inline Seg make_seg(int64_t x_start, int64_t y_start, int64_t y_end, bool fwd) {
    const int64_t len = y_end - y_start + 1;                       // run_hash_lineplot.py:46
    Seg s;
    s.xs = x_start;
    s.ys = y_start;
    s.fwd = fwd;
    s.xe = fwd ? x_start + (len - 1) : x_start - (len - 1);        // classes.py:50-53
    s.ye = y_start + (len - 1);
    return s;
}

inline bool is_linear(const Seg& i, const Seg& j) {                // output_clusters.py:11-27
    const int64_t on_ref = j.ys - i.ye;
    int64_t on_read = j.xs - i.xe;
    if (on_read == 0) on_read = 1;
    const double ratio = static_cast<double>(on_ref) / static_cast<double>(on_read);
    if (i.fwd != j.fwd) return false;
    return !(ratio >= 1.5 || ratio <= 0.7);
}

int fail(int64_t sig, const std::string& what) {
    svx::set_error("svx_pairs_generate: signature " + std::to_string(sig) + ": " + what);
    return SVX_ERR_INVALID;
}

}  // namespace

extern "C" int svx_pairs_generate(int64_t n_sig, const int64_t* sig_aln_off, const int64_t* aln,
                                  const int64_t* sig_bkp_off, int64_t capacity, int32_t* rows, int64_t* meta,
                                  int64_t* n_rows) {
    if (n_sig < 0 || !n_rows || (n_sig && (!sig_aln_off || !aln || !sig_bkp_off)) || capacity < 0 ||
        (capacity && (!rows || !meta))) {
        svx::set_error("svx_pairs_generate: bad argument");
        return SVX_ERR_INVALID;
    }
    int64_t out = 0, needed = 0;
    std::vector<Seg> seg;
    for (int64_t s = 0; s < n_sig; ++s) {
        const int64_t a0 = sig_aln_off[s], n = sig_aln_off[s + 1] - a0;
        if (n <= 0) return fail(s, "has no alignment");
        const int64_t* first = aln + a0 * SVX_ALN_FIELDS;
        const int64_t* last = aln + (a0 + n - 1) * SVX_ALN_FIELDS;
        const int64_t ref0 = first[0], read0 = first[2];
        // all segments in the reference's order: main (first, last), then the inner ones
        seg.clear();
        auto add = [&](const int64_t* a, bool main_seg) {
            const int64_t ys = a[0] - ref0, ye = a[1] - ref0, qs = a[2] - read0, qe = a[3] - read0;
            if (main_seg || !a[4]) seg.push_back(make_seg(qs, ys, ye, true));    // classes.py:103-111: main
            else seg.push_back(make_seg(qe, ys, ye, false));                     // segments are always forward
        };
        add(first, true);
        if (n > 1) add(last, true);
        for (int64_t k = 1; k + 1 < n; ++k) add(aln + (a0 + k) * SVX_ALN_FIELDS, false);
        const int n_main = n > 1 ? 2 : 1;
        const int64_t n_other = n > 2 ? n - 2 : 0;
        const int64_t read_len = last[3] - read0, ref_len = last[1] - ref0;       // classes.py:113-114

        // non-linear score (output_clusters.py:213-251)
        double sum = 0.0;
        int64_t lo = seg[0].ys, hi = seg[0].ys;
        for (const Seg& g : seg) {
            lo = std::min(lo, std::min(g.ys, g.ye));
            hi = std::max(hi, std::max(g.ys, g.ye));
            const double ref_mid = static_cast<double>(g.xs + g.xe) / 2.0;
            const double read_mid = static_cast<double>(g.ys + g.ye) / 2.0;
            sum += std::fabs(ref_mid - read_mid) * static_cast<double>(std::llabs(g.xe - g.xs));
        }
        if (hi - lo == 0) continue;                                              // v1.3.4: the signature is skipped
        const int64_t score = static_cast<int64_t>(sum / static_cast<double>(hi - lo));

        const int64_t n_bkp = sig_bkp_off[s + 1] - sig_bkp_off[s];
        auto emit = [&](const Seg& p, const Seg& q, int64_t sub, bool main_pair, bool forward, int64_t bkp_k) -> int {
            ++needed;
            if (bkp_k >= n_bkp) return fail(s, "needs breakpoint " + std::to_string(bkp_k) + " but has " + std::to_string(n_bkp));
            if (out >= capacity) return SVX_OK;                                  // counting pass / undersized output
            const int64_t v[SVX_ROW_FIELDS] = {p.xs, p.xe, p.ys, p.ye, p.fwd, q.xs, q.xe, q.ys, q.ye, q.fwd, read_len, ref_len};
            int32_t* r = rows + out * SVX_ROW_FIELDS;
            for (int k = 0; k < SVX_ROW_FIELDS; ++k) {
                if (v[k] < INT32_MIN || v[k] > INT32_MAX) return fail(s, "coordinate does not fit int32");
                r[k] = static_cast<int32_t>(v[k]);
            }
            int64_t* m = meta + out * SVX_PAIR_META;
            m[0] = s;
            m[1] = sub;
            m[2] = (main_pair ? 1 : 0) | (forward ? 2 : 0);
            m[3] = sig_bkp_off[s] + bkp_k;
            m[4] = score;
            ++out;
            return SVX_OK;
        };
        int64_t sub = 0;
        for (int i = 0; i + 1 < n_main; ++i) {                                   // output_clusters.py:168-175
            ++sub;
            if (!is_linear(seg[i], seg[i + 1]))
                if (int rc = emit(seg[i], seg[i + 1], sub, true, true, 0)) return rc;
        }
        for (int i = 0; i < n_main; ++i)                                         // output_clusters.py:182-204
            for (int64_t k = 0; k < n_other; ++k) {
                ++sub;
                const Seg& o = seg[n_main + k];
                if (!is_linear(seg[i], o))
                    if (int rc = emit(seg[i], o, sub, false, seg[i].fwd && o.fwd, k + 1)) return rc;
            }
    }
    *n_rows = needed;
    if (needed > capacity && capacity > 0) {
        svx::set_error("svx_pairs_generate: " + std::to_string(needed) + " rows needed, capacity " + std::to_string(capacity));
        return SVX_ERR_INVALID;
    }
    return SVX_OK;
}
