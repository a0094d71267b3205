// Host-side per-row replay and region aggregation of one chromosome, the integer
// and floating-point part of the step immediately after the GPU path.  It replaces
//   * the per-row loop of Predict.run                       src/network/predict.py:213-300
//   * Predict.get_region_potential_svtypes                  src/network/predict.py:29-145
//   * the numeric part of write_results_to_vcf              src/network/output.py:473-474,495-496,525-529,551-552
// and leaves text assembly, type refinement and genotyping to svision_b200/calls.py, which now only
// touches the candidates that reach min_support (a few per cent of the rows).
//
// Everything that ends up printed must equal what the reference's Python prints, so the two numpy
// reductions involved are restated operation by operation:
//   numpy.mean(list of numpy.float32)  = float32 pairwise sum (8 accumulators, blocks of 128), divided by the
//                                        count in float64 and cast back to float32
//   numpy.std(list of int)             = float64: mean by the same reduce, squared deviations, reduce,
//                                        / n, sqrt
// (tests/test_calls.py fuzzes both against numpy itself).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/svx.h"

namespace svx {
void set_error(const std::string& msg);   // svx_api.cu
}
using svx::set_error;

namespace {

// numpy/_core/src/umath/loops_utils.h.src: @TYPE@_pairwise_sum
template <typename T>
T pairwise_sum(const T* a, int64_t n) {
    if (n < 8) {
        T res = 0;
        for (int64_t i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        T r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        T res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int64_t n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
}

// numpy.add.reduce over a contiguous 1-D array of the reduction's own dtype: one inner-loop call, i.e.
// the pairwise sum of all n elements (checked against numpy for n up to 70 000)
template <typename T>
T add_reduce(const T* a, int64_t n) {
    return n <= 0 ? T(0) : pairwise_sum(a, n);
}

// numpy's _mean divides the float32 sum by a numpy.intp count: float32 / int64 promotes to float64, and the
// quotient is then cast back to float32
float np_mean_f32(const float* a, int64_t n) {
    return static_cast<float>(static_cast<double>(add_reduce(a, n)) / static_cast<double>(n));
}

double np_std_i64(const int64_t* a, int64_t n, std::vector<double>& tmp) {
    tmp.resize(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) tmp[i] = static_cast<double>(a[i]);
    const double mean = add_reduce(tmp.data(), n) / static_cast<double>(n);
    for (int64_t i = 0; i < n; ++i) {
        const double d = static_cast<double>(a[i]) - mean;
        tmp[i] = d * d;
    }
    return std::sqrt(add_reduce(tmp.data(), n) / static_cast<double>(n));
}

struct Span {
    const char* p;
    int64_t n;
    bool operator==(const Span& o) const { return n == o.n && std::memcmp(p, o.p, static_cast<size_t>(n)) == 0; }
};

bool parse_int(const Span& f, int64_t* out) {
    const char* p = f.p;
    const char* e = f.p + f.n;
    if (p == e) return false;
    bool neg = false;
    if (*p == '-' || *p == '+') { neg = *p == '-'; ++p; }
    if (p == e || e - p > 18) return false;
    int64_t v = 0;
    for (; p < e; ++p) {
        const unsigned d = static_cast<unsigned>(*p - '0');
        if (d > 9) return false;
        v = v * 10 + d;
    }
    *out = neg ? -v : v;
    return true;
}

// one read of a region: its key (the read id without 'm'), the last row that named it, its calls
struct Read {
    std::string key;
    int64_t last_row = -1;
    bool has_call = false;
    int64_t call_order = -1;           // order of first insertion into the calls dict
    bool present[SVX_NUM_CLASSES] = {};
    int64_t bkp[SVX_NUM_CLASSES][3] = {};
};

struct Group {
    int kinds_mask = 0;
    std::vector<int> read_idx;          // indices into the region's reads, in arrival order
    int64_t bkp[SVX_NUM_CLASSES][3] = {};
    int first_seen = 0;
};

}  // namespace

extern "C" {

int svx_np_mean_f32(const float* values, int64_t n, float* out) {
    if (!values || !out || n <= 0) { set_error("svx_np_mean_f32: bad arguments"); return SVX_ERR_INVALID; }
    *out = np_mean_f32(values, n);
    return SVX_OK;
}

int svx_np_std_i64(const int64_t* values, int64_t n, double* out) {
    if (!values || !out || n <= 0) { set_error("svx_np_std_i64: bad arguments"); return SVX_ERR_INVALID; }
    std::vector<double> tmp;
    *out = np_std_i64(values, n, tmp);
    return SVX_OK;
}

int svx_calls_aggregate(const char* text, int64_t len, int64_t n, const int64_t* spans, const int32_t* flags,
                        const int64_t* bkp_start, const int64_t* bkp_end, const int64_t* bkp_len,
                        const int32_t* labels, const float* win, int64_t min_support, int64_t* cand,
                        double* qual, int64_t* reads_out, int64_t* n_cand, int64_t* n_reads) {
    if (!n_cand || !n_reads) { set_error("svx_calls_aggregate: NULL count pointers"); return SVX_ERR_INVALID; }
    *n_cand = 0;
    *n_reads = 0;
    if (n < 0 || (n > 0 && (!text || !spans || !flags || !bkp_start || !bkp_end || !bkp_len || !labels || !win ||
                            !cand || !qual || !reads_out))) {
        set_error("svx_calls_aggregate: bad arguments");
        return SVX_ERR_INVALID;
    }
    auto span = [&](int64_t row, int col) -> Span {
        const int64_t* s = spans + (row * SVX_BED_SPANS + col) * 2;
        return Span{text + s[0], s[1]};
    };
    for (int64_t i = 0; i < n; ++i) {
        if (labels[i] < 0 || labels[i] >= SVX_NUM_CLASSES) {
            set_error("svx_calls_aggregate: label out of range at row " + std::to_string(i));
            return SVX_ERR_INVALID;
        }
        for (int c = 0; c < SVX_BED_SPANS; ++c) {
            const int64_t* s = spans + (i * SVX_BED_SPANS + c) * 2;
            if (s[0] < 0 || s[1] < 0 || s[0] + s[1] > len) {
                set_error("svx_calls_aggregate: span outside the text at row " + std::to_string(i));
                return SVX_ERR_INVALID;
            }
        }
    }

    std::vector<Read> reads;
    std::vector<float> wins;
    std::vector<Group> groups;
    std::vector<int64_t> scores;
    std::vector<double> tmp;
    int64_t region_row = -1, n_types = 0, n_uncovered = 0, call_counter = 0;
    int64_t nc = 0, nr = 0;
    int rc = SVX_OK;

    auto flush = [&]() {
        if (region_row < 0) return;
        // ---- get_region_potential_svtypes: reads with the same class set form one candidate --------
        std::vector<int> order;                              // reads that carry a call, in dict order
        for (int r = 0; r < (int)reads.size(); ++r)
            if (reads[r].has_call) order.push_back(r);
        // insertion order of the calls dict = order of first call, not of first appearance
        for (size_t a = 1; a < order.size(); ++a) {          // tiny: insertion sort, stable
            const int v = order[a];
            size_t b = a;
            while (b > 0 && reads[order[b - 1]].call_order > reads[v].call_order) { order[b] = order[b - 1]; --b; }
            order[b] = v;
        }
        groups.clear();
        for (int r : order) {
            const Read& rd = reads[r];
            int mask = 0;
            for (int k = 0; k < SVX_NUM_CLASSES; ++k) mask |= rd.present[k] ? (1 << k) : 0;
            Group* g = nullptr;
            for (Group& x : groups)
                if (x.kinds_mask == mask) { g = &x; break; }
            if (!g) {
                groups.emplace_back();
                g = &groups.back();
                g->kinds_mask = mask;
                g->first_seen = (int)groups.size();
                for (int k = 0; k < SVX_NUM_CLASSES; ++k)
                    for (int j = 0; j < 3; ++j) g->bkp[k][j] = rd.bkp[k][j];
                g->read_idx.push_back(r);
                continue;
            }
            const int64_t m = (int64_t)g->read_idx.size();      // running integer mean, read by read
            for (int k = 0; k < SVX_NUM_CLASSES; ++k) {
                if (!(mask >> k & 1)) continue;
                for (int j = 0; j < 3; ++j) {
                    const double q = (double)(rd.bkp[k][j] + g->bkp[k][j] * m) / (double)(m + 1);
                    g->bkp[k][j] = (int64_t)q;                   // int(): toward zero
                }
            }
            g->read_idx.push_back(r);
        }
        if (groups.empty()) return;
        // ranked by support, ties in first-seen order (stable)
        std::vector<int> rank(groups.size());
        for (size_t i = 0; i < rank.size(); ++i) rank[i] = (int)i;
        for (size_t a = 1; a < rank.size(); ++a) {
            const int v = rank[a];
            size_t b = a;
            while (b > 0 && groups[rank[b - 1]].read_idx.size() < groups[v].read_idx.size()) { rank[b] = rank[b - 1]; --b; }
            rank[b] = v;
        }
        // ---- numeric part of write_results_to_vcf ---------------------------------------------------
        const float mean_score = np_mean_f32(wins.data(), (int64_t)wins.size());
        const float rounded = std::nearbyint(mean_score * 100.0f) / 100.0f;    // round(np.float32, 2)
        const float class_penalty = (1.0f - rounded) * 100.0f;
        const bool uncovered = n_uncovered > 0 && (double)n_uncovered >= 0.75 * (double)n_types;
        for (int gi : rank) {
            const Group& g = groups[gi];
            const int64_t support = (int64_t)g.read_idx.size();
            if (support < min_support) continue;
            scores.resize((size_t)support);
            for (int64_t s = 0; s < support; ++s) {
                const Read& rd = reads[g.read_idx[s]];
                if (!parse_int(span(rd.last_row, 4), &scores[s])) {
                    set_error("svx_calls_aggregate: signature score is not an integer at row " +
                              std::to_string(rd.last_row));
                    rc = SVX_ERR_INVALID;
                    return;
                }
                reads_out[nr + s] = rd.last_row;
            }
            const double spread = np_std_i64(scores.data(), support, tmp) / (double)support;
            int64_t* c = cand + nc * SVX_CAND_FIELDS;
            c[0] = region_row;
            c[1] = support;
            c[3] = uncovered ? 1 : 0;
            c[4] = nr;
            int nk = 0;
            for (int k = 0; k < SVX_NUM_CLASSES; ++k) {
                if (!(g.kinds_mask >> k & 1)) continue;
                c[5 + nk] = k;
                for (int j = 0; j < 3; ++j) c[10 + nk * 3 + j] = g.bkp[k][j];
                ++nk;
            }
            c[2] = nk;
            qual[nc] = spread + (double)class_penalty;
            nr += support;
            ++nc;
        }
    };

    for (int64_t i = 0; i < n && rc == SVX_OK; ++i) {
        const int32_t fl = flags[i];
        const int pred = labels[i];
        if (fl & SVX_BED_FLAG_COMPLEMENT) continue;                        // predict.py:214
        if ((fl & SVX_BED_FLAG_FORWARD) && pred == 2) continue;           // predict.py:229-231
        if (region_row < 0 || !(span(i, 0) == span(region_row, 0))) {     // predict.py:235-247
            flush();
            if (rc != SVX_OK) break;
            region_row = i;
            reads.clear();
            wins.clear();
            n_types = n_uncovered = 0;
            call_counter = 0;
        }
        const Span rn = span(i, 1);
        const bool main_pair = (fl & SVX_BED_FLAG_MAIN) != 0;
        std::string key;
        key.reserve((size_t)rn.n);
        for (int64_t k = 0; k < rn.n; ++k)
            if (!(main_pair && rn.p[k] == 'm')) key.push_back(rn.p[k]);
        Read* rd = nullptr;
        for (Read& x : reads)
            if (x.key == key) { rd = &x; break; }
        if (!rd) {
            reads.emplace_back();
            rd = &reads.back();
            rd->key = std::move(key);
        }
        rd->last_row = i;                                                  // names / scores: last write wins
        ++n_types;
        n_uncovered += (fl & SVX_BED_FLAG_UNCOVERED) ? 1 : 0;
        wins.push_back(win[i]);
        if (!main_pair && pred < 2) continue;                              // predict.py:279-281
        if (!rd->has_call) {
            rd->has_call = true;
            rd->call_order = call_counter++;
        }
        rd->present[pred] = true;
        rd->bkp[pred][0] = bkp_start[i];
        rd->bkp[pred][1] = bkp_end[i];
        rd->bkp[pred][2] = bkp_len[i];
    }
    if (rc == SVX_OK) flush();                                             // predict.py:298-300
    if (rc != SVX_OK) return rc;
    *n_cand = nc;
    *n_reads = nr;
    return SVX_OK;
}

}  // extern "C"
