// Host-side reader of SVision's <chrom>.segments.all.bed.
//
// One pass over the text turns the 23 tab-separated columns of every line into the packed
// int32[12] row the GPU path consumes, the three breakpoint integers, a few per-row flags and the
// byte spans of the string columns the post-classification step still needs.  It replaces
// BatchGenerator.read_class_list (src/network/create_batch.py:29-61: per-line split + '_'.join) and
// the per-image int()/strand parsing of next_batch (create_batch.py:103-137).
//
// Column map (writer: src/collection/output_clusters.py:180-182,207-209):
//   0 region | 1-5 seg1 xS xE yS yE fwd | 6-10 seg2 | 11 read_len | 12 ref_len | 13 read id |
//   14 sub id | 15 qname | 16 sig type | 17-18 bkp start/end | 19 score | 20 forward | 21 mechanism |
//   22 bkp len
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/svx.h"

namespace svx {
void set_error(const std::string& msg);   // svx_api.cu
}
using svx::set_error;

namespace {

constexpr int kCols = 23;
constexpr int kSpanCols[SVX_BED_SPANS] = {0, 13, 15, 16, 19, 20, 21};

struct Field {
    const char* p;
    int64_t n;
};

inline bool equals(const Field& f, const char* lit) {
    const size_t n = std::strlen(lit);
    return static_cast<size_t>(f.n) == n && std::memcmp(f.p, lit, n) == 0;
}

// optional sign + decimal digits, nothing else (Python's int() also takes blanks and '_': the
// reference's writer never emits them)
inline bool parse_int(const Field& f, int64_t* out) {
    const char* p = f.p;
    const char* e = f.p + f.n;
    if (p == e) return false;
    bool neg = false;
    if (*p == '-' || *p == '+') {
        neg = *p == '-';
        ++p;
    }
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

inline const char* line_end(const char* p, const char* end) {
    const void* nl = std::memchr(p, '\n', static_cast<size_t>(end - p));
    return nl ? static_cast<const char*>(nl) : end;
}

int fail(int64_t line_no, const std::string& what) {
    set_error("segments BED line " + std::to_string(line_no) + ": " + what);
    return SVX_ERR_INVALID;
}

}  // namespace

extern "C" int svx_bed_count_rows(const char* text, int64_t len, int64_t* n_rows) {
    if ((!text && len) || len < 0 || !n_rows) {
        set_error("svx_bed_count_rows: bad argument");
        return SVX_ERR_INVALID;
    }
    int64_t n = 0;
    const char* end = text + len;
    for (const char* p = text; p < end;) {
        const char* e = line_end(p, end);
        if (e > p) ++n;                       // blank lines carry no site
        p = e + 1;
    }
    *n_rows = n;
    return SVX_OK;
}

extern "C" int svx_bed_parse(const char* text, int64_t len, int64_t n_rows, int32_t* rows, int64_t* bkp,
                             int64_t* spans, int32_t* flags) {
    if ((!text && len) || len < 0 || n_rows < 0 || (n_rows && (!rows || !bkp || !spans || !flags))) {
        set_error("svx_bed_parse: bad argument");
        return SVX_ERR_INVALID;
    }
    const char* end = text + len;
    int64_t i = 0, line_no = 0;
    Field prev_region{nullptr, 0};
    for (const char* p = text; p < end;) {
        const char* e = line_end(p, end);
        ++line_no;
        if (e == p) {
            p = e + 1;
            continue;
        }
        if (i >= n_rows) return fail(line_no, "more rows than the caller sized the outputs for");
        Field f[kCols];
        int nf = 0;
        const char* q = p;
        while (nf < kCols) {
            const void* tab = std::memchr(q, '\t', static_cast<size_t>(e - q));
            const char* fe = tab ? static_cast<const char*>(tab) : e;
            f[nf++] = Field{q, fe - q};
            if (!tab) break;
            q = fe + 1;
        }
        if (nf < kCols) return fail(line_no, "expected 23 tab-separated columns, found " + std::to_string(nf));
        // a '\r' before the newline would stay in the last column, as line.strip('\n') leaves it
        // (create_batch.py:42); int() tolerates it, so do we
        Field last = f[22];
        if (last.n && last.p[last.n - 1] == '\r') --last.n;

        static const int int_cols[10] = {1, 2, 3, 4, 6, 7, 8, 9, 11, 12};
        static const int row_slot[10] = {0, 1, 2, 3, 5, 6, 7, 8, 10, 11};
        int32_t* r = rows + i * SVX_ROW_FIELDS;
        for (int k = 0; k < 10; ++k) {
            int64_t v;
            if (!parse_int(f[int_cols[k]], &v))
                return fail(line_no, "column " + std::to_string(int_cols[k]) + " is not an integer");
            if (v < INT32_MIN || v > INT32_MAX)
                return fail(line_no, "column " + std::to_string(int_cols[k]) + " does not fit int32");
            r[row_slot[k]] = static_cast<int32_t>(v);
        }
        // 'True' -> forward; 'False' and anything else take the reverse branch (create_batch.py:111-116,
        // src/segmentplot/classes.py:50-53)
        r[4] = equals(f[5], "True") ? 1 : 0;
        r[9] = equals(f[10], "True") ? 1 : 0;

        int64_t* b = bkp + i * 3;
        if (!parse_int(f[17], &b[0])) return fail(line_no, "column 17 is not an integer");
        if (!parse_int(f[18], &b[1])) return fail(line_no, "column 18 is not an integer");
        if (!parse_int(last, &b[2])) return fail(line_no, "column 22 is not an integer");

        int64_t* s = spans + i * SVX_BED_SPANS * 2;
        for (int k = 0; k < SVX_BED_SPANS; ++k) {
            s[2 * k] = f[kSpanCols[k]].p - text;
            s[2 * k + 1] = f[kSpanCols[k]].n;
        }
        int32_t fl = 0;
        if (std::memchr(f[13].p, 'm', static_cast<size_t>(f[13].n))) fl |= SVX_BED_FLAG_MAIN;
        if (equals(f[20], "True")) fl |= SVX_BED_FLAG_FORWARD;
        if (equals(f[16], "sigUncovered")) fl |= SVX_BED_FLAG_UNCOVERED;
        if (prev_region.p && prev_region.n == f[0].n && std::memcmp(prev_region.p, f[0].p, static_cast<size_t>(f[0].n)) == 0)
            fl |= SVX_BED_FLAG_SAME_REGION;
        prev_region = f[0];
        for (int k = 0; k < SVX_BED_SPANS; ++k) {                        // predict.py:214
            const Field& c = f[kSpanCols[k]];
            static const char kWord[] = "complement";
            if (c.n >= 10 && std::search(c.p, c.p + c.n, kWord, kWord + 10) != c.p + c.n) fl |= SVX_BED_FLAG_COMPLEMENT;
        }
        flags[i] = fl;
        ++i;
        p = e + 1;
    }
    if (i != n_rows) {
        set_error("svx_bed_parse: the text holds " + std::to_string(i) + " rows, the caller announced " + std::to_string(n_rows));
        return SVX_ERR_INVALID;
    }
    return SVX_OK;
}
