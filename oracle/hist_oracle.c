/*
 * oracle/hist_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's weighted histogramming and
 * grid->events lookup (icecube/pisa):
 *   - pisa/core/translation.py:171-205 histogram_fh -> fast_histogram.histogramdd
 *   - pisa/core/translation.py:417-501 lookup_regular_{1,2,3}d[_array]
 *   - pisa/core/translation.py:503-597 find_index / find_index_unsafe
 *   - pisa/stages/utils/hist.py:93-113 searchsorted pre-digitisation
 * Checker only; see prob3_oracle.c for the rules on who may call it.
 *
 * Third-party arithmetic: the regular-binning histogram is computed in the
 * reference by `fast_histogram` (setup.py:89 pins `fast-histogram>=0.10`; it is
 * not vendored under /root/reference and is not installed here).  Its published
 * algorithm for histogramdd is, per dimension d, with norm_d = n_d/(hi_d-lo_d):
 *     if (x_d >= lo_d && x_d < hi_d) i_d = (int)((x_d - lo_d) * norm_d); else skip
 *     flat = sum_d i_d * stride_d ; hist[flat] += w
 * which is also exactly what the reference's own lookup_regular_* use
 * (translation.py:417-455), so both directions share one index rule.
 * Parity status: index rule PINNED against the reference's test_histogram
 * (== numpy.histogramdd on non-edge samples, translation.py:779-818) and
 * test_find_index (translation.py:821-940); the upper-edge case x == hi
 * (numpy keeps it, the half-open rule drops it) and an index that rounds up to
 * n_d for x just below hi (an out-of-bounds access in the reference's own
 * lookup_regular_*) are parity-UNPINNED; the oracle folds the
 * latter into the last bin and reports the count through the return value.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef ORACLE_F32
typedef float real_t;
#define SUFFIX(name) name##_f32
#else
typedef double real_t;
#define SUFFIX(name) name##_f64
#endif

/* translation.py:560-597 find_index_unsafe */
static int64_t find_index_unsafe(double val, const double *edges, int64_t n_edges) {
    int64_t left = 0, right = n_edges - 1;
    while (left < right) {
        int64_t test = (left + right) >> 1;
        if (val >= edges[test]) left = test + 1;
        else right = test;
    }
    return left - 1;
}

/* translation.py:503-557 find_index: [b0)[b1)...[b_last], -1 under/nan, n_bins over */
int64_t SUFFIX(oracle_find_index)(double val, const double *edges, int64_t n_edges) {
    int64_t n_bins = n_edges - 1;
    if (val >= edges[0]) {
        if (val <= edges[n_edges - 1]) {
            int64_t b = find_index_unsafe(val, edges, n_edges);
            if (b < 0) b = 0;
            if (b > n_bins - 1) b = n_bins - 1;
            return b;
        }
        return n_bins;
    }
    return -1;
}

void SUFFIX(oracle_find_index_array)(const real_t *vals, int64_t n, const double *edges,
                                     int64_t n_edges, int64_t *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = SUFFIX(oracle_find_index)((double)vals[i], edges, n_edges);
}

/* hist.py:93-113: np.searchsorted(edges, x, side="right") - 1, top edge folded into last bin */
void SUFFIX(oracle_digitize_irregular)(const real_t *x, int64_t n, const double *edges,
                                       int64_t n_edges, int64_t *out) {
    for (int64_t i = 0; i < n; ++i) {
        /* searchsorted side='right': first index with edges[idx] > x; nan sorts last */
        int64_t lo = 0, hi = n_edges;
        double v = (double)x[i];
        if (v != v) lo = n_edges;
        else
            while (lo < hi) {
                int64_t mid = (lo + hi) >> 1;
                if (edges[mid] <= v) lo = mid + 1;
                else hi = mid;
            }
        int64_t idx = lo - 1;
        if (v == edges[n_edges - 1]) idx -= 1;
        out[i] = idx;
    }
}

/*
 * Flat row-major bin index per event on a regular (linear) D-dim binning,
 * -1 if outside in any dimension.  Rule of translation.py:427-455.
 * coords is D pointers to N values each.
 */
int64_t SUFFIX(oracle_regular_index)(const real_t *const *coords, int n_dims, int64_t n,
                                     const double *lo, const double *hi, const int64_t *nbins,
                                     int64_t *out) {
    /* Arithmetic is double in both modes: numba promotes float32 (op) float64 and the
     * range / norm are float64 scalars (translation.py:419,429-430); fast_histogram
     * converts its inputs to double. */
    int64_t n_overflow_rounding = 0;
    double norm[8];
    for (int d = 0; d < n_dims; ++d) norm[d] = (double)nbins[d] / (hi[d] - lo[d]);
    for (int64_t i = 0; i < n; ++i) {
        int64_t flat = 0;
        int ok = 1;
        for (int d = 0; d < n_dims; ++d) {
            double x = (double)coords[d][i];
            if (x >= lo[d] && x < hi[d]) {
                int64_t id = (int64_t)((x - lo[d]) * norm[d]);
                if (id >= nbins[d]) { ++n_overflow_rounding; id = nbins[d] - 1; }
                flat = flat * nbins[d] + id;
            } else { ok = 0; break; }
        }
        out[i] = ok ? flat : -1;
    }
    return n_overflow_rounding;
}

/* hist[idx] += w, serial event order (what a single-threaded fast_histogram /
 * numpy.histogramdd does); idx < 0 skipped.  weights may be NULL (counts). */
void SUFFIX(oracle_accumulate)(const int64_t *idx, const real_t *weights, int64_t n,
                               int64_t n_bins_total, real_t *hist) {
    for (int64_t b = 0; b < n_bins_total; ++b) hist[b] = 0;
    /* fast_histogram / numpy accumulate in double and the reference casts the
     * result to FTYPE (translation.py:205,223) */
    double *acc = (double *)__builtin_alloca(sizeof(double) * (size_t)n_bins_total);
    for (int64_t b = 0; b < n_bins_total; ++b) acc[b] = 0;
    for (int64_t i = 0; i < n; ++i)
        if (idx[i] >= 0) acc[idx[i]] += weights ? (double)weights[i] : 1.0;
    for (int64_t b = 0; b < n_bins_total; ++b) hist[b] = (real_t)acc[b];
}

/* translation.py:417-501 lookup_regular_*: out = flat_hist[index] (0 outside);
 * `width` > 1 is the *_array variant (flat_hist is [n_bins, width]). */
void SUFFIX(oracle_lookup)(const int64_t *idx, const real_t *flat_hist, int64_t n, int width,
                           real_t *out) {
    for (int64_t i = 0; i < n; ++i)
        for (int w = 0; w < width; ++w)
            out[(size_t)i * width + w] = idx[i] >= 0 ? flat_hist[(size_t)idx[i] * width + w] : (real_t)0;
}
