/*
 * ORACLE (test infrastructure, not product): C restatement of hot path A in float64.
 *
 * Follows /root/reference/graphrole/features/extract.py:105-118: for every requested node the
 * previous-generation feature rows of its out-neighbours are summed and averaged; an empty
 * neighbourhood gives 0 for both (fillna(0), :113).  Input features are the float32 values the
 * GPU kernel reads; accumulation is float64 so the result is the exact-arithmetic answer to
 * ~1e-16 relative.  Pinned against the Python restatements (which are pinned against the
 * reference's golden vectors) in tests/test_oracle_refex.py.
 *
 * Also the multi-threaded "fair" CPU comparator timed by bench.py (OpenMP over rows).
 */
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int refex_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* rows: n_sel row numbers; X: [*, ldx] float32; out_sum/out_mean: [n_sel, d] float64 */
int refex_oracle_rows_f32(int64_t n_sel, const int64_t* rows, const int64_t* rowptr,
                          const int32_t* colidx, const float* X, int64_t ldx, int32_t d,
                          double* out_sum, double* out_mean, int32_t threads) {
    if (n_sel < 0 || d < 1 || ldx < d) return 1;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t s = 0; s < n_sel; ++s) {
        const int64_t r = rows ? rows[s] : s;
        const int64_t beg = rowptr[r], end = rowptr[r + 1];
        double* sum = out_sum + s * (int64_t)d;
        double* mean = out_mean + s * (int64_t)d;
        for (int32_t c = 0; c < d; ++c) sum[c] = 0.0;
        for (int64_t k = beg; k < end; ++k) {
            const float* x = X + (int64_t)colidx[k] * ldx;
            for (int32_t c = 0; c < d; ++c) sum[c] += (double)x[c];
        }
        const double deg = (double)(end - beg);
        for (int32_t c = 0; c < d; ++c) mean[c] = end > beg ? sum[c] / deg : 0.0;
    }
    return 0;
}

/* One whole recursion level in float64 (the reference's arithmetic, extract.py:105-118, carried
 * through the recursion of extract.py:77-83 without ever rounding to float32), compared in place
 * with what the GPU produced for the same level:
 *   X        [n, d] float64   level input (previous level's float64 means, or the base features)
 *   gpu      [n, 2d] float32  the GPU's [sum block | mean block] for this level
 *   next     [n, d] float64   out: this level's float64 means (the next level's input)
 *   err[0..1]                 out: max relative error of the GPU's sums / means against float64,
 *                             |g - r| / |r| over entries with r != 0
 *   err[2]                    out: number of entries where r == 0 but g != 0 (must be 0)
 * Rows are independent; OpenMP over rows. */
int refex_oracle_level_check_f64(int64_t n, const int64_t* rowptr, const int32_t* colidx,
                                 const double* X, int32_t d, const float* gpu, double* next,
                                 double* err, int32_t threads) {
    if (n < 0 || d < 1 || d > 4096) return 1;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    double worst_sum = 0.0, worst_mean = 0.0, bad_zero = 0.0;
#pragma omp parallel for schedule(dynamic, 256) reduction(max : worst_sum, worst_mean) \
    reduction(+ : bad_zero)
    for (int64_t r = 0; r < n; ++r) {
        double sum[4096];
        const int64_t beg = rowptr[r], end = rowptr[r + 1];
        for (int32_t c = 0; c < d; ++c) sum[c] = 0.0;
        for (int64_t k = beg; k < end; ++k) {
            const double* x = X + (int64_t)colidx[k] * d;
            for (int32_t c = 0; c < d; ++c) sum[c] += x[c];
        }
        const double deg = (double)(end - beg);
        const float* g = gpu + r * 2 * (int64_t)d;
        for (int32_t c = 0; c < d; ++c) {
            const double mean = end > beg ? sum[c] / deg : 0.0;
            next[r * (int64_t)d + c] = mean;
            const double es = (double)g[c] - sum[c], em = (double)g[d + c] - mean;
            if (sum[c] != 0.0) {
                const double rel = (es < 0 ? -es : es) / (sum[c] < 0 ? -sum[c] : sum[c]);
                if (rel > worst_sum) worst_sum = rel;
            } else if (es != 0.0) {
                bad_zero += 1.0;
            }
            if (mean != 0.0) {
                const double rel = (em < 0 ? -em : em) / (mean < 0 ? -mean : mean);
                if (rel > worst_mean) worst_mean = rel;
            } else if (em != 0.0) {
                bad_zero += 1.0;
            }
        }
    }
    err[0] = worst_sum;
    err[1] = worst_mean;
    err[2] = bad_zero;
    return 0;
}
