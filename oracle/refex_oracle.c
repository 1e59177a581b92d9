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
