/* TEST INFRASTRUCTURE - C restatement of the reference-bin search, for sizes the numpy oracle is too slow for
 * and as the multi-core CPU baseline ("port") of bench.py.  Not part of the product; never linked into
 * libwisecondor_b200.so.
 *
 * Follows /root/reference/wisetools.py:
 *   :364-398 getReference   rows of the part, candidates = bins of all other chromosomes (concatenated)
 *   :298-325 getRefForBins  d_j = sum_s (other[j][s] - row[s])^2 accumulated sequentially over s with
 *                           separately rounded subtract, multiply, add (what numpy executes on the reference's
 *                           Fortran-ordered operands); streaming insert `if d < curMax: insert at
 *                           bisect_right(dists, d); drop last` starting from index -1 / distance 1e10.
 * Pinned against the reference itself by tests/test_oracle_vs_ref.py (bit-exact).
 * Build: gcc -O2 -ffp-contract=off -pthread -fPIC -shared (see oracle/Makefile); FMA contraction must stay off.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

static int bisect_right(const double* a, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (x < a[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

typedef struct {
    const double* X;
    int N, S, k, row_begin, row_end;
    const int* cs_of;
    const int* ce_of;
    int32_t* idx;
    double* dist;
    int* next_row;     /* shared work counter (rows handed out 4 at a time) */
} job_t;

static void search_row(const job_t* J, int t) {
    const int S = J->S, k = J->k, N = J->N;
    const double* row = J->X + (size_t)t * S;
    int32_t* oi = J->idx + (size_t)(t - J->row_begin) * k;
    double* od = J->dist + (size_t)(t - J->row_begin) * k;
    for (int i = 0; i < k; ++i) { oi[i] = -1; od[i] = 1e10; }
    double cur_max = 1e10;
    const int cs = J->cs_of[t], ce = J->ce_of[t];
    int other = 0;                          /* position in the other-chromosome concatenation */
    for (int j = 0; j < N; ++j) {
        if (j >= cs && j < ce) continue;
        const double* xj = J->X + (size_t)j * S;
        double acc = 0.0;
        for (int s = 0; s < S; ++s) {
            double v = xj[s] - row[s];
            acc = acc + v * v;
        }
        if (acc < cur_max) {
            int p = bisect_right(od, k, acc);
            memmove(od + p + 1, od + p, sizeof(double) * (size_t)(k - 1 - p));
            memmove(oi + p + 1, oi + p, sizeof(int32_t) * (size_t)(k - 1 - p));
            od[p] = acc;
            oi[p] = other;
            cur_max = od[k - 1];
        }
        ++other;
    }
}

static void* worker(void* arg) {
    const job_t* J = (const job_t*)arg;
    for (;;) {
        int t0 = __atomic_fetch_add(J->next_row, 4, __ATOMIC_RELAXED);
        if (t0 >= J->row_end) break;
        int t1 = t0 + 4 < J->row_end ? t0 + 4 : J->row_end;
        for (int t = t0; t < t1; ++t) search_row(J, t);
    }
    return NULL;
}

int wc_oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* Returns 0 on success.  idx/dist are (row_end-row_begin) x k, row-major.  nthreads <= 0: all online cores. */
int wc_oracle_get_reference(const double* X, int N, int S, const int* chrom_bins, int nchrom, int row_begin,
                            int row_end, int k, int32_t* idx, double* dist, int nthreads) {
    int* cs_of = (int*)malloc(sizeof(int) * (size_t)N);
    int* ce_of = (int*)malloc(sizeof(int) * (size_t)N);
    if (!cs_of || !ce_of) return -1;
    int pos = 0;
    for (int c = 0; c < nchrom; ++c) {
        for (int i = 0; i < chrom_bins[c]; ++i) { cs_of[pos + i] = pos; ce_of[pos + i] = pos + chrom_bins[c]; }
        pos += chrom_bins[c];
    }
    if (pos != N) { free(cs_of); free(ce_of); return -2; }
    if (nthreads <= 0) nthreads = wc_oracle_max_threads();
    if (nthreads > 256) nthreads = 256;
    int next = row_begin;
    job_t J = {X, N, S, k, row_begin, row_end, cs_of, ce_of, idx, dist, &next};
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nthreads - 1; ++i)
        if (pthread_create(&th[started], NULL, worker, &J) == 0) ++started;
    worker(&J);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
    free(cs_of);
    free(ce_of);
    return 0;
}
