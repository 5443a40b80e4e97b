/* CPU oracle (float64, C + OpenMP) for the PSMC coalescent-HMM hot path of jthlab/phlash.
 *
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs through oracle/c_oracle.py.  The product
 * (phlash_b200/) never links or loads it.
 *
 * Restates, from the maths:
 *   - the O(M) structured transition  x -> x A,
 *       (xA)[j] = b[j] * sum_{i>j} x[i] + d[j] * x[j] + v[j] * sum_{i<j} u[i] x[i]
 *     (reference: src/phlash/hmm.py:52-65, device twin src/phlash/gpu.py:504-522);
 *   - the scaled forward recursion, transition -> emission -> rescale, missing (-1)
 *     observations emitting 1 (reference: src/phlash/hmm.py:68-82);
 *   - the gradient contract of the reference kernel after the host-side roll:
 *     d ll / d log(theta) for rows b,d,u,v,emis0,emis1 and pi * d ll / d pi for the pi row
 *     (reference: src/phlash/gpu.py:575-692 and :303-313), computed by the adjoint recursion.
 *
 * Pinning: agrees with oracle/psmc_oracle.py (NumPy) to round-off, which in turn is pinned
 * against golden vectors produced by the reference's own sources (tests/golden/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { ROW_B = 0, ROW_D, ROW_U, ROW_V, ROW_E0, ROW_E1, ROW_PI, NROWS };

/* out = (x A) .* emis(ob); returns the sum of out */
static inline double step_forward(int M, const double *pp, const double *x, int ob, double *out)
{
    const double *b = pp + ROW_B * M, *d = pp + ROW_D * M, *u = pp + ROW_U * M, *v = pp + ROW_V * M;
    double acc = 0.0;
    for (int j = 0; j < M; ++j) {
        out[j] = d[j] * x[j] + v[j] * acc;
        acc += u[j] * x[j];
    }
    acc = 0.0;
    for (int j = M - 1; j >= 0; --j) {
        out[j] += b[j] * acc;
        acc += x[j];
    }
    double total = 0.0;
    if (ob >= 0) {
        const double *e = pp + (ROW_E0 + ob) * M;
        for (int j = 0; j < M; ++j) {
            out[j] *= e[j];
            total += out[j];
        }
    } else {
        for (int j = 0; j < M; ++j) total += out[j];
    }
    return total;
}

static double pair_loglik(int M, const double *pp, const int8_t *row, int64_t L, double *final_alpha)
{
    double cur[256], nxt[256];
    memcpy(cur, pp + ROW_PI * M, sizeof(double) * M);
    double ll = 0.0;
    for (int64_t s = 0; s < L; ++s) {
        double total = step_forward(M, pp, cur, row[s], nxt);
        double inv = 1.0 / total;
        for (int j = 0; j < M; ++j) cur[j] = nxt[j] * inv;
        ll += log(total);
    }
    if (final_alpha) memcpy(final_alpha, cur, sizeof(double) * M);
    return ll;
}

/* adjoint recursion; work = (L+1)*M doubles */
static double pair_loglik_grad(int M, const double *pp, const int8_t *row, int64_t L, double *grad, double *work)
{
    const double *b = pp + ROW_B * M, *d = pp + ROW_D * M, *u = pp + ROW_U * M, *v = pp + ROW_V * M;
    double *alpha = work;
    memcpy(alpha, pp + ROW_PI * M, sizeof(double) * M);
    double ll = 0.0;
    for (int64_t s = 0; s < L; ++s) {
        double *nxt = alpha + (s + 1) * M;
        double total = step_forward(M, pp, alpha + s * M, row[s], nxt);
        double inv = 1.0 / total;
        for (int j = 0; j < M; ++j) nxt[j] *= inv;
        ll += log(total);
    }
    memset(grad, 0, sizeof(double) * NROWS * M);
    double beta[256], w[256], tail[256], bnew[256];
    for (int j = 0; j < M; ++j) beta[j] = 1.0;
    for (int64_t s = L - 1; s >= 0; --s) {
        const int ob = row[s];
        const double *x = alpha + s * M, *post = alpha + (s + 1) * M;
        if (ob >= 0) {
            const double *e = pp + (ROW_E0 + ob) * M;
            double *ge = grad + (ROW_E0 + ob) * M;
            for (int j = 0; j < M; ++j) {
                ge[j] += post[j] * beta[j];
                w[j] = e[j] * beta[j];
            }
        } else {
            for (int j = 0; j < M; ++j) w[j] = beta[j];
        }
        /* bnew = A w */
        double acc = 0.0;
        for (int j = M - 1; j >= 0; --j) {
            tail[j] = acc;
            acc += v[j] * w[j];
        }
        acc = 0.0;
        double z = 0.0;
        for (int i = 0; i < M; ++i) {
            bnew[i] = acc + d[i] * w[i] + u[i] * tail[i];
            acc += b[i] * w[i];
            z += x[i] * bnew[i];
        }
        const double scale = 1.0 / z;
        double above = 0.0;
        for (int j = M - 1; j >= 0; --j) {
            grad[ROW_B * M + j] += b[j] * above * w[j] * scale;
            above += x[j];
        }
        double weighted = 0.0;
        for (int j = 0; j < M; ++j) {
            const double ws = w[j] * scale;
            grad[ROW_D * M + j] += d[j] * x[j] * ws;
            grad[ROW_V * M + j] += v[j] * weighted * ws;
            grad[ROW_U * M + j] += u[j] * x[j] * tail[j] * scale;
            weighted += u[j] * x[j];
            beta[j] = bnew[j] * scale;
        }
    }
    for (int j = 0; j < M; ++j) grad[ROW_PI * M + j] = pp[ROW_PI * M + j] * beta[j];
    return ll;
}

/* params: [n_pairs, 7, M]; rows addressed as data + row_index[p] * pitch.
 * dlog may be NULL (forward only); final_alpha may be NULL.  Returns 0, or -1 on bad M / OOM. */
int oracle_loglik_batch(int M, const int8_t *data, int64_t pitch, int64_t L, int64_t n_pairs,
                        const int64_t *row_index, const double *params, double *ll, double *dlog,
                        double *final_alpha, int n_threads)
{
    if (M < 1 || M > 256) return -1;
    int failed = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        double *work = dlog ? (double *)malloc(sizeof(double) * (size_t)(L + 1) * M) : NULL;
        if (dlog && !work) {
#pragma omp atomic write
            failed = 1;
        }
#pragma omp for schedule(dynamic, 1)
        for (int64_t p = 0; p < n_pairs; ++p) {
            if (dlog && !work) continue;
            const double *pp = params + p * NROWS * M;
            const int8_t *row = data + row_index[p] * pitch;
            if (dlog)
                ll[p] = pair_loglik_grad(M, pp, row, L, dlog + p * NROWS * M, work);
            else
                ll[p] = pair_loglik(M, pp, row, L, final_alpha ? final_alpha + p * M : NULL);
        }
        free(work);
    }
    return failed ? -1 : 0;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
