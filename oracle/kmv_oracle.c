/* kmv_oracle.c -- plain-C CPU restatement of the K.V hot path.  TEST INFRASTRUCTURE / CPU BASELINE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this; the product
 * (librpgp.so) never does.  It follows the reference's own dense arithmetic, evaluated row block by row block the way
 * gpytorch's checkpoint_kernel schedule does (gp_experiment_runner.py:250,330), without materialising K:
 *
 *   per coordinate   (x1_[:, i] - x2_[:, i])^2 / -2 -> exp -> accumulate      gp_models/kernels/memory_efficient_gam_kernel.py:21-29
 *   per group (K>1)  exp(-1/2 * sum_m diff_m^2) scaled by the outputscale      polynomial_projection_kernels.py:84-103
 *   then             out[i, :] += k(i, i') * V[i', :]                           (`K @ rhs` of LazyTensor._matmul)
 *
 * Parity pinning: checked against the numpy oracle (oracle/rpgp_oracle.py), which is itself pinned to the reference's
 * golden vectors G1-G5 and to fixtures generated from the reference (tests/test_oracle_golden.py).
 * POSIX threads over interleaved row blocks (no OpenMP runtime dependency); `threads` <= 0 means "all online cores".
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define COL_BLOCK 256
#define ROW_BLOCK 16
#define MAX_THREADS 256

int oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > MAX_THREADS) n = MAX_THREADS;
    return (int)n;
}

typedef struct {
    const void *Z1, *Z2, *c, *V;
    void* out;
    long m, n;
    int J, K, t, tid, nthreads, is_f64;
} job_t;

/* float32 arithmetic, as the reference runs without --double.  The column block of Z2 is transposed into zt[q][b] first,
 * so that, like the reference's per-dimension vector ops (memory_efficient_gam_kernel.py:21-29), the inner loops run
 * over contiguous columns and the compiler can vectorise expf. */
static void rows_f32(const job_t* jb, long r0, long r1) {
    const float *Z1 = jb->Z1, *Z2 = jb->Z2, *c = jb->c, *V = jb->V;
    float* out = jb->out;
    const int J = jb->J, K = jb->K, t = jb->t, JK = J * K;
    const long n = jb->n;
    float kv[COL_BLOCK], sq[COL_BLOCK];
    float* zt = (float*)malloc(sizeof(float) * (size_t)JK * COL_BLOCK);
    for (long i = r0; i < r1; ++i)
        for (int q = 0; q < t; ++q) out[i * t + q] = 0.f;
    for (long c0 = 0; c0 < n; c0 += COL_BLOCK) {
        const long nb = (n - c0 < COL_BLOCK) ? (n - c0) : COL_BLOCK;
        for (long b = 0; b < nb; ++b)
            for (int q = 0; q < JK; ++q) zt[(long)q * COL_BLOCK + b] = Z2[(c0 + b) * JK + q];
        for (long i = r0; i < r1; ++i) {
            const float* a = Z1 + i * JK;
            float* o = out + i * t;
            for (long b = 0; b < nb; ++b) kv[b] = 0.f;
            for (int j = 0; j < J; ++j) {
                const float cj = c[j];
                for (long b = 0; b < nb; ++b) sq[b] = 0.f;
                for (int mm = 0; mm < K; ++mm) {
                    const float aq = a[j * K + mm];
                    const float* zq = zt + (long)(j * K + mm) * COL_BLOCK;
                    for (long b = 0; b < nb; ++b) {
                        const float d = aq - zq[b];
                        sq[b] += d * d;
                    }
                }
                for (long b = 0; b < nb; ++b) kv[b] += cj * expf(-0.5f * sq[b]);
            }
            for (long b = 0; b < nb; ++b) {
                const float* v = V + (c0 + b) * t;
                const float kb = kv[b];
                for (int q = 0; q < t; ++q) o[q] += kb * v[q];
            }
        }
    }
    free(zt);
}

/* float64 arithmetic (the --double path) */
static void rows_f64(const job_t* jb, long r0, long r1) {
    const double *Z1 = jb->Z1, *Z2 = jb->Z2, *c = jb->c, *V = jb->V;
    double* out = jb->out;
    const int J = jb->J, K = jb->K, t = jb->t, JK = J * K;
    const long n = jb->n;
    for (long i = r0; i < r1; ++i) {
        const double* a = Z1 + i * JK;
        double* o = out + i * t;
        for (int q = 0; q < t; ++q) o[q] = 0.0;
        for (long col = 0; col < n; ++col) {
            const double* z = Z2 + col * JK;
            double kv = 0.0;
            for (int j = 0; j < J; ++j) {
                double sq = 0.0;
                for (int mm = 0; mm < K; ++mm) {
                    const double d = a[j * K + mm] - z[j * K + mm];
                    sq += d * d;
                }
                kv += c[j] * exp(-0.5 * sq);
            }
            const double* v = V + col * t;
            for (int q = 0; q < t; ++q) o[q] += kv * v[q];
        }
    }
}

static void* worker(void* arg) {
    const job_t* jb = (const job_t*)arg;
    /* interleaved row blocks: thread k takes blocks k, k+T, k+2T, ... */
    for (long r0 = (long)jb->tid * ROW_BLOCK; r0 < jb->m; r0 += (long)jb->nthreads * ROW_BLOCK) {
        const long r1 = (r0 + ROW_BLOCK < jb->m) ? r0 + ROW_BLOCK : jb->m;
        if (jb->is_f64) rows_f64(jb, r0, r1); else rows_f32(jb, r0, r1);
    }
    return NULL;
}

static int run(job_t base, int threads) {
    if (base.m < 0 || base.n < 0 || base.J < 1 || base.K < 1 || base.t < 1) return 1;
    if (threads <= 0) threads = oracle_max_threads();
    if (threads > MAX_THREADS) threads = MAX_THREADS;
    pthread_t tids[MAX_THREADS];
    job_t jobs[MAX_THREADS];
    for (int k = 0; k < threads; ++k) {
        jobs[k] = base;
        jobs[k].tid = k;
        jobs[k].nthreads = threads;
        if (k > 0 && pthread_create(&tids[k], NULL, worker, &jobs[k]) != 0) return 2;
    }
    worker(&jobs[0]);
    for (int k = 1; k < threads; ++k) pthread_join(tids[k], NULL);
    return 0;
}

int oracle_kmv_f32(const float* Z1, long m, const float* Z2, long n, int J, int K, const float* c, const float* V,
                   int t, float* out, int threads) {
    job_t jb = {Z1, Z2, c, V, out, m, n, J, K, t, 0, 1, 0};
    return run(jb, threads);
}

int oracle_kmv_f64(const double* Z1, long m, const double* Z2, long n, int J, int K, const double* c, const double* V,
                   int t, double* out, int threads) {
    job_t jb = {Z1, Z2, c, V, out, m, n, J, K, t, 0, 1, 1};
    return run(jb, threads);
}
