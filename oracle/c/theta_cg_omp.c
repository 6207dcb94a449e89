/* CPU ORACLE (TEST / BASELINE INFRASTRUCTURE ONLY): multi-threaded port of the implicit
 * theta step with Jacobi-PCG on the reference's operators (A = M^-1 L as FFEIO stores it,
 * thermca/fem/ffe_io.py:85-117; product as scipy csr_matvec does it, thermca/solver.py:281).
 * Stands in for the reference's optional MKL backend (thermca/_utils/sparse.py:6-9), which is
 * not installable here.  Same algorithm as oracle/theta_oracle.py:theta_cg_steps and as the
 * device kernel csrc/pcg.cuh; used only by bench.py's CPU legs and tests.
 *   gcc -O3 -march=native -fopenmp -shared -fPIC theta_cg_omp.c -o ../_build/libtheta_cg_omp.so */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>

static void spmv(int64_t n, const int32_t* ip, const int32_t* ix, const double* a, const double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double s = 0.0;
        for (int32_t k = ip[i]; k < ip[i + 1]; ++k) s += a[k] * x[ix[k]];
        y[i] = s;
    }
}

/* nsteps theta steps of y' = b - A y; returns total PCG iterations. work: 6*n doubles. */
int64_t theta_cg_steps(int64_t n, const int32_t* ip, const int32_t* ix, const double* a, const double* mass,
                       const double* adiag, const double* b, double* y, int32_t nsteps, double h, double theta,
                       double rtol, int32_t maxit, double* work, int32_t threads) {
    if (threads > 0) omp_set_num_threads(threads);
    double *x = work, *r = work + n, *p = work + 2 * n, *q = work + 3 * n, *dinv = work + 4 * n, *t = work + 5 * n;
    const double shift = theta * h;
    int64_t iters = 0;
    for (int32_t st = 0; st < nsteps; ++st) {
        spmv(n, ip, ix, a, y, t);
        double rho = 0.0, bb = 0.0;
#pragma omp parallel for reduction(+ : rho, bb) schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            const double ri = mass[i] * (b[i] - t[i]);
            const double d = 1.0 / (mass[i] * (1.0 + shift * adiag[i]));
            r[i] = ri; x[i] = 0.0; dinv[i] = d; p[i] = d * ri;
            rho += ri * d * ri; bb += ri * ri;
        }
        const double tol2 = rtol * rtol * bb;
        for (int32_t it = 0; it < maxit && bb > 0.0; ++it) {
            spmv(n, ip, ix, a, p, t);
            double pq = 0.0;
#pragma omp parallel for reduction(+ : pq) schedule(static)
            for (int64_t i = 0; i < n; ++i) { q[i] = mass[i] * (p[i] + shift * t[i]); pq += p[i] * q[i]; }
            const double alpha = rho / pq;
            double rn = 0.0, rr = 0.0;
#pragma omp parallel for reduction(+ : rn, rr) schedule(static)
            for (int64_t i = 0; i < n; ++i) {
                x[i] += alpha * p[i];
                const double ri = r[i] - alpha * q[i];
                r[i] = ri; rn += ri * dinv[i] * ri; rr += ri * ri;
            }
            ++iters;
            if (!(rr > tol2)) break;
            const double beta = rn / rho;
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i) p[i] = dinv[i] * r[i] + beta * p[i];
            rho = rn;
        }
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) y[i] += h * x[i];
    }
    return iters;
}
