// Jacobi-preconditioned conjugate gradients for the implicit step of one FE part.
//
// System: mass * (I + shift * A) x = mass * rhs, i.e. (M + shift*K) x = M rhs with
// A = M^-1 K as the reference stores it (thermca/fem/ffe_io.py:85-117); M + shift*K is
// symmetric positive definite, diag = mass_i * (1 + shift * A_ii).
// The reference has no implicit scheme: this is a new method on the
// reference's operators; its CPU restatement is oracle/theta_oracle.py.
//
// Two drivers over the same phases:
//   * tc_pcg_coop_kernel  - ONE cooperative persistent launch per solve, reduce-barriers
//     (gridbar.cuh) between phases, convergence loop on the device (no host round trip);
//   * tc_pcg_k1/k2/k3     - one launch per phase, used when cooperative launch is not
//     available and for cross-checking.
// Dot products: per-block partial (fixed shuffle tree) -> the last block to reach the barrier sums
// the G partials in index order and publishes the total, so all blocks see the same scalar and
// results are run-to-run identical.
#pragma once
#include <cooperative_groups.h>
#include "parts.cuh"
#include "gridbar.cuh"
namespace cg = cooperative_groups;

struct TcPcg {
    TcSell A;
    const double* mass;
    const double* adiag;
    const double* rhs;      // unscaled right-hand side
    double* x;              // solution
    double *r, *p, *q, *dinv;
    double* partial;        // 4 banks of TC_MAX_PARTIALS
    double* sc;             // [0],[1] rho (parity), [2] tol2, [3] bnorm2, [4] rr
    int* isc;               // [0] pcg_done, [1] iterations of this solve, [2] total iterations, [3] solves,
                            // [4] status of the last failed solve (1 iteration limit, 2 breakdown), [5] failed solves
    double rtol;
    int maxit;
    const double* ctrl;     // shift = coef * ctrl[1]
    double coef;
    unsigned long long* prof;   // optional [16] phase times (ns) of block 0
    const int* skip;            // optional: non-zero -> leave x untouched (block is explicit this attempt)
    TcGridBar bar;              // reduce-barrier state of the cooperative kernel (gridbar.cuh)
};

#define TC_PCG_THREADS 256

__device__ __forceinline__ double tc_bank_sum(const double* bank, int G, double* sh) {
    __shared__ double bcast;
    const double s = ordered_partial_sum(bank, G, sh);
    __syncthreads();
    if (threadIdx.x == 0) bcast = s;
    __syncthreads();
    return bcast;
}

__device__ __forceinline__ void tc_pcg_init_phase(const TcPcg& S, double shift, double* sh) {
    const int n = S.A.n;
    double rho = 0.0, bb = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double m = S.mass[i];
        const double b = m * S.rhs[i];
        const double d = 1.0 / (m * (1.0 + shift * S.adiag[i]));
        const double z = d * b;
        S.r[i] = b; S.x[i] = 0.0; S.dinv[i] = d; S.p[i] = z;
        rho += b * z; bb += b * b;
    }
    const double r0 = block_sum(rho, sh);
    const double r1 = block_sum(bb, sh);
    if (threadIdx.x == 0) { S.partial[blockIdx.x] = r0; S.partial[TC_MAX_PARTIALS + blockIdx.x] = r1; }
}

// x += alpha p ; r -= alpha q ; partial sums of r.(dinv r) and r.r
__device__ __forceinline__ void tc_pcg_update_phase(const TcPcg& S, double alpha, double* sh) {
    const int n = S.A.n;
    double rho = 0.0, rr = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double pi = S.p[i];
        S.x[i] += alpha * pi;
        const double ri = S.r[i] - alpha * S.q[i];
        S.r[i] = ri;
        rho += ri * (S.dinv[i] * ri);
        rr += ri * ri;
    }
    const double r0 = block_sum(rho, sh);
    const double r1 = block_sum(rr, sh);
    if (threadIdx.x == 0) {
        S.partial[2 * TC_MAX_PARTIALS + blockIdx.x] = r0;
        S.partial[3 * TC_MAX_PARTIALS + blockIdx.x] = r1;
    }
}

__device__ __forceinline__ void tc_pcg_dir_phase(const TcPcg& S, double beta) {
    const int n = S.A.n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        S.p[i] = S.dinv[i] * S.r[i] + beta * S.p[i];
}

__device__ __forceinline__ unsigned long long tc_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// accumulated in shared memory, flushed once per launch: a global read-modify-write per phase is an exposed memory
// round trip of thread 0 in front of every barrier
#define PCG_PROF(k) do { if (prof_on) { const unsigned long long t_ = tc_gtime(); s_prof[(k)] += t_ - prof_t; prof_t = t_; } } while (0)

// MB = resident blocks per SM the register allocation is tuned for (occupancy matters: the kernel
// is latency/bandwidth bound; measured on B200, see profiles/).
// Every reduction is folded into its grid barrier (gridbar.cuh): three barriers per iteration, the first two
// carry p.q and (r.z, r.r).  Summation order is fixed (block partial tree, then block-index order), so all
// blocks hold bit-identical scalars and runs are reproducible.
// Status (isc[4]): 0 converged, 1 iteration limit reached with r.r above the tolerance, 2 breakdown
// (non-finite or non-positive p.q / r.r) — surfaced by tc_pcg_stats, checked by the host drivers.
template <int MB, int THREADS = TC_PCG_THREADS>
__global__ void __launch_bounds__(THREADS, MB)
tc_pcg_coop_kernel(TcPcg S, const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;       // uniform over the grid
    if (S.skip && *S.skip) return;             // uniform over the grid
    __shared__ double sh[32];
    unsigned int gen = 0u;
    const double shift = S.coef * S.ctrl[1];
    __shared__ unsigned long long s_prof[8];
    const bool prof_on = S.prof && blockIdx.x == 0 && threadIdx.x == 0;
    if (prof_on) { for (int k = 0; k < 8; ++k) s_prof[k] = 0ULL; }
    unsigned long long prof_t = prof_on ? tc_gtime() : 0ULL;
    double red[2];
    {
        const int n = S.A.n;
        double rho = 0.0, bb = 0.0;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const double m = S.mass[i];
            const double b = m * S.rhs[i];
            const double d = 1.0 / (m * (1.0 + shift * S.adiag[i]));
            const double z = d * b;
            S.r[i] = b; S.x[i] = 0.0; S.dinv[i] = d; S.p[i] = z;
            rho += b * z; bb += b * b;
        }
        red[0] = block_sum(rho, sh);
        red[1] = block_sum(bb, sh);
    }
    tc_grid_reduce<2>(S.bar, gen, red);
    double rho = red[0];
    const double bb = red[1];
    const double tol2 = S.rtol * S.rtol * bb;
    int it = 0, status = 0;
    if (bb > 0.0) {
        status = 1;
        for (; it < S.maxit;) {
            PCG_PROF(0);
            const double dot = tc_sell_spmv_body<SPMV_SHIFT, 0>(S.A, S.p, S.q, nullptr, S.mass, shift);
            PCG_PROF(1);
            red[0] = block_sum(dot, sh);
            tc_grid_reduce<1>(S.bar, gen, red);
            PCG_PROF(2);
            const double pq = red[0];
            if (!(pq > 0.0) || !isfinite(pq)) { status = 2; break; }     // uniform: every block holds the same pq
            const double alpha = rho / pq;
            {
                const int n = S.A.n;
                double rn = 0.0, rr = 0.0;
                // x += alpha p is taken in the direction phase below, which reads p anyway (one stream less per
                // iteration; same operations per element, so x is bit-identical)
                for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
                    const double ri = S.r[i] - alpha * S.q[i];
                    S.r[i] = ri;
                    rn += ri * (S.dinv[i] * ri);
                    rr += ri * ri;
                }
                red[0] = block_sum(rn, sh);
                red[1] = block_sum(rr, sh);
            }
            PCG_PROF(3);
            tc_grid_reduce<2>(S.bar, gen, red);
            PCG_PROF(4);
            const double rho_new = red[0], rr = red[1];
            ++it;
            if (!(rr > tol2)) {                                           // uniform
                status = isfinite(rr) ? 0 : 2;
                const int n = S.A.n;                                      // the last x update
                for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
                    S.x[i] += alpha * S.p[i];
                break;
            }
            {
                const double beta = rho_new / rho;
                const int n = S.A.n;
                for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
                    const double pi = S.p[i];
                    S.x[i] += alpha * pi;
                    S.p[i] = S.dinv[i] * S.r[i] + beta * pi;
                }
            }
            rho = rho_new;
            PCG_PROF(5);
            tc_grid_barrier(S.bar, gen);
            PCG_PROF(6);
        }
    }
    if (prof_on) { for (int k = 0; k < 8; ++k) S.prof[k] += s_prof[k]; }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S.isc[1] = it; S.isc[2] += it; S.isc[3] += 1;
        if (status) { S.isc[4] = status; S.isc[5] += 1; }
    }
}

// ---- block-Jacobi variant (north_star (b): "Jacobi / block-Jacobi preconditioned CG") ---------------------------
// Preconditioner: the 32 x 32 diagonal blocks of M + shift*K that belong to the SELL slices (one warp = one block),
// inverted once per (h, films) on the host side and stored symmetric in FP32, [slice][j][i]: a preconditioner does not
// need FP64 (the residual and all recurrences stay FP64) and the application is bandwidth: 4 KB per slice = 128 B/row
// on top of the ~266 B/row of a point-Jacobi iteration.  Application: lane i holds r_i, z_i = sum_j Binv[j][i] r_j with
// r_j broadcast by shuffle.  The vector phases therefore run warp-per-slice; z is stored (the direction phase reuses it).
// Measured (profiles/README.md): fewer iterations (about -25..-35 %) but +48 % bytes per iteration: it only pays for
// stiff shifts (large h, steady state); point Jacobi stays the default.
__device__ __forceinline__ double tc_block_apply(const float* __restrict__ binv, int slice, int lane, double rv) {
    const float* B = binv + (size_t)slice * 1024 + lane;
    double acc = 0.0;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) acc += (double)__ldg(B + 32 * j) * __shfl_sync(0xffffffffu, rv, j);
    return acc;
}

template <int MB, int THREADS = TC_PCG_THREADS>
__global__ void __launch_bounds__(THREADS, MB)
tc_pcg_coop_bj_kernel(TcPcg S, const float* __restrict__ binv, double* __restrict__ z, const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    if (S.skip && *S.skip) return;
    __shared__ double sh[32];
    unsigned int gen = 0u;
    const double shift = S.coef * S.ctrl[1];
    const int n = S.A.n, nsl = S.A.nslices;
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = gridDim.x * (blockDim.x >> 5);
    double red[2];
    {
        double rho = 0.0, bb = 0.0;
        for (int sl = warp; sl < nsl; sl += nwarps) {
            const int i = sl * 32 + lane;
            const bool live = i < n;
            const double b = live ? S.mass[i] * S.rhs[i] : 0.0;
            const double zi = tc_block_apply(binv, sl, lane, b);
            if (live) { S.r[i] = b; S.x[i] = 0.0; z[i] = zi; S.p[i] = zi; rho += b * zi; bb += b * b; }
        }
        red[0] = block_sum(rho, sh);
        red[1] = block_sum(bb, sh);
    }
    tc_grid_reduce<2>(S.bar, gen, red);
    double rho = red[0];
    const double bb = red[1];
    const double tol2 = S.rtol * S.rtol * bb;
    int it = 0, status = 0;
    if (bb > 0.0) {
        status = 1;
        for (; it < S.maxit;) {
            const double dot = tc_sell_spmv_body<SPMV_SHIFT, 0>(S.A, S.p, S.q, nullptr, S.mass, shift);
            red[0] = block_sum(dot, sh);
            tc_grid_reduce<1>(S.bar, gen, red);
            const double pq = red[0];
            if (!(pq > 0.0) || !isfinite(pq)) { status = 2; break; }
            const double alpha = rho / pq;
            {
                double rn = 0.0, rr = 0.0;
                for (int sl = warp; sl < nsl; sl += nwarps) {
                    const int i = sl * 32 + lane;
                    const bool live = i < n;
                    double ri = 0.0;
                    if (live) {
                        S.x[i] += alpha * S.p[i];
                        ri = S.r[i] - alpha * S.q[i];
                        S.r[i] = ri;
                    }
                    const double zi = tc_block_apply(binv, sl, lane, ri);
                    if (live) { z[i] = zi; rn += ri * zi; rr += ri * ri; }
                }
                red[0] = block_sum(rn, sh);
                red[1] = block_sum(rr, sh);
            }
            tc_grid_reduce<2>(S.bar, gen, red);
            const double rho_new = red[0], rr = red[1];
            ++it;
            if (!(rr > tol2)) { status = isfinite(rr) ? 0 : 2; break; }
            const double beta = rho_new / rho;
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
                S.p[i] = z[i] + beta * S.p[i];
            rho = rho_new;
            tc_grid_barrier(S.bar, gen);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S.isc[1] = it; S.isc[2] += it; S.isc[3] += 1;
        if (status) { S.isc[4] = status; S.isc[5] += 1; }
    }
}

// ---- one-launch-per-phase variant ------------------------------------------------------
__global__ void __launch_bounds__(TC_PCG_THREADS)
tc_pcg_init_kernel(TcPcg S, const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    __shared__ double sh[32];
    tc_pcg_init_phase(S, S.coef * S.ctrl[1], sh);
    if (blockIdx.x == 0 && threadIdx.x == 0) { S.isc[0] = 0; S.isc[1] = 0; S.isc[3] += 1; }
}
// k1: (first iteration only: finish init scalars) q = P p, partial p.q
__global__ void __launch_bounds__(TC_PCG_THREADS)
tc_pcg_k1(TcPcg S, int it, const int* __restrict__ done_flag) {
    if ((done_flag && *done_flag) || S.isc[0]) return;
    __shared__ double sh[32];
    const double dot = tc_sell_spmv_body<SPMV_SHIFT>(S.A, S.p, S.q, nullptr, S.mass, S.coef * S.ctrl[1]);
    const double bs = block_sum(dot, sh);
    if (threadIdx.x == 0) S.partial[blockIdx.x] = bs;
    (void)it;
}
// pgrid = grid size of the kernel that produced the bank being summed
__global__ void __launch_bounds__(TC_PCG_THREADS)
tc_pcg_k2(TcPcg S, int it, int spmv_grid, const int* __restrict__ done_flag) {
    if ((done_flag && *done_flag) || S.isc[0]) return;
    __shared__ double sh[32];
    const double rho = S.sc[it & 1];
    const double pq = tc_bank_sum(S.partial, spmv_grid, sh);
    tc_pcg_update_phase(S, rho / pq, sh);
}
__global__ void __launch_bounds__(TC_PCG_THREADS)
tc_pcg_k3(TcPcg S, int it, int upd_grid, const int* __restrict__ done_flag) {
    if ((done_flag && *done_flag) || S.isc[0]) return;
    __shared__ double sh[32];
    const double rho_new = tc_bank_sum(S.partial + 2 * TC_MAX_PARTIALS, upd_grid, sh);
    const double rr = tc_bank_sum(S.partial + 3 * TC_MAX_PARTIALS, upd_grid, sh);
    const double rho = S.sc[it & 1];
    const bool conv = !(rr > S.sc[2]);
    if (!conv) tc_pcg_dir_phase(S, rho_new / rho);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S.sc[(it + 1) & 1] = rho_new;
        S.sc[4] = rr;
        S.isc[1] = it + 1; S.isc[2] += 1;
    }
    // the done flag must only flip after every block has read isc[0]==0 for this launch:
    // it is raised by the NEXT launch's prologue instead (tc_pcg_flag_kernel).
    if (blockIdx.x == 0 && threadIdx.x == 0 && conv) S.sc[5] = 1.0;
}
// tiny: stage init scalars / raise the done flag between phases
__global__ void tc_pcg_scalar_kernel(TcPcg S, int what, int grid_prev, const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    __shared__ double sh[32];
    if (what == 0) {                 // after init: rho, tol2
        const double rho = tc_bank_sum(S.partial, grid_prev, sh);
        const double bb = tc_bank_sum(S.partial + TC_MAX_PARTIALS, grid_prev, sh);
        if (threadIdx.x == 0) {
            S.sc[0] = rho; S.sc[2] = S.rtol * S.rtol * bb; S.sc[3] = bb; S.sc[5] = 0.0;
            if (!(bb > 0.0)) S.isc[0] = 1;
        }
    } else {                         // after k3: publish convergence
        if (threadIdx.x == 0 && S.sc[5] != 0.0) S.isc[0] = 1;
    }
}
