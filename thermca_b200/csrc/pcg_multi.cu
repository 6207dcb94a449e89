// Jacobi-preconditioned CG for K right-hand sides at once on one SELL-32 operator (K <= 8).
//
// Used by the MOR basis build (thermca_b200/mor_basis.py): the inverse-Krylov Arnoldi processes of the S
// coupling surfaces of a body (thermca/fem/mor_io.py:169-220) all solve with the same symmetrised state
// matrix, so their k-th solves run together and the matrix - 12 of the ~25 bytes per non-zero and iteration -
// is streamed once per iteration instead of S times.  The K recurrences are independent (this is not block
// CG): every right-hand side has its own alpha / beta / convergence test and is frozen once converged.
// Vectors are interleaved, v[row*K + k], so the gather of a neighbour row fetches its K values from one or two
// sectors.  ONE cooperative persistent launch per solve; dot products are per-block partials summed in block
// order by every block (same scheme as csrc/pcg.cuh): run-to-run identical, no atomics.
#include <cooperative_groups.h>
#include "common.cuh"
#include "../../include/thermca_b200.h"
namespace cg = cooperative_groups;

#define PM_THREADS 256
#define PM_KMAX 8

struct TcPcgMulti {
    int n, nslices, K;
    const long long* slice_ptr;
    const int* col;
    const double* val;
    const double* diag;
    const double* b;        // [n][K]
    double* x;              // [n][K]
    double *r, *p, *q;      // [n][K]
    double* partial;        // [3][PM_KMAX][TC_MAX_PARTIALS]
    int* iters;             // [K] iterations until convergence (maxit if not converged)
    double rtol;
    int maxit;
};

// every block sums the G partials of K quantities in block order; result in out[k] for all threads
template <int KT>
__device__ __forceinline__ void pm_bank_sums(const double* bank, int G, int K, double* sh, double* bc, double* out) {
    for (int k = 0; k < K; ++k) {
        const double s = ordered_partial_sum(bank + (size_t)k * TC_MAX_PARTIALS, G, sh);
        if (threadIdx.x == 0) bc[k] = s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KT; ++k) out[k] = bc[k];      // static indices: `out` stays in registers
    __syncthreads();
}

// N stored entries of one row against the KT interleaved vectors: N value + N column loads in flight, then the gathers
#define PM_GROUP 4
template <int N, int KT>
__device__ __forceinline__ void pm_row_group(const TcPcgMulti& S, long long pos, double* sum) {
    double v[N];
    long long c[N];
#pragma unroll
    for (int t = 0; t < N; ++t) v[t] = ld_stream_f64(S.val + pos + 32LL * t);
#pragma unroll
    for (int t = 0; t < N; ++t) c[t] = ld_stream_s32(S.col + pos + 32LL * t);
#pragma unroll
    for (int t = 0; t < N; ++t) {
#pragma unroll
        for (int k = 0; k < KT; ++k) sum[k] += v[t] * S.p[c[t] * KT + k];
    }
}

// KT = number of right-hand sides (compile time: exact unrolling, registers and occupancy scale with it)
template <int KT>
__global__ void __launch_bounds__(PM_THREADS, (KT <= 2 ? 6 : (KT <= 4 ? 4 : 2)))
tc_pcg_multi_kernel(TcPcgMulti S) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[32];
    __shared__ double bc[KT];
    const int G = gridDim.x, n = S.n;
    constexpr int K = KT;
    const int tid = threadIdx.x, lane = tid & 31;
    const long long gtid = blockIdx.x * (long long)PM_THREADS + tid, gthreads = (long long)G * PM_THREADS;
    double* bank0 = S.partial;
    double* bank1 = S.partial + (size_t)PM_KMAX * TC_MAX_PARTIALS;
    double* bank2 = S.partial + 2 * (size_t)PM_KMAX * TC_MAX_PARTIALS;

    // init: x = 0, r = b, p = r / diag; rho = r.z, bb = b.b
    double acc0[KT], acc1[KT];
#pragma unroll
    for (int k = 0; k < KT; ++k) acc0[k] = acc1[k] = 0.0;
    for (long long i = gtid; i < n; i += gthreads) {
        const double d = 1.0 / S.diag[i];
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if (k < K) {
                const double bv = S.b[i * K + k];
                S.x[i * K + k] = 0.0; S.r[i * K + k] = bv; S.p[i * K + k] = d * bv;
                acc0[k] += bv * d * bv; acc1[k] += bv * bv;
            }
    }
#pragma unroll
    for (int k = 0; k < KT; ++k)
        if (k < K) {                                   // uniform over the grid
            const double s0 = block_sum(acc0[k], sh);
            const double s1 = block_sum(acc1[k], sh);
            if (tid == 0) { bank0[(size_t)k * TC_MAX_PARTIALS + blockIdx.x] = s0; bank1[(size_t)k * TC_MAX_PARTIALS + blockIdx.x] = s1; }
        }
    grid.sync();
    double rho[KT], tol2[KT];
    bool active[KT];
    pm_bank_sums<KT>(bank0, G, K, sh, bc, rho);
    pm_bank_sums<KT>(bank1, G, K, sh, bc, tol2);
    int n_active = 0;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        if (k < K) { active[k] = tol2[k] > 0.0; tol2[k] *= S.rtol * S.rtol; n_active += active[k]; }
        else active[k] = false;
    }
    if (blockIdx.x == 0 && tid < K) S.iters[tid] = 0;
    const int warps = PM_THREADS / 32;
    int it = 0;
    while (n_active > 0 && it < S.maxit) {
        // q = A p (one warp per 32-row slice), partial p.q
#pragma unroll
        for (int k = 0; k < KT; ++k) acc0[k] = 0.0;
        for (int sl = blockIdx.x * warps + (tid >> 5); sl < S.nslices; sl += G * warps) {
            const long long base = S.slice_ptr[sl];
            const int width = (int)((S.slice_ptr[sl + 1] - base) >> 5);
            const long long row = (long long)sl * 32 + lane;
            double sum[KT];
#pragma unroll
            for (int k = 0; k < KT; ++k) sum[k] = 0.0;
            // values and columns of a group of entries travel together (one round trip per group instead of one per
            // entry: the warps are latency bound, parts.cuh); the tail is one group behind a warp-uniform branch;
            // the sums keep the stored entry order
            int j = 0;
            for (; j + PM_GROUP <= width; j += PM_GROUP)
                pm_row_group<PM_GROUP, KT>(S, base + 32LL * j + lane, sum);
            switch (width - j) {                                   // warp uniform (all lanes share the width)
                case 3: pm_row_group<3, KT>(S, base + 32LL * j + lane, sum); break;
                case 2: pm_row_group<2, KT>(S, base + 32LL * j + lane, sum); break;
                case 1: pm_row_group<1, KT>(S, base + 32LL * j + lane, sum); break;
                default: break;
            }
            if (row < n) {
#pragma unroll
                for (int k = 0; k < KT; ++k)
                    if (k < K) { S.q[row * K + k] = sum[k]; acc0[k] += S.p[row * K + k] * sum[k]; }
            }
        }
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if (k < K) {
                const double s0 = block_sum(acc0[k], sh);
                if (tid == 0) bank0[(size_t)k * TC_MAX_PARTIALS + blockIdx.x] = s0;
            }
        grid.sync();
        double alpha[KT];
        pm_bank_sums<KT>(bank0, G, K, sh, bc, alpha);
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            alpha[k] = (k < K && active[k]) ? rho[k] / alpha[k] : 0.0;
            acc0[k] = acc1[k] = 0.0;
        }
        // x += alpha p ; r -= alpha q ; partial r.(r/diag), r.r
        for (long long i = gtid; i < n; i += gthreads) {
            const double d = 1.0 / S.diag[i];
#pragma unroll
            for (int k = 0; k < KT; ++k)
                if (k < K) {
                    const size_t e = (size_t)i * K + k;
                    S.x[e] += alpha[k] * S.p[e];
                    const double ri = S.r[e] - alpha[k] * S.q[e];
                    S.r[e] = ri;
                    acc0[k] += ri * d * ri; acc1[k] += ri * ri;
                }
        }
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if (k < K) {
                const double s0 = block_sum(acc0[k], sh);
                const double s1 = block_sum(acc1[k], sh);
                if (tid == 0) { bank1[(size_t)k * TC_MAX_PARTIALS + blockIdx.x] = s0; bank2[(size_t)k * TC_MAX_PARTIALS + blockIdx.x] = s1; }
            }
        grid.sync();
        double rho_new[KT], rr[KT], beta[KT];
        pm_bank_sums<KT>(bank1, G, K, sh, bc, rho_new);
        pm_bank_sums<KT>(bank2, G, K, sh, bc, rr);
        ++it;
        n_active = 0;
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            beta[k] = 0.0;
            if (k < K && active[k]) {
                if (!(rr[k] > tol2[k])) {
                    active[k] = false;
                    if (blockIdx.x == 0 && tid == 0) S.iters[k] = it;
                } else {
                    beta[k] = rho_new[k] / rho[k];
                    rho[k] = rho_new[k];
                    ++n_active;
                }
            }
        }
        if (n_active == 0) break;                      // uniform: every block holds the same sums
        for (long long i = gtid; i < n; i += gthreads) {
            const double d = 1.0 / S.diag[i];
#pragma unroll
            for (int k = 0; k < KT; ++k)
                if (k < K && active[k]) {
                    const size_t e = (size_t)i * K + k;
                    S.p[e] = d * S.r[e] + beta[k] * S.p[e];
                }
        }
        grid.sync();
    }
    if (blockIdx.x == 0 && tid == 0) {
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if (k < K && active[k]) S.iters[k] = S.maxit;
    }
}

extern "C" int64_t tc_pcg_multi_work_doubles(int64_t n, int32_t K) {
    return 3LL * n * K + 3LL * PM_KMAX * TC_MAX_PARTIALS;
}

// Solves A X = B for K right-hand sides; B and X are [n][K] (K values of a row contiguous).  A: SELL-32,
// symmetric positive definite, diag = its diagonal.  iters_dev[K] (device) receives the iteration counts.
extern "C" int tc_pcg_solve_multi(void* cuda_stream, int32_t n, int32_t nslices, const int64_t* slice_ptr,
                                  const int32_t* col, const double* val, const double* diag, int32_t K,
                                  const double* B, double* X, double* work, double rtol, int32_t maxit,
                                  int32_t* iters_dev) {
    TC_CHECK(K >= 1 && K <= PM_KMAX, "between 1 and 8 right-hand sides");
    TC_CHECK(n > 0 && nslices == (n + 31) / 32, "slice count does not match the row count");
    TcPcgMulti S{};
    S.n = n; S.nslices = nslices; S.K = K;
    S.slice_ptr = (const long long*)slice_ptr; S.col = col; S.val = val; S.diag = diag;
    S.b = B; S.x = X;
    S.r = work; S.p = work + (size_t)n * K; S.q = work + 2 * (size_t)n * K;
    S.partial = work + 3 * (size_t)n * K;
    S.iters = iters_dev; S.rtol = rtol;
    S.maxit = maxit;
    if ((long long)S.maxit > 2LL * n + 50) S.maxit = 2 * n + 50;
    int dev = 0, sms = 148, per_sm = 0;
    TC_CUDA(cudaGetDevice(&dev));
    TC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const void* fn = nullptr;
    switch (K) {
        case 1: fn = (const void*)tc_pcg_multi_kernel<1>; break;
        case 2: fn = (const void*)tc_pcg_multi_kernel<2>; break;
        case 3: fn = (const void*)tc_pcg_multi_kernel<3>; break;
        case 4: fn = (const void*)tc_pcg_multi_kernel<4>; break;
        case 5: fn = (const void*)tc_pcg_multi_kernel<5>; break;
        case 6: fn = (const void*)tc_pcg_multi_kernel<6>; break;
        case 7: fn = (const void*)tc_pcg_multi_kernel<7>; break;
        default: fn = (const void*)tc_pcg_multi_kernel<8>; break;
    }
    TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, PM_THREADS, 0));
    TC_CHECK(per_sm > 0, "kernel does not fit an SM");
    int g = sms * per_sm;
    const int need = (nslices + PM_THREADS / 32 - 1) / (PM_THREADS / 32);
    if (need < g) g = need;
    if (g > TC_MAX_PARTIALS) g = TC_MAX_PARTIALS;
    if (g < 1) g = 1;
    void* args[] = {(void*)&S};
    TC_CUDA(cudaLaunchCooperativeKernel(fn, dim3(g), dim3(PM_THREADS), args, 0, (cudaStream_t)cuda_stream));
    return 0;
}
