// Orthogonalisation kernels of the inverse-Krylov (Arnoldi) basis build - SURVEY.md §8f row 3.
//
// The reference builds one orthonormal basis per coupling surface with modified Gram-Schmidt in
// Python (thermca/fem/mor_io.py:169-220): v = A^-1 q_k; v -= (q_j.v) q_j for j <= k; q_{k+1} = v/|v|.
// On the device the projection runs as classical Gram-Schmidt applied twice ("twice is enough": the
// same orthogonality as MGS, but all k dot products of a pass read v once and need one reduction):
//     h = Q_k^T v  (tc_krylov_dots_kernel + ordered reduction)
//     v = v - Q_k h, partial |v|^2  (tc_krylov_update_kernel)
// Q is column-major (column j at Q + j*ld), so every load is coalesced.  All reductions are per-block
// partials summed in block order: results are run-to-run identical.
#include "common.cuh"
#include "../../include/thermca_b200.h"

#define KR_THREADS 256
#define KR_COLS 8            // columns handled per pass over a row chunk

// partial[j*G + blockIdx.x] = sum over the block's rows of Q[i,j]*v[i]
__global__ void __launch_bounds__(KR_THREADS)
tc_krylov_dots_kernel(long long n, int k, const double* __restrict__ Q, long long ld, const double* __restrict__ v,
                      double* __restrict__ partial) {
    __shared__ double sh[32];
    const int G = gridDim.x;
    for (int j0 = 0; j0 < k; j0 += KR_COLS) {
        double acc[KR_COLS];
#pragma unroll
        for (int c = 0; c < KR_COLS; ++c) acc[c] = 0.0;
        for (long long i = blockIdx.x * (long long)KR_THREADS + threadIdx.x; i < n; i += (long long)G * KR_THREADS) {
            const double vi = v[i];
#pragma unroll
            for (int c = 0; c < KR_COLS; ++c)
                if (j0 + c < k) acc[c] += Q[(long long)(j0 + c) * ld + i] * vi;
        }
#pragma unroll
        for (int c = 0; c < KR_COLS; ++c) {
            const double s = block_sum(acc[c], sh);
            if (threadIdx.x == 0 && j0 + c < k) partial[(long long)(j0 + c) * G + blockIdx.x] = s;
        }
    }
}

// h[j] (+)= sum_b partial[j*G + b] in block order; one block per column
__global__ void __launch_bounds__(KR_THREADS)
tc_krylov_reduce_kernel(int G, const double* __restrict__ partial, double* __restrict__ h, int accumulate) {
    __shared__ double sh[32];
    const int j = blockIdx.x;
    const double s = ordered_partial_sum(partial + (long long)j * G, G, sh);
    if (threadIdx.x == 0) h[j] = (accumulate ? h[j] : 0.0) + s;
}

// v -= Q_k h ; partial_norm[blockIdx.x] = sum of v_i^2 over the block's rows
__global__ void __launch_bounds__(KR_THREADS)
tc_krylov_update_kernel(long long n, int k, const double* __restrict__ Q, long long ld, const double* __restrict__ h,
                        double* __restrict__ v, double* __restrict__ partial_norm) {
    __shared__ double sh[32];
    extern __shared__ double hs[];
    for (int j = threadIdx.x; j < k; j += KR_THREADS) hs[j] = h[j];
    __syncthreads();
    double nn = 0.0;
    for (long long i = blockIdx.x * (long long)KR_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * KR_THREADS) {
        double vi = v[i];
        for (int j = 0; j < k; ++j) vi -= hs[j] * Q[(long long)j * ld + i];
        v[i] = vi;
        nn += vi * vi;
    }
    const double s = block_sum(nn, sh);
    if (threadIdx.x == 0) partial_norm[blockIdx.x] = s;
}

// out = v * (1 / sqrt(nrm2[0]))
__global__ void tc_krylov_scale_kernel(long long n, const double* __restrict__ v, const double* __restrict__ nrm2,
                                       double* __restrict__ out) {
    const double s = 1.0 / sqrt(nrm2[0]);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = v[i] * s;
}

extern "C" int64_t tc_krylov_work_doubles(int32_t k_max) {
    return (int64_t)(k_max + 1) * TC_MAX_PARTIALS + 2LL * (k_max + 1) + 8;
}

// Orthogonalise v against the first k columns of Q (`passes` rounds of classical Gram-Schmidt); writes
// |v|^2 after the projection to nrm2_out[0] (device) and the accumulated coefficients to h_out[0..k).
extern "C" int tc_krylov_orthogonalize(void* cuda_stream, int64_t n, int32_t k, const double* Q, int64_t ld, double* v,
                                       int32_t passes, double* work, double* h_out, double* nrm2_out) {
    TC_CHECK(n > 0 && k >= 0 && passes >= 1, "bad arguments");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    long long g = (n + KR_THREADS - 1) / KR_THREADS;
    if (g > 148 * 8) g = 148 * 8;
    if (g > TC_MAX_PARTIALS) g = TC_MAX_PARTIALS;
    const int G = (int)g;
    double* partial = work;                                   // [k][G]
    double* pnorm = work + (long long)(k + 1) * TC_MAX_PARTIALS - TC_MAX_PARTIALS;   // last bank: [G]
    double* h = work + (long long)(k + 1) * TC_MAX_PARTIALS;  // [k]
    for (int p = 0; p < passes; ++p) {
        if (k > 0) {
            tc_krylov_dots_kernel<<<G, KR_THREADS, 0, st>>>(n, k, Q, ld, v, partial);
            tc_krylov_reduce_kernel<<<k, KR_THREADS, 0, st>>>(G, partial, h, 0);
            if (h_out) tc_krylov_reduce_kernel<<<k, KR_THREADS, 0, st>>>(G, partial, h_out, p > 0);
        }
        tc_krylov_update_kernel<<<G, KR_THREADS, sizeof(double) * (k > 0 ? k : 1), st>>>(n, k, Q, ld, h, v, pnorm);
    }
    tc_krylov_reduce_kernel<<<1, KR_THREADS, 0, st>>>(G, pnorm, nrm2_out, 0);
    TC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tc_krylov_normalize(void* cuda_stream, int64_t n, const double* v, const double* nrm2, double* out) {
    long long g = (n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    tc_krylov_scale_kernel<<<(int)g, 256, 0, (cudaStream_t)cuda_stream>>>(n, v, nrm2, out);
    TC_CUDA(cudaGetLastError());
    return 0;
}
