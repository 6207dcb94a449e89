// Micro-benchmark of the reduce-barrier (gridbar.cuh) and of cooperative_groups' grid.sync():
// an empty persistent kernel that does nothing but `iters` barriers.  Used to size the grid of the
// PCG kernels (blocks per SM x threads) — profiles/README.md keeps the numbers.
#include <cooperative_groups.h>
#include "../../include/thermca_b200.h"
#include "gridbar.cuh"
namespace cg = cooperative_groups;

template <int K, int MODE>
__global__ void tc_bar_bench_kernel(TcGridBar B, int iters, double* sink) {
    unsigned int gen = 0u, tgen = tc_bar_generation(B);
    double v[K > 0 ? K : 1];
    double acc = 0.0;
    cg::grid_group grid = cg::this_grid();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 1) { grid.sync(); continue; }
#pragma unroll
        for (int k = 0; k < (K > 0 ? K : 1); ++k) v[k] = 1.0 + k;
        if (MODE == 2) tc_grid_reduce_last<(K > 0 ? K : 1)>(B, tgen, v, [](double*, int) {});
        else tc_grid_reduce<K>(B, gen, v);
        acc += v[0];
    }
    if (sink && blockIdx.x == 0 && threadIdx.x == 0) sink[0] = acc;
}

// ns per barrier; mode 0 = reduce-barrier with K values (0, 1, 2), 1 = cooperative_groups grid.sync(),
// 2 = last-arriver variant (K = 1, 2) with an empty exchange hook
extern "C" int tc_bar_bench(void* cuda_stream, int32_t blocks, int32_t threads, int32_t iters, int32_t K, int32_t mode,
                            double* ns_per_barrier) {
    cudaStream_t st = (cudaStream_t)cuda_stream;
    TC_CHECK(blocks > 0 && blocks <= TC_MAX_PARTIALS && threads >= 32 && threads <= 1024, "grid");
    TcGridBar B{};
    TC_CUDA(cudaMalloc(&B.ctr, sizeof(unsigned int) * TC_BAR_CTR_WORDS));
    TC_CUDA(cudaMemset(B.ctr, 0, sizeof(unsigned int) * TC_BAR_CTR_WORDS));
    TC_CUDA(cudaMalloc(&B.part, sizeof(double) * TC_BAR_PART_DOUBLES));
    TC_CUDA(cudaMalloc(&B.res, sizeof(double) * 2 * TC_BAR_MAXK + 8));
    double* sink = B.res + 2 * TC_BAR_MAXK;
    void* fn = mode == 1 ? (void*)tc_bar_bench_kernel<0, 1>
               : mode == 2 ? (K <= 1 ? (void*)tc_bar_bench_kernel<1, 2> : (void*)tc_bar_bench_kernel<2, 2>)
               : K == 0 ? (void*)tc_bar_bench_kernel<0, 0> : K == 1 ? (void*)tc_bar_bench_kernel<1, 0>
                                                                    : (void*)tc_bar_bench_kernel<2, 0>;
    cudaEvent_t e0, e1;
    TC_CUDA(cudaEventCreate(&e0)); TC_CUDA(cudaEventCreate(&e1));
    int rc = 0;
    for (int rep = 0; rep < 2 && rc == 0; ++rep) {          // first launch warms up
        void* args[] = {(void*)&B, (void*)&iters, (void*)&sink};
        cudaEventRecord(e0, st);
        cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(threads), args, 0, st);
        cudaEventRecord(e1, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { tc_set_error(__FILE__, __LINE__, cudaGetErrorString(e)); rc = -1; }
    }
    float ms = 0.f;
    if (rc == 0) { cudaEventElapsedTime(&ms, e0, e1); *ns_per_barrier = 1e6 * ms / iters; }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(B.ctr); cudaFree(B.part); cudaFree(B.res);
    return rc;
}
