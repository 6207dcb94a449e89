// Shared device helpers for the thermca_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define TC_WARP 32
#define TC_SLICE 32          // SELL-C slice height = one warp
#define TC_MAX_PARTIALS 2048 // upper bound of blocks that publish reduction partials

// ---- error plumbing (host) ---------------------------------------------------------
extern "C" void tc_set_error(const char* file, int line, const char* msg);
#define TC_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) {                                             \
            tc_set_error(__FILE__, __LINE__, cudaGetErrorString(e_));        \
            return -1;                                                       \
        }                                                                    \
    } while (0)
#define TC_CHECK(cond, msg)                                                  \
    do {                                                                     \
        if (!(cond)) {                                                       \
            tc_set_error(__FILE__, __LINE__, msg);                           \
            return -2;                                                       \
        }                                                                    \
    } while (0)

// ---- streaming loads that do not allocate in L1 (matrix values / column indices) ----
__device__ __forceinline__ double ld_stream_f64(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_stream_s16(const short* p) {
    short v;
    asm volatile("ld.global.nc.L1::no_allocate.s16 %0, [%1];" : "=h"(v) : "l"(p));
    return (int)v;
}
__device__ __forceinline__ int ld_stream_s32(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// Matrix streams with an L2 evict-first policy (EF): when the vectors of a rank fit into the 126 MB L2 (strong
// scaling: ~1 M rows per GPU) the operator must not push them out, it is read exactly once per sweep.
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <bool EF> __device__ __forceinline__ double ldm_f64(const double* p, unsigned long long pol) {
    if (!EF) return ld_stream_f64(p);
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
template <bool EF> __device__ __forceinline__ int ldm_s16(const short* p, unsigned long long pol) {
    if (!EF) return ld_stream_s16(p);
    short v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol));
    return (int)v;
}
template <bool EF> __device__ __forceinline__ int ldm_s32(const int* p, unsigned long long pol) {
    if (!EF) return ld_stream_s32(p);
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}

// ---- deterministic block reduction (fixed shuffle tree, fixed warp order) -----------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Sum over the block; result valid in thread 0. `sh` needs >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        r = lane < nw ? sh[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// "Last block done" finaliser: every block stores its partial(s); the block that
// arrives last sums them in block-index order (deterministic) and calls `fin`.
// Returns true in ALL threads of the last block.
__device__ __forceinline__ bool last_block_arrives(unsigned int* counter) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
        if (is_last) *counter = 0u;   // re-arm for the next launch
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}
// Ordered sum of n partials by one block; result valid in thread 0.  The partials were written by other blocks
// and made visible by a grid barrier / last-block fence, so they are read from L2 (ld.global.cg); four loads are
// in flight at a time (a chain of dependent volatile loads cost ~0.6 us each), the additions keep the index order.
__device__ __forceinline__ double ordered_partial_sum(const double* part, int n, double* sh) {
    double v = 0.0;
    for (int i0 = threadIdx.x; i0 < n; i0 += 4 * blockDim.x) {
        double buf[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = i0 + k * blockDim.x;
            buf[k] = i < n ? __ldcg(part + i) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i0 + k * (int)blockDim.x < n) v += buf[k];
    }
    return block_sum(v, sh);
}


// ---- grid size of a persistent kernel whose warps grid-stride over `nslices` work items -------------------
// With W = 32*G/32... warps every warp gets floor or ceil(nslices / W) items; when the remainder is a small fraction
// of W the last round runs on a handful of warps that cannot saturate HBM (measured at N = 2 x 1.25 M rows: 4.12
// slices per warp -> the 12 % of the warps with a fifth slice finish at 44 us, the others at 35 us).  Pick the grid
// (not below 3/4 of the resident maximum) that maximises  fill of the rounds x occupancy factor; the occupancy
// factor 1 - 0.3 (1 - G/Gmax) is fitted to the measured 0.856 (40 warps/SM) vs 0.94 (64 warps/SM) of round 1.
static inline int tc_balanced_grid(long long nslices, int warps_per_block, int gmax) {
    if (gmax < 1) gmax = 1;
    long long need = (nslices + warps_per_block - 1) / warps_per_block;
    if (need < 1) need = 1;
    if (need <= gmax) return (int)need;                       // fewer items than warps: one item per warp
    int best = gmax;
    double best_score = -1.0;
    for (int g = gmax; g >= (3 * gmax) / 4 && g >= 1; --g) {
        const long long w = (long long)g * warps_per_block;
        const long long rounds = (nslices + w - 1) / w;
        const double fill = (double)nslices / (double)(rounds * w);
        const double score = fill * (1.0 - 0.3 * (1.0 - (double)g / gmax));
        if (score > best_score + 1e-12) { best_score = score; best = g; }
    }
    return best;
}
