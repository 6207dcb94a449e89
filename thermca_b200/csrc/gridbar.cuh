// Grid-wide barrier with the reduction folded in ("reduce-barrier") for the persistent PCG kernels.
//
// cooperative_groups::grid.sync() followed by "every block sums all G partials" costs G^2 L2 loads
// and two dependent round trips per reduction (measured round 1: 12.6 % of a 1 M-DOF PCG iteration,
// profiles/README.md).  Here the LAST block to arrive sums the G partials once, in block-index order
// with a fixed tree (bit-reproducible, independent of which block is last), publishes the totals in a
// parity slot and releases the generation flag every other block spins on:
//   arrive : partial[k][block] (st) ; fence ; atomicAdd(count)          (one L2 atomic per block)
//   last   : ordered sum of G partials ; [exchange hook: other GPUs] ; result[parity][k] ; count = 0 ;
//            st.release.gpu generation + 1
//   others : ld.acquire.gpu generation until it moves ; fence (drops stale L1 lines) ; read result
// The launch must be cooperative (all blocks co-resident).  K = 0 gives a plain barrier.
// The exchange hook runs in the last block only, between the local sum and the release: the row-
// partitioned kernel (dist.cu) adds the peers' values there, so one barrier == one global all-reduce.
#pragma once
#include "common.cuh"

#define TC_BAR_MAXK 4

struct TcGridBar {
    unsigned int* ctr;      // [0] arrival count; generation flags at [32 * (1 + j)], j < TC_BAR_FLAGS (one 128-byte line
                            // each: the pollers are spread over several L2 lines); zero-initialised, TC_BAR_CTR_WORDS words
    double* part;           // TC_BAR_MAXK banks of TC_MAX_PARTIALS
    double* res;            // [2][TC_BAR_MAXK] totals, parity = generation & 1
};
#define TC_BAR_FLAGS 32                                   // blocks poll flag (blockIdx % TC_BAR_FLAGS)
#define TC_BAR_CTR_WORDS (32 * (1 + TC_BAR_FLAGS))

// Fences: fence.acq_rel.gpu (MEMBAR.ALL.GPU + L1 invalidate) is enough on both sides of the flag; CUDA's
// __threadfence() is the sequentially consistent fence.sc.gpu (measured ~2x the cost).  The wait loop polls with a
// relaxed (volatile) load: an acquire load per poll would invalidate the SM's L1 on every iteration and evict the
// gather lines of co-resident blocks that are still working.
__device__ __forceinline__ void tc_fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned int tc_ld_relaxed_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void tc_st_relaxed_gpu_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// release-only forms: MEMBAR.ALL.GPU + the access, WITHOUT the L1 invalidation an acq_rel fence carries (checked in
// SASS) — the arriving block must not wipe the L1 lines of co-resident blocks that are still in the phase
__device__ __forceinline__ unsigned int tc_atom_add_release_gpu(unsigned int* p, unsigned int v) {
    unsigned int old;
    asm volatile("atom.release.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void tc_st_release_gpu_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int tc_bar_generation(const TcGridBar& B) {
    return tc_ld_relaxed_gpu_u32(B.ctr + 32);
}

struct TcNoExchange {
    __device__ __forceinline__ void operator()(double*, int) const {}
};

// v[k]: this block's contribution, valid in thread 0.  On return every thread of every block holds the
// grid totals in v[k].  `gen` is the caller's running generation counter (same value in all blocks; the
// kernel reads it once with tc_bar_generation at entry).  sh: >= 32 doubles of shared scratch.
// xchg(tot, K): called by ALL threads of the last block with the local totals in shared memory `tot`;
// may replace them (multi-GPU sum).  Must contain its own __syncthreads() pairs if it uses several threads.
template <int K, class X>
__device__ __forceinline__ void tc_grid_reduce(const TcGridBar& B, unsigned int& gen, double* v, double* sh, X xchg) {
    __shared__ int s_last;
    __shared__ double s_tot[TC_BAR_MAXK > 0 ? TC_BAR_MAXK : 1];
    const int G = gridDim.x;
    __syncthreads();                                       // every thread's earlier stores are ordered before thread 0's fence
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) B.part[k * TC_MAX_PARTIALS + blockIdx.x] = v[k];
        // release: partials + everything this block wrote before (cumulative over the __syncthreads above)
        const unsigned int old = tc_atom_add_release_gpu(B.ctr, 1u);
        const int last = (old == (unsigned int)(G - 1));
        if (last) tc_fence_acq_rel_gpu();                  // acquire: everything the other blocks released
        s_last = last;
    }
    __syncthreads();
    if (s_last) {
        if (K > 0) {
            double acc[K > 0 ? K : 1];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = 0.0;
            // fixed order: thread t owns partials t, t+T, t+2T, ... (ascending), then the fixed block tree
            for (int i0 = threadIdx.x; i0 < G; i0 += 4 * blockDim.x) {
                double buf[4][K > 0 ? K : 1];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = i0 + j * blockDim.x;
#pragma unroll
                    for (int k = 0; k < K; ++k) buf[j][k] = i < G ? __ldcg(B.part + k * TC_MAX_PARTIALS + i) : 0.0;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (i0 + j * (int)blockDim.x < G) {
#pragma unroll
                        for (int k = 0; k < K; ++k) acc[k] += buf[j][k];
                    }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double t = block_sum(acc[k], sh);
                if (threadIdx.x == 0) s_tot[k] = t;
            }
            __syncthreads();
            xchg(s_tot, K);
        }
        if (threadIdx.x == 0) {
            double* res = B.res + (gen & 1u) * TC_BAR_MAXK;
#pragma unroll
            for (int k = 0; k < K; ++k) res[k] = s_tot[k];
            tc_st_relaxed_gpu_u32(B.ctr, 0u);               // re-arm: nobody arrives again before the release below
        }
        __syncthreads();
        if (threadIdx.x < TC_BAR_FLAGS)                     // release (cumulative over thread 0's stores) + flag
            tc_st_release_gpu_u32(B.ctr + 32 * (1 + threadIdx.x), gen + 1u);
    } else if (threadIdx.x == 0) {
        const unsigned int* flag = B.ctr + 32 * (1 + (blockIdx.x % TC_BAR_FLAGS));
        while (tc_ld_relaxed_gpu_u32(flag) == gen) { }
        tc_fence_acq_rel_gpu();                             // acquire + stale L1 lines of this SM are dropped
    }
    __syncthreads();
    if (K > 0) {
        const double* res = B.res + (gen & 1u) * TC_BAR_MAXK;
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = __ldcg(res + k);
    }
    gen += 1u;
}

template <int K>
__device__ __forceinline__ void tc_grid_reduce(const TcGridBar& B, unsigned int& gen, double* v, double* sh) {
    tc_grid_reduce<K>(B, gen, v, sh, TcNoExchange());
}
__device__ __forceinline__ void tc_grid_barrier(const TcGridBar& B, unsigned int& gen, double* sh) {
    double dummy[1];
    tc_grid_reduce<0>(B, gen, dummy, sh, TcNoExchange());
}
