// Grid-wide barrier with the reduction folded in ("reduce-barrier") for the persistent PCG kernels.
//
// cooperative_groups::grid.sync() followed by "every block sums all G partials" costs G^2 L2 loads
// and two dependent round trips per reduction (measured round 1: 12.6 % of a 1 M-DOF PCG iteration,
// profiles/README.md).  Here the LAST block to arrive sums the G partials once, in block-index order
// with a fixed tree (bit-reproducible, independent of which block is last), publishes the totals in a
// parity slot and releases the generation flag every other block spins on:
//   arrive : partial[k][block] (st) ; atom.release add on the arrival counter             (one L2 atomic per block)
//   last   : acquire fence ; ordered sum of G partials ; [exchange hook: other GPUs] ; totals + generation tag
//            written into the result lines (tag granule last, st.release)
//   others : relaxed polls of their result line until the tag shows up (the totals ride in the same line) ;
//            one acquire fence (drops stale L1 lines)
// The launch must be cooperative (all blocks co-resident).  K = 0 gives a plain barrier.
// The exchange hook runs in the last block only, between the local sum and the release: the row-
// partitioned kernel (dist.cu) adds the peers' values there, so one barrier == one global all-reduce.
#pragma once
#include "common.cuh"

#define TC_BAR_MAXK 4

struct TcGridBar {
    unsigned int* ctr;      // word [0]: top-level arrival counter (one arrival per GROUP of TC_BAR_GROUP consecutive
                            // blocks); result lines at word 32*(1+j), j < TC_BAR_FLAGS (one 128-byte line each: the
                            // pollers are spread over several L2 lines); group counters at word 32*(1+TC_BAR_FLAGS+g),
                            // one line each.  Counters are re-armed by their last arriver, so kernels with different
                            // grids can share the state.  Zero-initialised, TC_BAR_CTR_WORDS words.
    double* part;           // TC_BAR_MAXK banks of TC_MAX_PARTIALS
    double* res;            // unused by the barrier itself (kept for callers that want a scratch slot)
};
#define TC_BAR_FLAGS 16                                   // blocks poll line (blockIdx % TC_BAR_FLAGS)
#define TC_BAR_GROUP 8                                    // blocks per first-level arrival counter
#define TC_BAR_MAX_GROUPS (TC_MAX_PARTIALS / TC_BAR_GROUP)
#define TC_BAR_CTR_WORDS (32 * (1 + TC_BAR_FLAGS + TC_BAR_MAX_GROUPS))

// Measured on B200 (scripts/barrier_bench.py, profiles/README.md): 1184 blocks hitting ONE counter serialise at the
// L2 atomic unit (~1 ns each: grid.sync() 2.4 us at 8 blocks/SM vs 1.2 us at <= 2), so arrivals are two-level: the 8
// blocks of a group hit their own line in parallel with all other groups, the last of a group arrives at the top.
//
// A result line carries the totals "LL style" (as NCCL's low-latency protocol does): every 8-byte granule is
// {32 data bits, 32-bit generation tag}; 8-byte accesses are single-copy atomic, so a poller that sees the new tag in
// a granule also sees its data — the totals arrive WITH the release flag, no second round trip to fetch them.
// Line layout (u64 granules): [0] barrier tag only, [1 + 2k], [2 + 2k] = low / high half of total k.
//
// Fences: fence.acq_rel.gpu (MEMBAR.ALL.GPU + L1 invalidate) is enough on the acquire side; CUDA's __threadfence() is
// the sequentially consistent fence.sc.gpu.  The wait loop polls with relaxed loads: an acquire load per poll would
// invalidate the SM's L1 on every iteration and evict the gather lines of co-resident blocks that are still working.
// The arrive is a release-only atomic (MEMBAR.ALL.GPU + ATOMG, no invalidation — checked in SASS).
__device__ __forceinline__ void tc_fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned long long tc_ld_relaxed_gpu_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void tc_st_release_gpu_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void tc_st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned int tc_atom_add_release_gpu_u32(unsigned int* p, unsigned int v) {
    unsigned int old;
    asm volatile("atom.release.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ unsigned int tc_atom_add_relaxed_gpu_u32(unsigned int* p, unsigned int v) {
    unsigned int old;
    asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void tc_st_relaxed_gpu_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generation = completed barriers on this state; every block of a launch reads the same value at entry
// (nobody can complete a barrier before all blocks have arrived, i.e. after they read it)
__device__ __forceinline__ unsigned int tc_bar_generation(const TcGridBar& B) {
    const unsigned long long* line = reinterpret_cast<const unsigned long long*>(B.ctr + 32);
    return (unsigned int)(tc_ld_relaxed_gpu_u64(line) >> 32);
}

struct TcNoExchange {
    __device__ __forceinline__ void operator()(double*, int) const {}
};

// K block sums with one pair of barriers (same per-value summation order as block_sum); results in out[] of thread 0
template <int K>
__device__ __forceinline__ void tc_block_sum_k(double* v, double* sh, double* out) {
    __shared__ double shk[TC_BAR_MAXK][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) shk[k][w] = v[k];
    }
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double r = lane < nw ? shk[k][lane] : 0.0;
            r = warp_sum(r);
            out[k] = r;
        }
    }
    (void)sh;
}

// v[k]: this block's contribution, valid in thread 0.  On return every thread of every block holds the
// grid totals in v[k].  `gen` is the caller's running generation counter (same value in all blocks; the
// kernel reads it once with tc_bar_generation at entry).  sh: >= 32 doubles of shared scratch.
// xchg(tot, K): called by ALL threads of the last block with the local totals in shared memory `tot`;
// may replace them (multi-GPU sum).  Must contain its own __syncthreads() pairs if it uses several threads.
template <int K, class X>
__device__ __forceinline__ void tc_grid_reduce(const TcGridBar& B, unsigned int& gen, double* v, double* sh, X xchg) {
    static_assert(K <= 3, "a result line holds up to 3 totals");
    __shared__ int s_last;
    __shared__ double s_tot[TC_BAR_MAXK];
    const int G = gridDim.x;
    const unsigned int tag = gen + 1u;
    __syncthreads();                                       // every thread's earlier stores are ordered before thread 0's release
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) B.part[k * TC_MAX_PARTIALS + blockIdx.x] = v[k];
        const int g = blockIdx.x / TC_BAR_GROUP;
        const int gsize = min(TC_BAR_GROUP, G - g * TC_BAR_GROUP), ngroups = (G + TC_BAR_GROUP - 1) / TC_BAR_GROUP;
        unsigned int* gctr = B.ctr + 32 * (1 + TC_BAR_FLAGS + g);
        // release: partials + everything this block wrote before (cumulative over the __syncthreads above)
        int last = 0;
        if (tc_atom_add_release_gpu_u32(gctr, 1u) == (unsigned int)(gsize - 1)) {
            tc_st_relaxed_gpu_u32(gctr, 0u);               // re-arm (nobody arrives again before the grid-wide release)
            tc_fence_acq_rel_gpu();                        // acquire my group's releases, release them to the top level
            if (tc_atom_add_relaxed_gpu_u32(B.ctr, 1u) == (unsigned int)(ngroups - 1)) {
                tc_st_relaxed_gpu_u32(B.ctr, 0u);
                tc_fence_acq_rel_gpu();                    // acquire: everything every block released
                last = 1;
            }
        }
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
            double tot[K > 0 ? K : 1];
            tc_block_sum_k<K>(acc, sh, tot);
            if (threadIdx.x == 0) {
#pragma unroll
                for (int k = 0; k < K; ++k) s_tot[k] = tot[k];
            }
            __syncthreads();
            xchg(s_tot, K);
        }
        // publish: thread j < TC_BAR_FLAGS writes result line j: the tag granule first, with release semantics
        // (cumulative over the whole grid's earlier writes through thread 0's acquire fences and the barriers
        // above), then the self-validating value granules
        if (threadIdx.x < TC_BAR_FLAGS) {
            unsigned long long* line = reinterpret_cast<unsigned long long*>(B.ctr + 32 * (1 + threadIdx.x));
            tc_st_release_gpu_u64(line, (unsigned long long)tag << 32);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const unsigned long long bits = (unsigned long long)__double_as_longlong(s_tot[k]);
                tc_st_relaxed_gpu_u64(line + 1 + 2 * k, ((unsigned long long)tag << 32) | (bits & 0xffffffffULL));
                tc_st_relaxed_gpu_u64(line + 2 + 2 * k, ((unsigned long long)tag << 32) | (bits >> 32));
            }
        }
        if (K > 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) v[k] = s_tot[k];
        }
    } else {
        __shared__ double s_got[TC_BAR_MAXK];
        if (threadIdx.x == 0) {
            const unsigned long long* line =
                reinterpret_cast<const unsigned long long*>(B.ctr + 32 * (1 + (blockIdx.x % TC_BAR_FLAGS)));
            while ((unsigned int)(tc_ld_relaxed_gpu_u64(line) >> 32) != tag) { }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                unsigned long long lo, hi;
                do { lo = tc_ld_relaxed_gpu_u64(line + 1 + 2 * k); } while ((unsigned int)(lo >> 32) != tag);
                do { hi = tc_ld_relaxed_gpu_u64(line + 2 + 2 * k); } while ((unsigned int)(hi >> 32) != tag);
                s_got[k] = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffULL)));
            }
            tc_fence_acq_rel_gpu();                         // acquire + stale L1 lines of this SM are dropped
        }
        __syncthreads();
        if (K > 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) v[k] = s_got[k];
        }
    }
    __syncthreads();
    gen = tag;
}

template <int K>
__device__ __forceinline__ void tc_grid_reduce(const TcGridBar& B, unsigned int& gen, double* v, double* sh) {
    tc_grid_reduce<K>(B, gen, v, sh, TcNoExchange());
}
__device__ __forceinline__ void tc_grid_barrier(const TcGridBar& B, unsigned int& gen, double* sh) {
    double dummy[1];
    tc_grid_reduce<0>(B, gen, dummy, sh, TcNoExchange());
}
