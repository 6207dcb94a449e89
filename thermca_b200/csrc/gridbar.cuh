// Grid-wide barrier with a deterministic K-value reduction for the persistent PCG kernels.
//
// Measured on B200 (scripts/barrier_bench.py; numbers in profiles/README.md).  The barrier itself is the scheme
// cooperative_groups uses — ONE release atomic per block on a counter whose top bit flips when the last block has
// arrived, relaxed polls, one acquire fence — because every extra hop costs an L2 round trip (~0.6 us):
// "last arriver sums the partials and publishes the totals" variants (tried: separate flag lines, LL-style
// result lines, two-level arrival) measured 4.0-6.7 us per reduction against 2.4 us for the plain barrier at
// 1184 blocks.  What does pay: fewer, larger blocks (296 x 1024 threads: 1.25 us per barrier; the arrivals
// serialise at the L2 atomic unit at ~1 ns each) and NOT invalidating L1 while polling (an acquire load per poll
// would do it and evict the gather lines of co-resident blocks that are still working).
// The reduction rides on the barrier: every block publishes its K partials before arriving and, once released,
// sums all G partials itself in block-index order with a fixed tree — bit-identical totals in every block (uniform
// control flow) and from run to run.  Partial banks alternate with the barrier parity, so a fast block that already
// publishes for the next reduction cannot overwrite what a slow block is still summing.
// The launch must be cooperative (all blocks co-resident).  K = 0 gives a plain barrier.
#pragma once
#include "common.cuh"

#define TC_BAR_MAXK 8

struct TcGridBar {
    unsigned int* ctr;      // word [0]: arrival counter (top bit = barrier phase).  Last-arriver variant: word [32] =
                            // its own arrival counter, result lines at word 32*(2+j), j < TC_BAR_FLAGS (one 128-byte
                            // line each).  Zero-initialised, TC_BAR_CTR_WORDS words.
    double* part;           // [2 parities][TC_BAR_MAXK banks][TC_MAX_PARTIALS]
    double* res;            // scratch for callers (the barrier itself does not use it)
};
#define TC_BAR_FLAGS 16                                   // result lines of the last-arriver variant (below)
#define TC_BAR_CTR_WORDS (32 * (2 + TC_BAR_FLAGS))
#define TC_BAR_PART_DOUBLES (2 * TC_BAR_MAXK * TC_MAX_PARTIALS)

__device__ __forceinline__ void tc_fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned int tc_ld_relaxed_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// release-only atomic: MEMBAR.ALL.GPU + ATOMG, without the L1 invalidation an acq_rel fence carries (checked in SASS)
__device__ __forceinline__ unsigned int tc_atom_add_release_gpu_u32(unsigned int* p, unsigned int v) {
    unsigned int old;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

// K block sums with one pair of barriers (same per-value summation order as block_sum); totals in ALL threads
template <int K>
__device__ __forceinline__ void tc_block_sum_k(double* v) {
    __shared__ double shk[TC_BAR_MAXK][32];
    __shared__ double tot[TC_BAR_MAXK];
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
            if (lane == 0) tot[k] = r;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = tot[k];
}

// v[k]: this block's contribution, valid in thread 0.  On return every thread of every block holds the grid totals
// in v[k].  `gen` counts the barriers of this launch (start at 0; only its parity is used).
// err / limit (optional): bound the wait — after `limit` clock64 ticks (or as soon as *err is non-zero) the block
// raises *err and carries on with whatever it has (the row-partitioned kernel: a lost peer must not hang the GPU).
struct TcNoWork { __device__ __forceinline__ void operator()() const {} };
// `work` (optional) runs on ALL threads between the block's arrival and its wait: work that neither feeds this
// reduction nor is read by other blocks before the NEXT barrier hides behind the wait (split-phase barrier).
template <int K, int U = (K > 3 ? 1 : 4), class F = TcNoWork>       // U partial loads per value in flight
__device__ __forceinline__ void tc_grid_reduce(const TcGridBar& B, unsigned int& gen, double* v, int* err = nullptr,
                                               long long limit = 0, F work = F()) {
    const int G = gridDim.x;
    double* bank = B.part + (size_t)(gen & 1u) * TC_BAR_MAXK * TC_MAX_PARTIALS;
    __syncthreads();                                       // every thread's earlier stores are ordered before thread 0's release
    unsigned int old = 0u;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) bank[k * TC_MAX_PARTIALS + blockIdx.x] = v[k];
        const unsigned int nb = blockIdx.x == 0 ? 0x80000000u - (unsigned int)(G - 1) : 1u;
        old = tc_atom_add_release_gpu_u32(B.ctr, nb);
    }
    work();
    if (threadIdx.x == 0) {
        if (err) {
            const long long t0 = clock64();
            for (unsigned int spins = 1; ((old ^ tc_ld_relaxed_gpu_u32(B.ctr)) & 0x80000000u) == 0u; ++spins)
                if ((spins & 255u) == 0u) {
                    if (*((volatile int*)err)) break;
                    if (clock64() - t0 > limit) { *((volatile int*)err) = 1; break; }
                }
        } else {
            while (((old ^ tc_ld_relaxed_gpu_u32(B.ctr)) & 0x80000000u) == 0u) { }
        }
        tc_fence_acq_rel_gpu();                            // acquire; stale L1 lines of this SM are dropped
    }
    __syncthreads();
    if (K > 0) {
        double acc[K > 0 ? K : 1];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.0;
        // fixed order: thread t owns partials t, t+T, t+2T, ... (ascending), then the fixed block tree
        for (int i0 = threadIdx.x; i0 < G; i0 += U * blockDim.x) {
            double buf[U][K > 0 ? K : 1];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int i = i0 + j * blockDim.x;
#pragma unroll
                for (int k = 0; k < K; ++k) buf[j][k] = i < G ? __ldcg(bank + k * TC_MAX_PARTIALS + i) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < U; ++j)
                if (i0 + j * (int)blockDim.x < G) {
#pragma unroll
                    for (int k = 0; k < K; ++k) acc[k] += buf[j][k];
                }
        }
        tc_block_sum_k<K>(acc);
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = acc[k];
    }
    gen += 1u;
}
__device__ __forceinline__ void tc_grid_barrier(const TcGridBar& B, unsigned int& gen, int* err = nullptr,
                                                long long limit = 0) {
    double dummy[1];
    tc_grid_reduce<0>(B, gen, dummy, err, limit);
}


// ---- last-arriver variant with an exchange hook (row-partitioned kernel, dist.cu) --------------------------------
// Across GPUs ONE block must trade the local totals with the peers and decide (a timeout has to be seen by all blocks
// alike), so here the last block to arrive sums the partials, runs the hook (peer mailboxes) and publishes the totals;
// everybody else only polls a result line.  The line carries the totals "LL style" (as NCCL's low-latency protocol
// does): every 8-byte granule is {32 data bits, 32-bit generation tag}; 8-byte accesses are single-copy atomic, so a
// poller that sees the tag in a granule also sees its data — the totals arrive WITH the release flag.
// Line layout (u64 granules): [0] tag only (written first, with release), [1 + 2k], [2 + 2k] = low / high half of
// total k.  The tag is monotonic across launches: read it once at kernel entry with tc_bar_generation.
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
__device__ __forceinline__ void tc_st_relaxed_gpu_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int tc_bar_generation(const TcGridBar& B) {
    const unsigned long long* line = reinterpret_cast<const unsigned long long*>(B.ctr + 64);
    return (unsigned int)(tc_ld_relaxed_gpu_u64(line) >> 32);
}

// v[k]: this block's contribution, valid in thread 0; `tag_gen`: completed last-arriver reductions (monotonic).
// xchg(tot, K): called by ALL threads of the last block with the local totals in shared memory `tot`; may replace
// them (multi-GPU sum); must contain its own __syncthreads() pairs.  On return all threads hold the totals.
template <int K, class X>
__device__ __forceinline__ void tc_grid_reduce_last(const TcGridBar& B, unsigned int& tag_gen, double* v, X xchg) {
    static_assert(K >= 1 && K <= 3, "a result line holds up to 3 totals");
    __shared__ int s_last;
    __shared__ double s_tot[TC_BAR_MAXK];
    const int G = gridDim.x;
    const unsigned int tag = tag_gen + 1u;
    double* bank = B.part + (size_t)(tag & 1u) * TC_BAR_MAXK * TC_MAX_PARTIALS;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) bank[k * TC_MAX_PARTIALS + blockIdx.x] = v[k];
        const int last = tc_atom_add_release_gpu_u32(B.ctr + 32, 1u) == (unsigned int)(G - 1);
        if (last) {
            tc_fence_acq_rel_gpu();                        // acquire: everything the other blocks released
            tc_st_relaxed_gpu_u32(B.ctr + 32, 0u);         // re-arm; ordered before the release stores below
        }
        s_last = last;
    }
    __syncthreads();
    if (s_last) {
        double acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.0;
        for (int i0 = threadIdx.x; i0 < G; i0 += 4 * blockDim.x) {
            double buf[4][K];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j * blockDim.x;
#pragma unroll
                for (int k = 0; k < K; ++k) buf[j][k] = i < G ? __ldcg(bank + k * TC_MAX_PARTIALS + i) : 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (i0 + j * (int)blockDim.x < G) {
#pragma unroll
                    for (int k = 0; k < K; ++k) acc[k] += buf[j][k];
                }
        }
        tc_block_sum_k<K>(acc);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) s_tot[k] = acc[k];
        }
        __syncthreads();
        xchg(s_tot, K);
        if (threadIdx.x < TC_BAR_FLAGS) {
            unsigned long long* line = reinterpret_cast<unsigned long long*>(B.ctr + 32 * (2 + threadIdx.x));
            tc_st_release_gpu_u64(line, (unsigned long long)tag << 32);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const unsigned long long bits = (unsigned long long)__double_as_longlong(s_tot[k]);
                tc_st_relaxed_gpu_u64(line + 1 + 2 * k, ((unsigned long long)tag << 32) | (bits & 0xffffffffULL));
                tc_st_relaxed_gpu_u64(line + 2 + 2 * k, ((unsigned long long)tag << 32) | (bits >> 32));
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = s_tot[k];
    } else {
        __shared__ double s_got[TC_BAR_MAXK];
        if (threadIdx.x == 0) {
            const unsigned long long* line =
                reinterpret_cast<const unsigned long long*>(B.ctr + 32 * (2 + (blockIdx.x % TC_BAR_FLAGS)));
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
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = s_got[k];
    }
    __syncthreads();
    tag_gen = tag;
}
