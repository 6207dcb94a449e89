// Row-partitioned implicit step for ONE large FE body across GPUs (BASELINE config #5):
// fused compute + collective.  Each rank owns a block of operator rows and runs ONE cooperative
// persistent kernel per theta step.  Inside that kernel
//   * ghost values of the gathered vector are written straight into the peers' HBM with peer
//     stores over NVLink (no NCCL call, no host round trip), driven by send lists, so any
//     partition works (contiguous slabs of a banded ordering give rank +-1 neighbours, a general
//     partition any neighbour set); every value carries its own sequence tag ("LL" encoding), so
//     there are no halo flags and no system-scope fences on the critical path;
//   * every PCG reduction is a local reduce-barrier (gridbar.cuh) followed by a one-way mailbox
//     exchange: block 0 writes the local sums into every rank's mailbox, every block polls its own
//     rank's mailbox and adds the values in rank order: bit-identical scalars on all ranks ->
//     identical convergence decisions, deterministic results;
//   * the search direction is double buffered, so the rows a peer needs are recomputed from
//     stable inputs by the pushing threads at the START of the vector phase (no intra-phase race,
//     no extra barrier): they cross NVLink while the local rows are updated and are in place a
//     phase and a barrier before the product needs them.
// Peer memory is exchanged once at setup as CUDA IPC handles (plumbing via torch.distributed in
// thermca_b200/dist.py).  The reference has nothing comparable (no multi-process code, SURVEY.md
// fact 1); per rank the arithmetic is the single-GPU PCG of pcg.cuh on the local rows.
//
// Failure handling: every wait on a peer is bounded (TC_DIST_TIMEOUT_S, default 20 s).  A rank that
// times out raises its error flag and contributes NaN to the next all-reduce, which makes every
// rank leave the iteration in the same reduction with status 2; the step is not committed
// (y unchanged) and tc_dist_check / tc_dist_get_state / tc_dist_stats return an error.
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#include "../../include/thermca_b200.h"
#include "parts.cuh"
#include "gridbar.cuh"

#define TC_MAX_WORLD 16
#define TC_DIST_K 7                     // values per all-reduce (single-reduction PCG: seven sums at once)
#define TC_MB_SLOT 16                   // u64 granules per (parity, sender) mailbox slot (2 per value, padded to 128 B)
#define TC_MB2_VALUES 64                // surface means of a partitioned body per exchange (second mailbox region)
#define TC_MB2_SLOT (2 * TC_MB2_VALUES)

typedef unsigned long long u64;

// Everything that crosses NVLink is "LL" encoded (the scheme of NCCL's low-latency protocol): a double travels as two
// 8-byte granules {32 data bits | 32-bit sequence tag}.  8-byte stores are single-copy atomic, so a reader that finds
// the expected tag in a granule also has its data: no flag, no system-scope fence on the sender (a MEMBAR.SYS would
// wait for the remote acknowledgements, i.e. a full NVLink round trip on the critical path), and every ghost value is
// usable the moment it lands.  Tags are monotonic sequence numbers (halo: hseq, reductions: rseq).
struct TcDistK {
    TcSell A;                             // local rows and LOCAL columns only (col16 relative to the row when the band
                                          // allows it, else col); couplings to ghosts live in the side lists below
    int n, n_ghost, rank, world;
    int n_bnd;                            // boundary slices (reference at least one ghost column)
    const int* bidx;                      // nslices: -1, or the index of the slice among the boundary slices
    int rot;                              // cyclic rotation of the slice order (see tc_dist_set_operator)
    const long long* gptr;                // n_bnd + 1: ghost couplings of boundary slice b (SELL-32: entry k of lane l at
    const int* gslot; const double* gval; //   gptr[b] + 32 k + l) = (ghost slot, operator value); padding = (0, 0.0)
    const double *b, *mass, *adiag;
    double *y, *p[2];                     // local vectors (n each); the search direction is double buffered
    double *x, *r, *q, *dinv;
    u64 *y_gl, *p_gl[2];                  // my ghost values, LL encoded: 2 granules per ghost slot (exported memory)
    long long n_send;                     // halo send list: local row -> (peer, ghost slot of the peer)
    const int* send_row; const int* send_peer; const long long* send_dst;
    u64* peer_y_gl[TC_MAX_WORLD]; u64* peer_p_gl[2][TC_MAX_WORLD];
    int n_nbr;                            // ranks I trade ghosts with (only its being non-zero matters to the kernel)
    u64* peer_mb[TC_MAX_WORLD]; u64* my_mb;      // mailboxes [2 parities][TC_MAX_WORLD senders][TC_MB_SLOT]
    u64* peer_mb2[TC_MAX_WORLD]; u64* my_mb2;    // surface-mean mailboxes [2][TC_MAX_WORLD][TC_MB2_SLOT]
    TcGridBar bar;
    int* state;                           // [0] reduction seq, [1] halo seq, [2] iters last, [3] iters total,
                                          // [4] comm error, [5] solves, [7] status of the last failed solve, [8] failed solves
    double theta, h, rtol;
    int maxit;
    // solve-only instantiation (stage solves of the W-methods on a partitioned body): z = (I + theta h A)^-1 rhs
    const double* rhs_in; double* x_out;  // local rows
    const double* h_dev;                  // step size on the device (the controller's), overrides h when non-null
    const int* done;                      // skip the launch when *done != 0
    const int* skip;                      // stiffness gate: *skip != 0 leaves x_out (= rhs, the explicit branch) untouched
    long long spin_limit;                 // clock64 ticks
    int push_fence;                       // experiment switch (TC_DIST_PUSH_FENCE=1): system fence after ghost pushes
    unsigned long long* prof;             // [2][16] phase times in ns: block 0 and the last block
    unsigned long long* blkprof;          // optional (TC_DIST_DIAG=3): [block][10]: 8 phase times (ns), SM id
    double* bstash;                       // n_bnd * 32: local row sums of the boundary slices between the two passes
    int x_in_dir;                         // x += alpha p rides on the direction phase (many rows per GPU) instead of
                                          // hiding behind the second all-reduce (few rows per GPU)
    // pipelined PCG (tc_dist_theta_pipe_kernel): extra vectors and the send lists grouped by boundary slice
    double *zh, *pv, *cc;                 // D z, search direction, c_i = 1 / (1 + shift a_ii)
    const int* sptr;                      // n_bnd + 1: send entries of boundary slice b (sorted by lane)
    const int* s_lane; const int* s_peer; const long long* s_dst;
};

__device__ __forceinline__ u64 ld_relaxed_sys_u64(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(u64* p, u64 v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void ll_store(u64* g, double v, unsigned int tag) {
    const u64 bits = (u64)__double_as_longlong(v);
    st_relaxed_sys_u64(g, ((u64)tag << 32) | (bits & 0xffffffffULL));
    st_relaxed_sys_u64(g + 1, ((u64)tag << 32) | (bits >> 32));
}
// bounded wait for one granule; on timeout (or when another thread of this rank already gave up) raises the error flag
__device__ __forceinline__ u64 ll_wait(const u64* g, unsigned int tag, int* err, long long limit) {
    u64 v = ld_relaxed_sys_u64(g);
    if ((unsigned int)(v >> 32) == tag) return v;
    const long long t0 = clock64();
    for (unsigned int spins = 1;; ++spins) {
        v = ld_relaxed_sys_u64(g);
        if ((unsigned int)(v >> 32) == tag) return v;
        if ((spins & 63u) == 0u) {                         // the bookkeeping loads stay off the polling path
            if (*((volatile int*)err)) return v;
            if (clock64() - t0 > limit) { *((volatile int*)err) = 1; __threadfence(); return v; }
        }
    }
}
__device__ __forceinline__ double ll_load(const u64* g, unsigned int tag, int* err, long long limit) {
    const u64 lo = ll_wait(g, tag, err, limit), hi = ll_wait(g + 1, tag, err, limit);
    return __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffULL)));
}
// both granules of one value with ONE 16-byte load (each 8-byte half is single-copy atomic on its own); the caller
// checks the tags and falls back on ll_load for a value that has not landed yet
__device__ __forceinline__ void ll_peek(const u64* g, u64& lo, u64& hi) {
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(g) : "memory");
}

// Global reduction = local reduce-barrier (gridbar.cuh: flip barrier, every block sums the G partials itself) followed by
// a mailbox exchange: block 0 writes the local totals into every rank's mailbox (LL encoded, sequence tagged, parity
// double buffered) and EVERY block polls its own rank's mailbox for the W values and adds them in rank order — the
// values cross NVLink once, one way, and nobody waits for a second local hop; all ranks and blocks get bit-identical
// totals.  A rank whose error flag is up contributes NaN: every rank then leaves the iteration in this very reduction.
// (The variant in which the last-arriving block exchanges and then publishes measured ~5 us more per reduction.)
// Every wait is bounded; once the error flag is up all waits of this rank return at once, so a failure drains fast.
// `work` (optional): split-phase — runs on all threads after the local totals are on their way to the peers and
// before the block polls its mailbox (one GPU: between the block's arrival and its wait, gridbar.cuh).
template <int K, class F = TcNoWork>
__device__ __forceinline__ void dist_reduce(const TcDistK& S, unsigned int& gen, double* red, int seq, F work = F()) {
    int* err = S.state + 4;
    if (S.world == 1) { tc_grid_reduce<K, (K > 3 ? 1 : 4)>(S.bar, gen, red, nullptr, 0, work); return; }
    tc_grid_reduce<K>(S.bar, gen, red, err, S.spin_limit);
    __shared__ double s_in[TC_MAX_WORLD][TC_DIST_K];
    __shared__ double s_out[TC_DIST_K];
    const int par = seq & 1;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) s_out[k] = red[k];
    }
    __syncthreads();
    const bool mine = (int)threadIdx.x < S.world * K;             // one thread per (peer, value): all waits run side by side
    const int peer = threadIdx.x / K, k = threadIdx.x - peer * K;
    if (mine && blockIdx.x == 0) {
        const bool bad = *((volatile int*)err) != 0;
        u64* out = S.peer_mb[peer] + ((size_t)par * TC_MAX_WORLD + S.rank) * TC_MB_SLOT;
        ll_store(out + 2 * k, bad ? nan("") : s_out[k], (unsigned int)seq);
    }
    work();                                                       // hidden behind the NVLink flight and the rank skew
    if (mine) {
        const u64* in = S.my_mb + ((size_t)par * TC_MAX_WORLD + peer) * TC_MB_SLOT;
        s_in[peer][k] = ll_load(in + 2 * k, (unsigned int)seq, err, S.spin_limit);
    }
    __syncthreads();
    if ((int)threadIdx.x < K) {
        double s = 0.0;
        for (int w = 0; w < S.world; ++w) s += s_in[w][threadIdx.x];
        s_out[threadIdx.x] = *((volatile int*)err) ? nan("") : s;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) red[k] = s_out[k];
    __syncthreads();
}

#define TC_DIST_THREADS 1024

// Couplings of one row of boundary slice `bi` to rows of other ranks: a short side list (<= ~6 entries per row on a
// slab cut), four ghost values requested at a time (16-byte LL peeks).  Kept out of line: it runs for ~1 % of the
// slices and must not cost the streaming loop registers.
__device__ __noinline__ double ghost_part(const long long* gptr, const int* gslot, const double* gval, const u64* xg,
                                          int bi, int lane, unsigned int hseq, int* state, long long limit) {
    const long long gb = gptr[bi];
    const int gw = (int)((gptr[bi + 1] - gb) >> 5);
    double sum = 0.0;
    for (int k0 = 0; k0 < gw; k0 += 4) {
        u64 lo[4], hi[4]; int sl[4]; double gv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k0 + j < gw) {
                sl[j] = gslot[gb + 32 * (k0 + j) + lane];
                gv[j] = gval[gb + 32 * (k0 + j) + lane];
                ll_peek(xg + 2 * (size_t)sl[j], lo[j], hi[j]);
            }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k0 + j < gw) {
                double xv;
                if ((unsigned int)(lo[j] >> 32) == hseq && (unsigned int)(hi[j] >> 32) == hseq) {
                    xv = __longlong_as_double((long long)((hi[j] << 32) | (lo[j] & 0xffffffffULL)));
                } else {
                    if (state[11]) atomicAdd(state + 10, 1);              // diagnostic: ghosts that had not landed
                    xv = ll_load(xg + 2 * (size_t)sl[j], hseq, state + 4, limit);
                }
                sum = __dadd_rn(sum, __dmul_rn(gv[j], xv));
            }
    }
    return sum;
}

// sum_k A[row][k] x[col] over the LOCAL columns of one slice (one row per lane): parts.cuh tc_slice_row_sum (entries in
// stored order, no FMA); EF = matrix streams with the L2 evict-first policy `pol`
template <bool EF = false>
__device__ __forceinline__ double slice_local_sum(const TcSell& A, const double* xl, int s, int lane, u64 pol = 0ULL) {
    return tc_slice_row_sum<0, EF>(A, xl, 0, s, lane, pol);
}

// Local rows times [local vector | ghosts].  Every slice runs the same streaming loop over its LOCAL columns, in the
// natural slice order up to a cyclic rotation (no permutation array: an index load ahead of slice_ptr would add a
// dependent round trip per slice).  The rotation puts the boundary slices (the leading and trailing runs of a block of
// contiguous rows) into the FIRST round of the grid-stride loop, on the highest-numbered warps: those are the warps
// without a slice in the remainder round, so the ~4 dependent loads of the ghost side list are absorbed by their slack;
// a slice with ghost couplings (looked up next to slice_ptr) then adds them from its side list.  The ghosts were
// pushed at the start of the previous vector phase, a whole phase and a barrier earlier (measured: practically never
// late); one that is still missing is waited for individually (LL tag).  No halo flags.
template <int MODE>
__device__ __forceinline__ void slab_finish_row(const TcDistK& S, const double* xl, double* y, const double* b, double shift,
                                                int row, double sum, double& dot) {
    if (MODE == SPMV_B) {
        y[row] = __dsub_rn(b[row], sum);
    } else {
        const double xr = tc_ld_x<0>(xl + row);
        const double q = S.mass[row] * (xr + shift * sum);
        y[row] = q;
        dot += xr * q;
    }
}
// Two passes per warp: the streaming pass over ALL its slices leaves the local sums of boundary slices in y and
// remembers their rounds in a bit mask; the ghost side lists (a non-inlined call with ~4 dependent loads) are added in
// a second pass over those rounds only (the local sums wait in a small scratch array) — the call, its argument registers and the spills around it stay out of the
// streaming loop (with the call inside it the row accumulator lived in local memory: 12 % slower products than the
// single-GPU kernel at 10 M rows), and the ghosts have had the whole pass to land.
template <int MODE>
__device__ __forceinline__ double slab_spmv(const TcDistK& S, const double* xl, const u64* xg, double* y,
                                            const double* b, double shift, unsigned int hseq) {
    const TcSell& A = S.A;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int warp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_block;
    const int nsl = A.nslices;
    double dot = 0.0;
    unsigned int bmask = 0u, bit = 1u;
    for (int idx = warp; idx < nsl; idx += nwarps) {
        int s = idx + S.rot;
        if (s >= nsl) s -= nsl;
        const int bi = S.n_bnd ? S.bidx[s] : -1;                     // warp uniform
        const double sum = slice_local_sum(A, xl, s, lane);
        const int row = s * 32 + lane;
        if (bi >= 0) {
            bmask |= bit;
            S.bstash[bi * 32 + lane] = sum;                          // local part, finished in the second pass (its own
                                                                     // scratch: y may alias b, tc_dist_residual_kernel)
        } else if (row < A.n) {
            slab_finish_row<MODE>(S, xl, y, b, shift, row, sum, dot);
        }
        bit = (bit << 1) | (bit >> 31);                              // round mod 32
    }
    if (bmask) {
        bit = 1u;
        for (int idx = warp; idx < nsl; idx += nwarps, bit = (bit << 1) | (bit >> 31)) {
            if (!(bmask & bit)) continue;
            int s = idx + S.rot;
            if (s >= nsl) s -= nsl;
            const int bi = S.bidx[s];
            if (bi < 0) continue;                                    // a round that shares the mask bit (> 32 rounds)
            const double gs = ghost_part(S.gptr, S.gslot, S.gval, xg, bi, lane, hseq, S.state, S.spin_limit);
            const int row = s * 32 + lane;
            if (row < A.n) slab_finish_row<MODE>(S, xl, y, b, shift, row, __dadd_rn(S.bstash[bi * 32 + lane], gs), dot);
        }
    }
    return dot;
}

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// phase profiler: thread 0 of block 0 and of the last block (every block with TC_DIST_DIAG=3) accumulate wall time per
// phase in SHARED memory and flush once at the end of the launch — a global read-modify-write per phase put an
// exposed memory round trip of thread 0 in front of every barrier (measured ~1.3 us per phase at 10 M rows)
#define PROF_INIT() const int prof_row = blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x - 1 ? 1 : -1); \
    __shared__ unsigned long long s_prof[8]; \
    const bool prof_on = threadIdx.x == 0 && (prof_row >= 0 || S.blkprof != nullptr); \
    if (prof_on) { for (int k_ = 0; k_ < 8; ++k_) s_prof[k_] = 0ULL; } \
    unsigned long long prof_t = prof_on ? gtime() : 0ULL
#define PROF(k) do { if (prof_on) { const unsigned long long t_ = gtime(); s_prof[(k)] += t_ - prof_t; prof_t = t_; } } while (0)
#define PROF_FLUSH() do { if (prof_on) { \
    if (prof_row >= 0) { for (int k_ = 0; k_ < 8; ++k_) S.prof[prof_row * 16 + k_] += s_prof[k_]; } \
    if (S.blkprof) { unsigned int sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); \
        for (int k_ = 0; k_ < 8; ++k_) S.blkprof[blockIdx.x * 10 + k_] += s_prof[k_]; S.blkprof[blockIdx.x * 10 + 8] = sm_; } } } while (0)

// scaled right-hand side and Jacobi factor of one row (explicit roundings: the owner's loop and the halo push
// must produce the same bits)
__device__ __forceinline__ void init_row(const TcDistK& S, const double* f, int i, double shift, double& bsc, double& d) {
    const double m = S.mass[i];
    bsc = __dmul_rn(m, f[i]);
    d = __ddiv_rn(1.0, __dmul_rn(m, __fma_rn(shift, S.adiag[i], 1.0)));
}

// ONE cooperative launch per theta step: right-hand side, Jacobi-PCG (the iteration of pcg.cuh), update.
// A single-reduction CG (Chronopoulos & Gear) was measured too: it saves one all-reduce per iteration (~10 us) but its
// vector phase streams 12 arrays and loses the L2 residency of the 8-array form (23 vs 13 us at 1.25 M rows per GPU):
// 789 vs 957 M DOF*steps/s at N = 2 (profiles/README.md), so the classic two-reduction form stays.
// SOLVE = true: the same iteration as a stage solve of a W-method, z = (I + theta h A)^-1 rhs_in -> x_out (no product
// with y, no commit), h read from the device.
template <int MB, int THREADS = TC_DIST_THREADS, bool SOLVE = false>
__global__ void __launch_bounds__(THREADS, MB)
tc_dist_theta_kernel(TcDistK S) {
    __shared__ double sh[32];
    if (SOLVE && ((S.done && *S.done) || (S.skip && *S.skip))) return;     // uniform over the grid and the ranks
    const int n = S.n;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    unsigned int gen = 0u;
    int rseq = S.state[0];
    unsigned int hseq = (unsigned int)S.state[1];
    PROF_INIT();
    const double hstep = (SOLVE && S.h_dev) ? *S.h_dev : S.h;
    const double shift = S.theta * hstep;
    int pb = 0;
    const double* f = SOLVE ? S.rhs_in : S.q;

    // ---- right-hand side f = b - A y (thermca/solver.py:281) on the local rows -----------------
    if (!SOLVE) {
        ++hseq;
        for (long long e = gtid; e < S.n_send; e += gsz)
            ll_store(S.peer_y_gl[S.send_peer[e]] + 2 * S.send_dst[e], S.y[S.send_row[e]], hseq);
        slab_spmv<SPMV_B>(S, S.y, S.y_gl, S.q, S.b, 0.0, hseq);
        tc_grid_barrier(S.bar, gen, S.world > 1 ? S.state + 4 : nullptr, S.spin_limit);
    }
    double red[TC_DIST_K];
    ++hseq;
    {
        double* p0 = S.p[0];
        double rho_l = 0.0, bb_l = 0.0;
        // rows a peer needs first: recomputed from the same stable inputs (bit-identical), pushed over NVLink
        for (long long e = gtid; e < S.n_send; e += gsz) {
            double bsc, d;
            init_row(S, f, S.send_row[e], shift, bsc, d);
            ll_store(S.peer_p_gl[0][S.send_peer[e]] + 2 * S.send_dst[e], __dmul_rn(d, bsc), hseq);
        }
        for (int i = gtid; i < n; i += gsz) {
            double bsc, d;
            init_row(S, f, i, shift, bsc, d);
            const double z = __dmul_rn(d, bsc);
            S.r[i] = bsc; S.x[i] = 0.0; S.dinv[i] = d; p0[i] = z;
            rho_l += bsc * z; bb_l += bsc * bsc;
        }
        red[0] = block_sum(rho_l, sh); red[1] = block_sum(bb_l, sh); red[2] = 0.0;
    }
    dist_reduce<2>(S, gen, red, ++rseq);
    double rho = red[0];
    const double bb = red[1];
    const double tol2 = S.rtol * S.rtol * bb;
    int it = 0, status = 0;
    if (bb > 0.0) {
        status = 1;
        for (; it < S.maxit;) {
            PROF(0);
            const double* pe = S.p[pb];
            const double dot = slab_spmv<SPMV_SHIFT>(S, pe, S.p_gl[pb], S.q, nullptr, shift, hseq);
            PROF(1);
            red[0] = block_sum(dot, sh);
            PROF(2);
            dist_reduce<1>(S, gen, red, ++rseq);
            PROF(3);
            const double pq = red[0];
            if (!(pq > 0.0) || !isfinite(pq)) { status = 2; break; }       // uniform over the grid AND over the ranks
            const double alpha = rho / pq;
            double rn = 0.0, rr = 0.0;
            for (int i = gtid; i < n; i += gsz) {
                const double ri = S.r[i] - alpha * S.q[i];
                S.r[i] = ri;
                rn += ri * (S.dinv[i] * ri);
                rr += ri * ri;
            }
            red[0] = block_sum(rn, sh); red[1] = block_sum(rr, sh);
            PROF(4);
            // x += alpha p feeds neither this reduction nor another block: it runs while the block waits for the
            // slowest rank's sums (split-phase barrier), 3 of the 7 streams of the update phase
            // (with many rows per GPU the phases are bandwidth bound and nothing hides: there the update rides on the
            // direction phase, which reads p anyway — one stream less, as in the single-GPU kernel)
            dist_reduce<2>(S, gen, red, ++rseq, [&]() {
                if (!S.x_in_dir)
                    for (int i = gtid; i < n; i += gsz) S.x[i] += alpha * pe[i];
            });
            PROF(5);
            ++it;
            if (!(red[1] > tol2)) {
                status = isfinite(red[1]) ? 0 : 2;
                if (S.x_in_dir)
                    for (int i = gtid; i < n; i += gsz) S.x[i] += alpha * pe[i];
                break;
            }
            const double beta = red[0] / rho;
            double* pn = S.p[pb ^ 1];
            ++hseq;
            // ghost rows first, so they are in flight while the local rows are updated
            for (long long e = gtid; e < S.n_send; e += gsz) {
                const int i = S.send_row[e];
                ll_store(S.peer_p_gl[pb ^ 1][S.send_peer[e]] + 2 * S.send_dst[e],
                         __fma_rn(beta, pe[i], __dmul_rn(S.dinv[i], S.r[i])), hseq);
            }
            if (S.x_in_dir) {
                for (int i = gtid; i < n; i += gsz) {
                    const double pi = pe[i];
                    S.x[i] += alpha * pi;
                    pn[i] = __fma_rn(beta, pi, __dmul_rn(S.dinv[i], S.r[i]));
                }
            } else {
                for (int i = gtid; i < n; i += gsz) pn[i] = __fma_rn(beta, pe[i], __dmul_rn(S.dinv[i], S.r[i]));
            }
            rho = red[0];
            pb ^= 1;
            PROF(6);
            tc_grid_barrier(S.bar, gen, S.world > 1 ? S.state + 4 : nullptr, S.spin_limit);
            PROF(7);
        }
    }
    // ---- y += h * z (not committed after a communication failure / breakdown) ------------------
    if (SOLVE) {
        // a failed stage solve leaves NaN: the controller rejects the attempt on every rank alike
        for (int i = gtid; i < n; i += gsz) S.x_out[i] = status != 2 ? S.x[i] : nan("");
    } else if (status != 2) {
        for (int i = gtid; i < n; i += gsz) S.y[i] += S.h * S.x[i];
    }
    PROF_FLUSH();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S.state[0] = rseq; S.state[1] = (int)hseq; S.state[2] = it; S.state[3] += it; S.state[5] += 1;
        if (status) { S.state[7] = status; S.state[8] += 1; }
    }
}

// ---- pipelined PCG: ONE sweep over the operator and ONE reduction per iteration -------------------------------
// Strong scaling (10 M rows over 8 GPUs = 1.3 M rows per GPU, 40 us of streaming per iteration) is bounded by the
// fixed costs of an iteration: the classic form has two all-reduces and a barrier (~25 us) between its three phases.
// The pipelined conjugate gradients of Ghysels & Vanroose (Parallel Computing 40, 2014) keep, next to the residual
// recurrences, w = A u and z = A q by recurrences of their own, so BOTH dot products of an iteration come from
// vectors that are known before the product: they are summed in one reduction.  Here all vector recurrences are
// fused into the product sweep (every row is updated by the thread that just formed its row sum, the gathered
// vector is double buffered), which leaves one sweep + one reduce-barrier / all-reduce per iteration at the
// streamed bytes of the classic form (14 vector streams instead of 13, no separately stored product).
// Jacobi-scaled variables (D = diag(d), d_i = 1 / (M_i (1 + shift A_ii)) = 1 / g_i):
//   u = D r,  wh = D w,  zh = D z,  q = D s;   B = D (M + shift K) = c o (I + shift A),  c_i = 1 / (1 + shift A_ii)
//   per iteration:  nh = B wh;  zh = nh + beta zh;  q = wh + beta q;  p = u + beta p;  x += alpha p;
//                   u -= alpha q;  wh -= alpha zh;   gamma = sum g u^2,  delta = sum g wh u,  rr = sum (g u)^2
//   beta = gamma / gamma_old,  alpha = gamma / (delta - beta gamma / alpha_old)   (first iteration: beta = 0)
// Same iteration counts and accuracy as the classic form on the cuboid operators (CPU study, oracle/theta_oracle.py
// theta_pipecg_steps: 34/34, 95/95, 434/434 iterations at h = 1, 10, 100 s; identical error against the direct solver).
// Ghosts: the warp that forms a boundary row pushes the new wh value straight from its register (per-slice send
// lists, LL encoded); the boundary slices run first in every warp's sequence (rot), so the values are on their
// way a whole sweep and a reduction before a peer gathers them.
__device__ __forceinline__ double shfl_f64(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}

template <bool FIRST, bool EF>
__device__ __forceinline__ void pipe_sweep(const TcDistK& S, const double* wo, const u64* wo_g, double* wn, int pbw,
                                           double alpha, double beta, double shift, unsigned int hread,
                                           unsigned int hwrite, u64 pol, double* acc) {
    const TcSell& A = S.A;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int warp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_block;
    const int nsl = A.nslices;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int idx = warp; idx < nsl; idx += nwarps) {
        int s = idx + S.rot;
        if (s >= nsl) s -= nsl;
        const int bi = S.n_bnd ? S.bidx[s] : -1;
        double sum = slice_local_sum<EF>(A, wo, s, lane, pol);
        if (bi >= 0)
            sum = __dadd_rn(sum, ghost_part(S.gptr, S.gslot, S.gval, wo_g, bi, lane, hread, S.state, S.spin_limit));
        const int row = s * 32 + lane;
        double wnew = 0.0;
        if (row < A.n) {
            const double g = S.dinv[row], c = S.cc[row];
            const double woi = tc_ld_x<0>(wo + row);
            const double nh = c * (woi + shift * sum);
            if (FIRST) {                                   // wh_0 = B u_0 (wo holds u_0)
                wnew = nh;
                const double gu = g * woi;
                a0 += woi * gu; a1 += wnew * gu; a2 += gu * gu;
            } else {
                const double z = nh + beta * S.zh[row];
                const double qn = woi + beta * S.q[row];
                const double uo = S.r[row];
                const double pn = uo + beta * S.pv[row];
                S.x[row] += alpha * pn;
                const double un = uo - alpha * qn;
                wnew = woi - alpha * z;
                S.zh[row] = z; S.q[row] = qn; S.pv[row] = pn; S.r[row] = un;
                const double gu = g * un;
                a0 += un * gu; a1 += wnew * gu; a2 += gu * gu;
            }
            wn[row] = wnew;
        }
        if (bi >= 0) {                                     // rows of this slice that peers gather
            const int e0 = S.sptr[bi], e1 = S.sptr[bi + 1];
            for (int e = e0; e < e1; e += 32) {
                const int ee = e + lane;
                const bool act = ee < e1;
                const double v = shfl_f64(wnew, act ? S.s_lane[ee] : 0);
                if (act) ll_store(S.peer_p_gl[pbw][S.s_peer[ee]] + 2 * S.s_dst[ee], v, hwrite);
            }
        }
    }
    acc[0] = a0; acc[1] = a1; acc[2] = a2;
}

template <int MB, int THREADS, bool EF>
__global__ void __launch_bounds__(THREADS, MB)
tc_dist_theta_pipe_kernel(TcDistK S) {
    const int n = S.n;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    unsigned int gen = 0u;
    int rseq = S.state[0];
    unsigned int hseq = (unsigned int)S.state[1];
    PROF_INIT();
    const double shift = S.theta * S.h;
    const u64 pol = EF ? l2_evict_first_policy() : 0ULL;
    int* berr = S.world > 1 ? S.state + 4 : nullptr;

    // ---- right-hand side f = b - A y (thermca/solver.py:281) on the local rows -> S.q ---------------------
    ++hseq;
    for (long long e = gtid; e < S.n_send; e += gsz)
        ll_store(S.peer_y_gl[S.send_peer[e]] + 2 * S.send_dst[e], S.y[S.send_row[e]], hseq);
    slab_spmv<SPMV_B>(S, S.y, S.y_gl, S.q, S.b, 0.0, hseq);
    tc_grid_barrier(S.bar, gen, berr, S.spin_limit);

    // ---- u_0 = D M f (double-buffer slot 0 + ghosts), coefficient vectors, zeroed recurrences ------------
    ++hseq;
    for (long long e = gtid; e < S.n_send; e += gsz) {
        const int i = S.send_row[e];
        const double m = S.mass[i];
        const double g = __dmul_rn(m, __fma_rn(shift, S.adiag[i], 1.0));
        ll_store(S.peer_p_gl[0][S.send_peer[e]] + 2 * S.send_dst[e], __ddiv_rn(__dmul_rn(m, S.q[i]), g), hseq);
    }
    for (int i = gtid; i < n; i += gsz) {
        const double m = S.mass[i];
        const double t = __fma_rn(shift, S.adiag[i], 1.0);
        const double g = __dmul_rn(m, t);
        const double u0 = __ddiv_rn(__dmul_rn(m, S.q[i]), g);
        S.r[i] = u0; S.p[0][i] = u0; S.dinv[i] = g; S.cc[i] = __ddiv_rn(1.0, t);
        S.x[i] = 0.0; S.zh[i] = 0.0; S.q[i] = 0.0; S.pv[i] = 0.0;
    }
    tc_grid_barrier(S.bar, gen, berr, S.spin_limit);

    // ---- wh_0 = B u_0 and the first dot products --------------------------------------------------------
    double red[TC_DIST_K];
    unsigned int hread = hseq;
    ++hseq;
    pipe_sweep<true, EF>(S, S.p[0], S.p_gl[0], S.p[1], 1, 0.0, 0.0, shift, hread, hseq, pol, red);
    tc_block_sum_k<3>(red);
    dist_reduce<3>(S, gen, red, ++rseq);
    hread = hseq;
    int pb = 1;
    double gamma = red[0], delta = red[1], gamma_old = 0.0, alpha_old = 0.0;
    const double bb = red[2];
    const double tol2 = S.rtol * S.rtol * bb;
    int it = 0, status = 0;
    if (bb > 0.0) {
        status = 1;
        for (; it < S.maxit;) {
            PROF(0);
            double alpha, beta;
            if (it == 0) { beta = 0.0; alpha = gamma / delta; }
            else { beta = gamma / gamma_old; alpha = gamma / (delta - beta * gamma / alpha_old); }
            if (!(alpha > 0.0) || !isfinite(alpha)) { status = 2; break; }     // uniform over the grid AND the ranks
            ++hseq;
            pipe_sweep<false, EF>(S, S.p[pb], S.p_gl[pb], S.p[pb ^ 1], pb ^ 1, alpha, beta, shift, hread, hseq, pol, red);
            PROF(1);
            tc_block_sum_k<3>(red);
            PROF(2);
            dist_reduce<3>(S, gen, red, ++rseq);
            PROF(3);
            hread = hseq;
            pb ^= 1;
            ++it;
            if (!(red[2] > tol2)) { status = isfinite(red[2]) ? 0 : 2; break; }
            gamma_old = gamma; alpha_old = alpha;
            gamma = red[0]; delta = red[1];
        }
    }
    // ---- y += h * z (not committed after a communication failure / breakdown) ----------------------------
    if (status != 2)
        for (int i = gtid; i < n; i += gsz) S.y[i] += S.h * S.x[i];
    PROF_FLUSH();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S.state[0] = rseq; S.state[1] = (int)hseq; S.state[2] = it; S.state[3] += it; S.state[5] += 1;
        if (status) { S.state[7] = status; S.state[8] += 1; }
    }
}

// ---- single-reduction classic PCG: ONE all-reduce and one local barrier per iteration -------------------------------
// The iterates of the classic form (same alpha, beta up to rounding) with both of its reductions taken in ONE: the sums
//   p.q,  rho = r.Dr,  r.Dq,  q.Dq,  r.r,  r.q,  q.q
// are all known once q = (M + shift K) p has been formed, and they give the scalars of the whole iteration:
//   alpha = rho / p.q,   rho' = rho - 2 alpha r.Dq + alpha^2 q.Dq = r'.Dr',   r'.r' = r.r - 2 alpha r.q + alpha^2 q.q,
//   beta = rho' / rho
// (the recurrences for the inner products are those of D'Azevedo, Eijkhout & Romine, LAPACK working note 56, 1993; rho
// and r.r are summed afresh from the vectors in every sweep, only beta and the stopping test use the expanded forms, so
// nothing drifts).  The product sweep reads r and d next to the operator (2 more streams), and the update and direction
// phases of the classic kernel merge into one vector phase (x += alpha p; r -= alpha q; p' = D r + beta p: 8 streams
// instead of 7 + 4), run by the thread that owns the row in the sweep: the owner of a boundary row pushes p' to the
// peers straight from its register.  Per iteration: sweep -> all-reduce of 7 -> vector phase -> local barrier, i.e. one
// all-reduce less than the classic kernel at fewer streamed bytes.  CPU study (oracle/theta_oracle.py theta_srcg_steps):
// identical iteration counts and errors against the direct solver for h = 1 .. 100 s, rtol 1e-8 .. 1e-12.
template <bool EF>
__device__ __forceinline__ void sr_sweep(const TcDistK& S, const double* pe, const u64* pe_g, unsigned int hread,
                                         double shift, u64 pol, double* acc) {
    const TcSell& A = S.A;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int warp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_block;
    const int nsl = A.nslices;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0, a6 = 0.0;
    for (int idx = warp; idx < nsl; idx += nwarps) {
        int s = idx + S.rot;
        if (s >= nsl) s -= nsl;
        const int bi = S.n_bnd ? S.bidx[s] : -1;
        double sum = slice_local_sum<EF>(A, pe, s, lane, pol);
        if (bi >= 0)
            sum = __dadd_rn(sum, ghost_part(S.gptr, S.gslot, S.gval, pe_g, bi, lane, hread, S.state, S.spin_limit));
        const int row = s * 32 + lane;
        if (row < A.n) {
            const double pi = tc_ld_x<0>(pe + row);
            const double q = S.mass[row] * (pi + shift * sum);
            S.q[row] = q;
            const double r = S.r[row], d = S.dinv[row];
            const double dr = d * r, dq = d * q;
            a0 += pi * q; a1 += r * dr; a2 += r * dq; a3 += q * dq; a4 += r * r; a5 += r * q; a6 += q * q;
        }
    }
    acc[0] = a0; acc[1] = a1; acc[2] = a2; acc[3] = a3; acc[4] = a4; acc[5] = a5; acc[6] = a6;
}

// x += alpha p; r -= alpha q; p' = D r + beta p (written to the other buffer and pushed to the peers that gather it).
// LAST: the iteration ends here, no new direction; !SOLVE: y += h x is committed in the same pass.
template <bool LAST, bool SOLVE>
__device__ __forceinline__ void sr_vector_phase(const TcDistK& S, const double* pe, double* pn, int pbw, double alpha,
                                                double beta, unsigned int hwrite, bool commit) {
    const TcSell& A = S.A;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int warp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_block;
    const int nsl = A.nslices;
    for (int idx = warp; idx < nsl; idx += nwarps) {
        int s = idx + S.rot;
        if (s >= nsl) s -= nsl;
        const int row = s * 32 + lane;
        double pnew = 0.0;
        if (row < A.n) {
            const double pi = pe[row];
            const double xn = S.x[row] + alpha * pi;
            if (LAST) {                                   // a failed solve leaves NaN (SOLVE) / an untouched state
                if (SOLVE) S.x_out[row] = commit ? xn : nan("");
                else if (commit) S.y[row] += S.h * xn;
            } else {
                S.x[row] = xn;
                const double rn = S.r[row] - alpha * S.q[row];
                S.r[row] = rn;
                pnew = __fma_rn(beta, pi, __dmul_rn(S.dinv[row], rn));
                pn[row] = pnew;
            }
        }
        if (!LAST && S.n_bnd) {
            const int bi = S.bidx[s];
            if (bi >= 0) {
                const int e0 = S.sptr[bi], e1 = S.sptr[bi + 1];
                for (int e = e0; e < e1; e += 32) {
                    const int ee = e + lane;
                    const bool act = ee < e1;
                    const double v = shfl_f64(pnew, act ? S.s_lane[ee] : 0);
                    if (act) ll_store(S.peer_p_gl[pbw][S.s_peer[ee]] + 2 * S.s_dst[ee], v, hwrite);
                }
            }
        }
    }
}

template <int MB, int THREADS, bool EF, bool SOLVE = false>
__global__ void __launch_bounds__(THREADS, MB)
tc_dist_theta_sr_kernel(TcDistK S) {
    if (SOLVE && ((S.done && *S.done) || (S.skip && *S.skip))) return;     // uniform over the grid and the ranks
    const int n = S.n;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    unsigned int gen = 0u;
    int rseq = S.state[0];
    unsigned int hseq = (unsigned int)S.state[1];
    PROF_INIT();
    const double hstep = (SOLVE && S.h_dev) ? *S.h_dev : S.h;
    const double shift = S.theta * hstep;
    const u64 pol = EF ? l2_evict_first_policy() : 0ULL;
    int* berr = S.world > 1 ? S.state + 4 : nullptr;
    const double* f = SOLVE ? S.rhs_in : S.q;

    // ---- right-hand side f = b - A y (thermca/solver.py:281) on the local rows ----------------------------
    if (!SOLVE) {
        ++hseq;
        for (long long e = gtid; e < S.n_send; e += gsz)
            ll_store(S.peer_y_gl[S.send_peer[e]] + 2 * S.send_dst[e], S.y[S.send_row[e]], hseq);
        slab_spmv<SPMV_B>(S, S.y, S.y_gl, S.q, S.b, 0.0, hseq);
        tc_grid_barrier(S.bar, gen, berr, S.spin_limit);
    }
    // ---- r_0 = M f, d, p_0 = D r_0 (+ ghosts, recomputed from the same stable inputs: bit-identical) ------
    ++hseq;
    for (long long e = gtid; e < S.n_send; e += gsz) {
        double bsc, d;
        init_row(S, f, S.send_row[e], shift, bsc, d);
        ll_store(S.peer_p_gl[0][S.send_peer[e]] + 2 * S.send_dst[e], __dmul_rn(d, bsc), hseq);
    }
    for (int i = gtid; i < n; i += gsz) {
        double bsc, d;
        init_row(S, f, i, shift, bsc, d);
        S.r[i] = bsc; S.x[i] = 0.0; S.dinv[i] = d; S.p[0][i] = __dmul_rn(d, bsc);
    }
    tc_grid_barrier(S.bar, gen, berr, S.spin_limit);

    double red[TC_DIST_K];
    int pb = 0, it = 0, status = 1;
    bool consumed = false;                           // the last vector phase has already committed x
    double tol2 = 0.0;
    for (;;) {
        PROF(0);
        sr_sweep<EF>(S, S.p[pb], S.p_gl[pb], hseq, shift, pol, red);
        PROF(1);
        tc_block_sum_k<7>(red);
        PROF(2);
        dist_reduce<7>(S, gen, red, ++rseq);
        PROF(3);
        const double pq = red[0], rho = red[1], rr = red[4];
        if (it == 0) {
            tol2 = S.rtol * S.rtol * rr;
            if (!(rr > 0.0)) { status = isfinite(rr) ? 0 : 2; break; }           // zero right-hand side: x = 0
        }
        if (!(pq > 0.0) || !isfinite(pq)) { status = 2; break; }                // uniform over the grid AND the ranks
        const double alpha = rho / pq;
        const double rho_new = rho - 2.0 * alpha * red[2] + alpha * alpha * red[3];
        const double rr_new = rr - 2.0 * alpha * red[5] + alpha * alpha * red[6];
        ++it;
        const bool conv = !(rr_new > tol2);
        if (conv || it >= S.maxit) {
            status = conv ? (isfinite(rr_new) ? 0 : 2) : 1;
            sr_vector_phase<true, SOLVE>(S, S.p[pb], nullptr, 0, alpha, 0.0, 0u, status != 2);
            PROF(4);
            consumed = true;
            break;
        }
        const double beta = rho_new / rho;
        ++hseq;
        sr_vector_phase<false, SOLVE>(S, S.p[pb], S.p[pb ^ 1], pb ^ 1, alpha, beta, hseq, false);
        pb ^= 1;
        PROF(4);
        tc_grid_barrier(S.bar, gen, berr, S.spin_limit);
        PROF(7);
    }
    if (!consumed) {
        // left before a vector phase consumed x: zero right-hand side (x = 0) or a breakdown
        if (SOLVE) { for (int i = gtid; i < n; i += gsz) S.x_out[i] = status != 2 ? S.x[i] : nan(""); }
        else if (status != 2) { for (int i = gtid; i < n; i += gsz) S.y[i] += S.h * S.x[i]; }
    }
    PROF_FLUSH();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S.state[0] = rseq; S.state[1] = (int)hseq; S.state[2] = it; S.state[3] += it; S.state[5] += 1;
        if (status) { S.state[7] = status; S.state[8] += 1; }
    }
}

// ---- host ------------------------------------------------------------------------------------
static int dist_threads() {
    static int t = -1;
    if (t < 0) {
        const char* e = getenv("TC_PCG_THREADS");
        t = e ? atoi(e) : TC_DIST_THREADS;
        if (t != 256 && t != 512 && t != 768 && t != 1024) t = TC_DIST_THREADS;
    }
    return t;
}
static void* dist_fn() {
    // 2 blocks x 1024 threads per SM by default (fewer barrier participants, profiles/README.md); TC_PCG_THREADS=512/256
    if (dist_threads() == 768) return (void*)tc_dist_theta_kernel<2, 768>;     // 42 registers, 48 warps per SM
    if (dist_threads() == 512) return (void*)tc_dist_theta_kernel<4, 512>;
    if (dist_threads() == 256) return (void*)tc_dist_theta_kernel<8, 256>;
    return (void*)tc_dist_theta_kernel<2, 1024>;
}

struct tc_dist {
    cudaStream_t stream;
    int rank, world;
    long long n, n_ghost;
    char* comm = nullptr;             // my exported buffer: LL ghost arrays + mailboxes
    size_t comm_bytes = 0;
    char* peer[TC_MAX_WORLD];
    long long peer_ghost[TC_MAX_WORLD];   // ghost count of every rank (fixes the layout of its buffer)
    double *y = nullptr, *p0 = nullptr, *p1 = nullptr;
    double *x = nullptr, *r = nullptr, *q = nullptr, *dinv = nullptr;
    int* state = nullptr;
    unsigned long long* prof = nullptr;
    unsigned long long* blkprof = nullptr;
    int* send_row = nullptr; int* send_peer = nullptr; long long* send_dst = nullptr;
    std::vector<int> h_send_row, h_send_peer;      // host copies: regrouped by boundary slice for the pipelined kernel
    std::vector<long long> h_send_dst;
    double *zh = nullptr, *pv = nullptr, *cc = nullptr;
    double* bstash = nullptr;
    int *sptr = nullptr, *s_lane = nullptr, *s_peer = nullptr; long long* s_dst = nullptr;
    int pcg_mode = 0;                 // 0 classic (two reductions + barrier), 1 pipelined (one sweep, one reduction)
    bool pipe_ok = false;             // every send row lies in a boundary slice (true for symmetric patterns)
    int pipe_shape = 216, pipe_threads = 512, pipe_grid = 0, rot_pipe = 0;     // shape shared by the pipelined and the
    int sr_grid = 0, rot_sr = 0;                                                  // single-reduction kernel (TC_PIPE_SHAPE)
    bool l2_hint = true;
    TcDistK K;
    int grid = 0;
    bool peers_open = false;
    bool peers_local = false;         // peers are sibling handles of this process (tc_dist_open_peers_local): no IPC mappings
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
// layout of a rank's exported buffer: [y ghosts | p ghosts 0 | p ghosts 1 | reduction mailboxes | surface-mean
// mailboxes], ghosts LL encoded (16 B each)
static size_t gl_bytes(long long n_ghost) { return align_up((size_t)(n_ghost > 0 ? n_ghost : 1) * 2 * sizeof(u64), 256); }
static size_t off_gl(long long n_ghost, int k) { return (size_t)k * gl_bytes(n_ghost); }
static size_t off_mb(long long n_ghost) { return 3 * gl_bytes(n_ghost); }
static size_t mb_bytes() { return align_up(sizeof(u64) * 2 * TC_MAX_WORLD * TC_MB_SLOT, 256); }
static size_t off_mb2(long long n_ghost) { return off_mb(n_ghost) + mb_bytes(); }
static size_t mb2_bytes() { return align_up(sizeof(u64) * 2 * TC_MAX_WORLD * TC_MB2_SLOT, 256); }

extern "C" int tc_dist_create(tc_dist** out, void* stream, int32_t rank, int32_t world, int64_t n_loc,
                              int64_t n_ghost) {
    TC_CHECK(world >= 1 && world <= TC_MAX_WORLD, "world size");
    TC_CHECK(rank >= 0 && rank < world && n_loc > 0 && n_ghost >= 0, "rank / sizes");
    tc_dist* d = new tc_dist();
    memset(d->peer, 0, sizeof d->peer);
    memset(&d->K, 0, sizeof d->K);
    d->stream = (cudaStream_t)stream; d->rank = rank; d->world = world;
    d->n = n_loc; d->n_ghost = n_ghost;
    d->comm_bytes = off_mb2(n_ghost) + mb2_bytes();
    TC_CUDA(cudaMalloc(&d->comm, d->comm_bytes));
    TC_CUDA(cudaMemset(d->comm, 0, d->comm_bytes));          // tag 0 everywhere: sequence numbers start at 1
    const size_t nb = sizeof(double) * (size_t)n_loc;
    TC_CUDA(cudaMalloc(&d->y, nb)); TC_CUDA(cudaMalloc(&d->p0, nb)); TC_CUDA(cudaMalloc(&d->p1, nb));
    TC_CUDA(cudaMalloc(&d->x, nb)); TC_CUDA(cudaMalloc(&d->r, nb));
    TC_CUDA(cudaMalloc(&d->q, nb)); TC_CUDA(cudaMalloc(&d->dinv, nb));
    TC_CUDA(cudaMalloc(&d->zh, nb)); TC_CUDA(cudaMalloc(&d->pv, nb)); TC_CUDA(cudaMalloc(&d->cc, nb));
    TC_CUDA(cudaMalloc(&d->state, sizeof(int) * 16));
    TC_CUDA(cudaMemset(d->state, 0, sizeof(int) * 16));
    TC_CUDA(cudaMalloc(&d->prof, sizeof(unsigned long long) * 32));
    TC_CUDA(cudaMemset(d->prof, 0, sizeof(unsigned long long) * 32));
    TC_CUDA(cudaMalloc(&d->K.bar.ctr, sizeof(unsigned int) * TC_BAR_CTR_WORDS));
    TC_CUDA(cudaMemset(d->K.bar.ctr, 0, sizeof(unsigned int) * TC_BAR_CTR_WORDS));
    TC_CUDA(cudaMalloc(&d->K.bar.part, sizeof(double) * TC_BAR_PART_DOUBLES));
    TC_CUDA(cudaMalloc(&d->K.bar.res, sizeof(double) * 2 * TC_BAR_MAXK));
    TC_CUDA(cudaMemset(d->K.bar.res, 0, sizeof(double) * 2 * TC_BAR_MAXK));
    d->peer[rank] = d->comm;
    d->peer_ghost[rank] = n_ghost;
    TC_CUDA(cudaDeviceSynchronize());
    *out = d;
    return 0;
}

extern "C" int tc_dist_ipc_handle(tc_dist* d, void* handle64) {
    cudaIpcMemHandle_t h;
    TC_CUDA(cudaIpcGetMemHandle(&h, d->comm));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    memcpy(handle64, &h, 64);
    return 0;
}

// handles: world * 64 bytes; peer_ghost: ghost count of all ranks
extern "C" int tc_dist_open_peers(tc_dist* d, const void* handles, const int64_t* peer_ghost) {
    for (int w = 0; w < d->world; ++w) {
        d->peer_ghost[w] = peer_ghost[w];
        if (w == d->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * w, 64);
        void* p = nullptr;
        TC_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        d->peer[w] = (char*)p;
    }
    d->peers_open = true;
    return 0;
}

// Same-process variant (test harness: several "ranks" as sibling handles on ONE GPU, each on its own stream, so that a
// single-GPU box reaches the ghost side lists, send lists and mailboxes): the peers' buffers are plain device pointers.
// The kernels of the sibling ranks must be able to run concurrently (small grids); a rank that never sees its peer
// times out (TC_DIST_TIMEOUT_S) and reports -3 as usual.
extern "C" int tc_dist_open_peers_local(tc_dist* d, tc_dist* const* all, const int64_t* peer_ghost) {
    for (int w = 0; w < d->world; ++w) {
        d->peer_ghost[w] = peer_ghost[w];
        if (w == d->rank) continue;
        TC_CHECK(all[w] && all[w]->rank == w && all[w]->world == d->world && all[w]->comm, "sibling handle");
        d->peer[w] = all[w]->comm;
    }
    d->peers_open = true;
    d->peers_local = true;
    return 0;
}

// Halo plan: send list (local row -> peer, ghost slot of that peer) and the number of neighbour ranks.
// All arrays are HOST pointers; they are copied.
extern "C" int tc_dist_set_halo(tc_dist* d, int64_t n_send, const int32_t* send_row, const int32_t* send_peer,
                                const int64_t* send_dst, int32_t n_nbr, const int32_t* nbr) {
    TC_CHECK(n_nbr >= 0 && n_nbr < TC_MAX_WORLD, "neighbour count");
    (void)nbr;
    for (int64_t e = 0; e < n_send; ++e) {
        const int w = send_peer[e];
        TC_CHECK(w >= 0 && w < d->world && w != d->rank && d->peer[w], "send list names a peer that is not open");
        TC_CHECK(send_row[e] >= 0 && send_row[e] < d->n && send_dst[e] >= 0 && send_dst[e] < d->peer_ghost[w],
                 "send list entry out of range");
    }
    cudaFree(d->send_row); cudaFree(d->send_peer); cudaFree(d->send_dst);
    d->send_row = nullptr; d->send_peer = nullptr; d->send_dst = nullptr;
    if (n_send > 0) {
        TC_CUDA(cudaMalloc(&d->send_row, sizeof(int) * n_send));
        TC_CUDA(cudaMalloc(&d->send_peer, sizeof(int) * n_send));
        TC_CUDA(cudaMalloc(&d->send_dst, sizeof(long long) * n_send));
        TC_CUDA(cudaMemcpy(d->send_row, send_row, sizeof(int) * n_send, cudaMemcpyHostToDevice));
        TC_CUDA(cudaMemcpy(d->send_peer, send_peer, sizeof(int) * n_send, cudaMemcpyHostToDevice));
        TC_CUDA(cudaMemcpy(d->send_dst, send_dst, sizeof(long long) * n_send, cudaMemcpyHostToDevice));
    }
    d->h_send_row.assign(send_row, send_row + n_send);
    d->h_send_peer.assign(send_peer, send_peer + n_send);
    d->h_send_dst.assign(send_dst, send_dst + n_send);
    TcDistK& K = d->K;
    K.n_send = n_send; K.send_row = d->send_row; K.send_peer = d->send_peer; K.send_dst = d->send_dst;
    K.n_nbr = n_nbr;
    return 0;
}

// ---- pipelined kernel: launch shape and per-slice send lists ---------------------------------------------------
#define PIPE_FN(MB, T) (ef ? (void*)tc_dist_theta_pipe_kernel<MB, T, true> : (void*)tc_dist_theta_pipe_kernel<MB, T, false>)
static void* dist_pipe_fn(int shape, bool ef) {          // shape = 100 * blocks per SM + threads / 32
    switch (shape) {
        case 132: return PIPE_FN(1, 1024);               // 64 registers, 32 warps per SM
        case 232: return PIPE_FN(2, 1024);               // 32 registers, 64 warps per SM
        case 316: return PIPE_FN(3, 512);                // 40 registers, 48 warps per SM
        case 416: return PIPE_FN(4, 512);                // 32 registers
        default:  return PIPE_FN(2, 512);                // 216: 64 registers, 32 warps per SM
    }
}
static int pipe_shape_threads(int shape) { return 32 * (shape % 100); }
#define SR_FN(MB, T) (ef ? (void*)tc_dist_theta_sr_kernel<MB, T, true> : (void*)tc_dist_theta_sr_kernel<MB, T, false>)
static void* dist_sr_fn(int shape, bool ef) {
    switch (shape) {
        case 132: return SR_FN(1, 1024);
        case 232: return SR_FN(2, 1024);
        case 316: return SR_FN(3, 512);
        case 416: return SR_FN(4, 512);
        default:  return SR_FN(2, 512);
    }
}
// Send entries regrouped by the boundary slice of their row (sorted by lane): the warp that forms the slice pushes.
// A send row outside every boundary slice (possible only for a structurally non-symmetric pattern) switches the
// pipelined kernel off for this body (pipe_ok = false); the classic kernel does not need the grouping.
static int dist_build_slice_sends(tc_dist* d, const std::vector<int>& hb, int n_bnd) {
    cudaFree(d->sptr); cudaFree(d->s_lane); cudaFree(d->s_peer); cudaFree(d->s_dst);
    d->sptr = d->s_lane = d->s_peer = nullptr; d->s_dst = nullptr;
    d->pipe_ok = true;
    const size_t ns = d->h_send_row.size();
    std::vector<int> cnt(n_bnd + 1, 0);
    for (size_t e = 0; e < ns; ++e) {
        const int b = hb[d->h_send_row[e] >> 5];
        if (b < 0) { d->pipe_ok = false; return 0; }
        cnt[b + 1] += 1;
    }
    for (int b = 0; b < n_bnd; ++b) cnt[b + 1] += cnt[b];
    std::vector<int> fill(cnt.begin(), cnt.end() - 1), lane(ns ? ns : 1), peer(ns ? ns : 1);
    std::vector<long long> dst(ns ? ns : 1);
    for (size_t e = 0; e < ns; ++e) {                 // the send list is sorted by (peer, row): lanes ascend per peer
        const int b = hb[d->h_send_row[e] >> 5];
        const int k = fill[b]++;
        lane[k] = d->h_send_row[e] & 31; peer[k] = d->h_send_peer[e]; dst[k] = d->h_send_dst[e];
    }
    TC_CUDA(cudaMalloc(&d->sptr, sizeof(int) * (n_bnd + 1)));
    TC_CUDA(cudaMalloc(&d->s_lane, sizeof(int) * lane.size()));
    TC_CUDA(cudaMalloc(&d->s_peer, sizeof(int) * peer.size()));
    TC_CUDA(cudaMalloc(&d->s_dst, sizeof(long long) * dst.size()));
    TC_CUDA(cudaMemcpy(d->sptr, cnt.data(), sizeof(int) * (n_bnd + 1), cudaMemcpyHostToDevice));
    TC_CUDA(cudaMemcpy(d->s_lane, lane.data(), sizeof(int) * lane.size(), cudaMemcpyHostToDevice));
    TC_CUDA(cudaMemcpy(d->s_peer, peer.data(), sizeof(int) * peer.size(), cudaMemcpyHostToDevice));
    TC_CUDA(cudaMemcpy(d->s_dst, dst.data(), sizeof(long long) * dst.size(), cudaMemcpyHostToDevice));
    d->K.sptr = d->sptr; d->K.s_lane = d->s_lane; d->K.s_peer = d->s_peer; d->K.s_dst = d->s_dst;
    return 0;
}

// Operator of the local rows (device pointers, kept by reference): SELL-32 over LOCAL columns (col, and optionally
// col16 = column - row); couplings to ghosts as side lists of the n_bnd boundary slices (gptr/gslot/gval, SELL-32 per
// boundary slice; bidx[s] = index of slice s among the boundary slices, or -1).
extern "C" int tc_dist_set_operator(tc_dist* d, int32_t nslices, const int64_t* slice_ptr, const int32_t* col,
                                    const int16_t* col16, const double* val, const double* mass,
                                    const double* adiag, const double* b, int32_t n_bnd, const int32_t* bidx,
                                    const int64_t* gptr, const int32_t* gslot, const double* gval) {
    TC_CHECK(n_bnd >= 0 && n_bnd <= nslices && (n_bnd == 0 || (bidx && gptr && gslot && gval)),
             "boundary slices need their index map and the ghost coupling lists");
    TcDistK& K = d->K;
    K.A = TcSell{(int)d->n, nslices, (const long long*)slice_ptr, col, val, (const short*)col16};
    K.n = (int)d->n; K.n_ghost = (int)d->n_ghost; K.rank = d->rank; K.world = d->world;
    K.n_bnd = n_bnd; K.bidx = bidx;
    K.gptr = (const long long*)gptr; K.gslot = gslot; K.gval = gval;
    K.b = b; K.mass = mass; K.adiag = adiag;
    K.y = d->y; K.p[0] = d->p0; K.p[1] = d->p1;
    K.x = d->x; K.r = d->r; K.q = d->q; K.dinv = d->dinv;
    K.zh = d->zh; K.pv = d->pv; K.cc = d->cc;
    K.y_gl = (u64*)(d->comm + off_gl(d->n_ghost, 0));
    K.p_gl[0] = (u64*)(d->comm + off_gl(d->n_ghost, 1)); K.p_gl[1] = (u64*)(d->comm + off_gl(d->n_ghost, 2));
    K.my_mb = (u64*)(d->comm + off_mb(d->n_ghost));
    K.my_mb2 = (u64*)(d->comm + off_mb2(d->n_ghost));
    for (int w = 0; w < TC_MAX_WORLD; ++w) {
        K.peer_y_gl[w] = nullptr; K.peer_p_gl[0][w] = nullptr; K.peer_p_gl[1][w] = nullptr; K.peer_mb[w] = nullptr;
        K.peer_mb2[w] = nullptr;
    }
    for (int w = 0; w < d->world; ++w) {
        if (!d->peer[w]) continue;
        const long long g = d->peer_ghost[w];
        K.peer_y_gl[w] = (u64*)(d->peer[w] + off_gl(g, 0));
        K.peer_p_gl[0][w] = (u64*)(d->peer[w] + off_gl(g, 1)); K.peer_p_gl[1][w] = (u64*)(d->peer[w] + off_gl(g, 2));
        K.peer_mb[w] = (u64*)(d->peer[w] + off_mb(g));
        K.peer_mb2[w] = (u64*)(d->peer[w] + off_mb2(g));
    }
    K.state = d->state;
    K.prof = d->prof;
    cudaFree(d->bstash); d->bstash = nullptr;
    TC_CUDA(cudaMalloc(&d->bstash, sizeof(double) * 32 * (size_t)(n_bnd > 0 ? n_bnd : 1)));
    K.bstash = d->bstash;
    {
        const char* xe = getenv("TC_DIST_X_IN_DIR");                 // rows per GPU from which the x update is merged
        const long long thr = xe ? atoll(xe) : 4000000LL;
        K.x_in_dir = d->n >= thr ? 1 : 0;
    }
    int dev = 0, sms = 148, occ = 0, khz = 1965000;
    TC_CUDA(cudaGetDevice(&dev));
    TC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TC_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    const char* to = getenv("TC_DIST_TIMEOUT_S");
    double secs = to ? atof(to) : 20.0;
    if (!(secs > 0.0)) secs = 20.0;
    K.spin_limit = (long long)(secs * 1e3 * (double)khz);
    const char* pf = getenv("TC_DIST_PUSH_FENCE");
    K.push_fence = pf && pf[0] == '1';
    const char* dg = getenv("TC_DIST_DIAG");            // 1: count late ghosts in state[10]; 2: also do not wait for them
    if (dg) { const int v = atoi(dg); cudaMemcpy(d->state + 11, &v, sizeof(int), cudaMemcpyHostToDevice); }
    if (dg && atoi(dg) == 3 && !d->blkprof) {           // per-block phase profile (scripts/pipe_bench.py --blocks)
        TC_CUDA(cudaMalloc(&d->blkprof, sizeof(unsigned long long) * 10 * TC_MAX_PARTIALS));
        TC_CUDA(cudaMemset(d->blkprof, 0, sizeof(unsigned long long) * 10 * TC_MAX_PARTIALS));
    }
    K.blkprof = d->blkprof;
    TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)dist_fn(), dist_threads(), 0));
    if (occ > 2048 / dist_threads()) occ = 2048 / dist_threads();
    if (occ < 1) occ = 1;
    d->grid = occ * sms;
    if (d->grid > TC_MAX_PARTIALS) d->grid = TC_MAX_PARTIALS;
    // boundary slices cost about one slice more (the dependent loads of the ghost side list): leave that many warps
    // without a slice in the remainder round — the rotation below puts the boundary slices exactly on those warps
    const char* bs = getenv("TC_DIST_BND_SLACK");
    const double bnd_slack = bs ? atof(bs) : 1.0;
    const long long nsl_eff = nslices + (long long)(bnd_slack * n_bnd);
    d->grid = tc_balanced_grid(nsl_eff, dist_threads() / 32, d->grid);
    if (const char* fg = getenv("TC_DIST_GRID")) { const int g = atoi(fg); if (g > 0 && g <= d->grid + 64 && g <= occ * sms) d->grid = g; }
    // pipelined kernel: 2 x 512 threads per SM by default (64 registers for the fused recurrences)
    {
        const char* pt = getenv("TC_PIPE_SHAPE");        // 100 * blocks per SM + warps per block
        d->pipe_shape = pt ? atoi(pt) : 216;
        if (d->pipe_shape != 132 && d->pipe_shape != 232 && d->pipe_shape != 316 && d->pipe_shape != 416) d->pipe_shape = 216;
        d->pipe_threads = pipe_shape_threads(d->pipe_shape);
        const char* lh = getenv("TC_PIPE_L2_HINT");
        d->l2_hint = !(lh && lh[0] == '0');
        int pocc = 0;
        TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pocc, (const void*)dist_pipe_fn(d->pipe_shape, d->l2_hint),
                                                              d->pipe_threads, 0));
        if (pocc > d->pipe_shape / 100) pocc = d->pipe_shape / 100;
        if (pocc < 1) pocc = 1;
        d->pipe_grid = pocc * sms;
        if (d->pipe_grid > TC_MAX_PARTIALS) d->pipe_grid = TC_MAX_PARTIALS;
        d->pipe_grid = tc_balanced_grid(nslices + (long long)(bnd_slack * n_bnd), d->pipe_threads / 32, d->pipe_grid);
        int socc = 0;
        TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&socc, (const void*)dist_sr_fn(d->pipe_shape, d->l2_hint),
                                                              d->pipe_threads, 0));
        if (socc > d->pipe_shape / 100) socc = d->pipe_shape / 100;
        if (socc < 1) socc = 1;
        d->sr_grid = socc * sms;
        if (d->sr_grid > TC_MAX_PARTIALS) d->sr_grid = TC_MAX_PARTIALS;
        d->sr_grid = tc_balanced_grid(nslices + (long long)(bnd_slack * n_bnd), d->pipe_threads / 32, d->sr_grid);
    }
    K.rot = 0; d->rot_pipe = 0; d->rot_sr = 0;
    d->pipe_ok = true;
    K.sptr = nullptr; K.s_lane = nullptr; K.s_peer = nullptr; K.s_dst = nullptr;
    if (n_bnd == 0 && !d->h_send_row.empty()) d->pipe_ok = false;
    if (n_bnd > 0) {
        std::vector<int> hb(nslices);
        TC_CUDA(cudaMemcpy(hb.data(), bidx, sizeof(int) * nslices, cudaMemcpyDeviceToHost));
        if (dist_build_slice_sends(d, hb, n_bnd)) return -1;
        int nl = 0, nu = 0;
        while (nl < nslices && hb[nl] >= 0) ++nl;
        while (nu < nslices - nl && hb[nslices - 1 - nu] >= 0) ++nu;
        auto rot_for = [&](long long nwarps) -> int {
            if (!(nl + nu > 0 && nl + nu <= nwarps && nwarps <= nslices)) return 0;
            // first-round index range [nwarps - (nl + nu), nwarps) <-> cyclic slice run starting at nslices - nu
            long long rot = ((long long)(nslices - nu) - (nwarps - (nl + nu))) % nslices;
            if (rot < 0) rot += nslices;
            return (int)rot;
        };
        K.rot = rot_for((long long)d->grid * (dist_threads() / 32));
        d->rot_pipe = rot_for((long long)d->pipe_grid * (d->pipe_threads / 32));
        d->rot_sr = rot_for((long long)d->sr_grid * (d->pipe_threads / 32));
    }
    return 0;
}

static int dist_read_state(tc_dist* d, int* st16) {
    TC_CUDA(cudaStreamSynchronize(d->stream));
    TC_CUDA(cudaMemcpy(st16, d->state, sizeof(int) * 16, cudaMemcpyDeviceToHost));
    return 0;
}

// 0 when every step since the last reset ended cleanly.  -3: a wait on a peer timed out (communication
// failure: the results of this rank AND its peers are invalid); -4: a PCG solve hit the iteration limit
// or broke down.  The flags stay set until tc_dist_reset_error.
extern "C" int tc_dist_check(tc_dist* d) {
    int st[16];
    if (dist_read_state(d, st)) return -1;
    if (st[4]) { tc_set_error(__FILE__, __LINE__, "row-partitioned step: wait on a peer GPU timed out (TC_DIST_TIMEOUT_S)"); return -3; }
    if (st[8]) {
        tc_set_error(__FILE__, __LINE__, st[7] == 1 ? "row-partitioned PCG reached the iteration limit above the tolerance"
                                                    : "row-partitioned PCG broke down (non-finite / peer failure)");
        return -4;
    }
    return 0;
}
extern "C" int tc_dist_reset_error(tc_dist* d) {
    TC_CUDA(cudaStreamSynchronize(d->stream));
    TC_CUDA(cudaMemset(d->state + 4, 0, sizeof(int)));
    TC_CUDA(cudaMemset(d->state + 7, 0, 2 * sizeof(int)));
    return 0;
}

// Let the kernel work on a caller-owned state block (the FE block of a tc_solver state vector: tc_attach_dist)
extern "C" int tc_dist_set_state_ptr(tc_dist* d, double* y_local_dev) {
    d->K.y = y_local_dev ? y_local_dev : d->y;
    return 0;
}
extern "C" int tc_dist_set_state(tc_dist* d, const double* y_local_dev) {
    TC_CUDA(cudaMemcpyAsync(d->K.y, y_local_dev, sizeof(double) * d->n, cudaMemcpyDeviceToDevice, d->stream));
    TC_CUDA(cudaStreamSynchronize(d->stream));
    return 0;
}
extern "C" int tc_dist_get_state(tc_dist* d, double* y_local_dev) {
    TC_CUDA(cudaMemcpyAsync(y_local_dev, d->K.y, sizeof(double) * d->n, cudaMemcpyDeviceToDevice, d->stream));
    return tc_dist_check(d);
}
// enqueue-only copies for pipelined host<->device streaming: dir 0 = set (caller -> state), 1 = get
extern "C" int tc_dist_copy_state_async(tc_dist* d, double* y_local_dev, int32_t dir) {
    if (dir == 0)
        TC_CUDA(cudaMemcpyAsync(d->K.y, y_local_dev, sizeof(double) * d->n, cudaMemcpyDeviceToDevice, d->stream));
    else
        TC_CUDA(cudaMemcpyAsync(y_local_dev, d->K.y, sizeof(double) * d->n, cudaMemcpyDeviceToDevice, d->stream));
    return 0;
}
// Surface means of a partitioned body (thermca/solver.py:211-220: mean = C_surf^T x is a sum over ALL rows of the
// body): every rank has the partial sums of its local rows (tc_surf_mean_kernel, one partial per block of a job);
// this one-CTA kernel adds them up per job in block order, trades the per-rank sums through the LL mailboxes and
// leaves the rank-ordered total in the job's first partial (the others zeroed), so the network kernel's C_SURF_MEAN
// command reads the global mean — bit-identical on every rank, which keeps the replicated network kernels in lockstep.
__global__ void tc_dist_mean_exchange_kernel(TcDistK S, double* __restrict__ partials, const int* __restrict__ first_block,
                                             const int* __restrict__ nblk, int job0, int njobs) {
    __shared__ double s_loc[TC_MB2_VALUES];
    int* err = S.state + 4;
    const int seq = S.state[12] + 1;
    const int par = seq & 1;
    for (int j = threadIdx.x; j < njobs; j += blockDim.x) {
        double loc = 0.0;
        const int fb = first_block[job0 + j];
        for (int b = 0; b < nblk[job0 + j]; ++b) loc += partials[fb + b];
        s_loc[j] = loc;
    }
    __syncthreads();
    const bool bad = *((volatile int*)err) != 0;
    for (int it = threadIdx.x; it < S.world * njobs; it += blockDim.x) {
        const int w = it / njobs, j = it - w * njobs;
        ll_store(S.peer_mb2[w] + ((size_t)par * TC_MAX_WORLD + S.rank) * TC_MB2_SLOT + 2 * j, bad ? nan("") : s_loc[j],
                 (unsigned int)seq);
    }
    for (int j = threadIdx.x; j < njobs; j += blockDim.x) {
        double tot = 0.0;
        for (int w = 0; w < S.world; ++w)
            tot += ll_load(S.my_mb2 + ((size_t)par * TC_MAX_WORLD + w) * TC_MB2_SLOT + 2 * j, (unsigned int)seq, err,
                           S.spin_limit);
        const int fb = first_block[job0 + j];
        partials[fb] = tot;
        for (int b = 1; b < nblk[job0 + j]; ++b) partials[fb + b] = 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) S.state[12] = seq;
}
extern "C" int tc_dist_mean_exchange(tc_dist* d, double* partials_dev, const int32_t* first_block_dev,
                                     const int32_t* nblk_dev, int32_t job0, int32_t njobs) {
    TC_CHECK(njobs >= 0 && njobs <= TC_MB2_VALUES, "too many surface means in one exchange");
    if (njobs == 0 || d->world == 1) return 0;
    TC_CHECK(d->peers_open, "peers not open");
    tc_dist_mean_exchange_kernel<<<1, 256, 0, d->stream>>>(d->K, partials_dev, first_block_dev, nblk_dev, job0, njobs);
    return 0;
}
// ---- explicit stages of a partitioned body (the adaptive pairs RK45 / RK23 under torchrun) ---------------------------
// f = b - A v on the local rows for ANY stage vector v (thermca/solver.py:281).  The ghost rows of v travel LL encoded
// as in the fused theta kernel; the two p ghost buffers alternate with the halo sequence number: a rank can run at most
// one product ahead of a neighbour (its next product needs that neighbour's next push, which is stream-ordered behind
// the neighbour's current product), so two buffers are enough and no flow control is needed.  `out` holds b on entry.
// The sequence number is advanced by a one-thread kernel ahead of the product (stream / graph ordered), so a replayed
// CUDA graph needs no host argument.  Cooperative launch: all blocks must be resident, a block that waits for a ghost
// pushed by a not yet scheduled block of the peer (which waits for this rank's unscheduled blocks) would deadlock.
__global__ void tc_dist_seq_kernel(int* state, const int* __restrict__ done) {
    if (done && *done) return;
    state[1] += 1;
}
template <int MB, int THREADS>
__global__ void __launch_bounds__(THREADS, MB)
tc_dist_residual_kernel(TcDistK S, const double* __restrict__ v, double* out, const int* __restrict__ done) {
    if (done && *done) return;
    const unsigned int hseq = (unsigned int)S.state[1];
    const int pb = (int)(hseq & 1u);
    const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
    for (long long e = gtid; e < S.n_send; e += gsz)
        ll_store(S.peer_p_gl[pb][S.send_peer[e]] + 2 * S.send_dst[e], v[S.send_row[e]], hseq);
    slab_spmv<SPMV_B>(S, v, S.p_gl[pb], out, out, 0.0, hseq);
}
extern "C" int tc_dist_residual(tc_dist* d, const double* v_local_dev, double* out_local_dev, const int32_t* done_flag_dev) {
    TC_CHECK(d->K.A.val && d->grid > 0, "operator not set");
    TC_CHECK(d->world == 1 || d->peers_open, "peers not open");
    tc_dist_seq_kernel<<<1, 1, 0, d->stream>>>(d->state, done_flag_dev);
    void* fn = dist_threads() == 512 ? (void*)tc_dist_residual_kernel<4, 512>
               : dist_threads() == 256 ? (void*)tc_dist_residual_kernel<8, 256> : (void*)tc_dist_residual_kernel<2, 1024>;
    void* args[] = {(void*)&d->K, (void*)&v_local_dev, (void*)&out_local_dev, (void*)&done_flag_dev};
    TC_CUDA(cudaLaunchCooperativeKernel(fn, dim3(d->grid), dim3(dist_threads()), args, 0, d->stream));
    return 0;
}

// vals[j] <- sum over the ranks of vals[j], added in rank order (bit-identical on every rank): the error norm and the
// initial-step norms of the adaptive pairs (scipy rk.py:145-165, common.py:63-134 are means over ALL rows).  Shares the
// mailbox and the sequence counter of the surface-mean exchange.
template <int OP>      // 0: sum, 1: max
__global__ void tc_dist_sum_exchange_kernel(TcDistK S, double* __restrict__ vals, int k, const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ double s_loc[TC_MB2_VALUES];
    int* err = S.state + 4;
    const int seq = S.state[12] + 1;
    const int par = seq & 1;
    const bool bad = *((volatile int*)err) != 0;
    for (int j = threadIdx.x; j < k; j += blockDim.x) s_loc[j] = vals[j];
    __syncthreads();
    for (int it = threadIdx.x; it < S.world * k; it += blockDim.x) {
        const int w = it / k, j = it - w * k;
        ll_store(S.peer_mb2[w] + ((size_t)par * TC_MAX_WORLD + S.rank) * TC_MB2_SLOT + 2 * j, bad ? nan("") : s_loc[j],
                 (unsigned int)seq);
    }
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        double tot = 0.0;
        for (int w = 0; w < S.world; ++w) {
            const double v = ll_load(S.my_mb2 + ((size_t)par * TC_MAX_WORLD + w) * TC_MB2_SLOT + 2 * j, (unsigned int)seq,
                                     err, S.spin_limit);
            tot = OP == 0 ? tot + v : (w == 0 || v > tot || v != v ? v : tot);      // NaN (a failed rank) wins the max
        }
        vals[j] = tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) S.state[12] = seq;
}
extern "C" int tc_dist_sum_exchange(tc_dist* d, double* vals_dev, int32_t k, const int32_t* done_flag_dev) {
    TC_CHECK(k >= 0 && k <= TC_MB2_VALUES, "too many values in one exchange");
    if (k == 0 || d->world == 1) return 0;
    TC_CHECK(d->peers_open, "peers not open");
    tc_dist_sum_exchange_kernel<0><<<1, 64, 0, d->stream>>>(d->K, vals_dev, k, done_flag_dev);
    return 0;
}
extern "C" int tc_dist_max_exchange(tc_dist* d, double* vals_dev, int32_t k, const int32_t* done_flag_dev) {
    TC_CHECK(k >= 0 && k <= TC_MB2_VALUES, "too many values in one exchange");
    if (k == 0 || d->world == 1) return 0;
    TC_CHECK(d->peers_open, "peers not open");
    tc_dist_sum_exchange_kernel<1><<<1, 64, 0, d->stream>>>(d->K, vals_dev, k, done_flag_dev);
    return 0;
}
extern "C" int32_t tc_dist_rank(tc_dist* d) { return d->rank; }
extern "C" int64_t tc_dist_local_rows(tc_dist* d) { return d->n; }
extern "C" void* tc_dist_stream(tc_dist* d) { return (void*)d->stream; }

// load vector of the local rows (device pointer, kept by reference): time-varying boundary data
extern "C" int tc_dist_set_load(tc_dist* d, const double* b_local_dev) {
    d->K.b = b_local_dev;
    return 0;
}

extern "C" int tc_dist_theta_steps(tc_dist* d, int32_t nsteps, double theta, double h, double rtol, int32_t maxit) {
    TC_CHECK(d->K.A.val && d->grid > 0, "operator not set");
    TC_CHECK(d->world == 1 || d->peers_open, "peers not open");
    d->K.theta = theta; d->K.h = h; d->K.rtol = rtol; d->K.maxit = maxit;
    for (int k = 0; k < nsteps; ++k) {
        void* args[] = {(void*)&d->K};
        if (d->pcg_mode == 1) {
            TcDistK Kp = d->K;
            Kp.rot = d->rot_pipe;
            void* pargs[] = {(void*)&Kp};
            TC_CUDA(cudaLaunchCooperativeKernel(dist_pipe_fn(d->pipe_shape, d->l2_hint), dim3(d->pipe_grid),
                                                dim3(d->pipe_threads), pargs, 0, d->stream));
        } else if (d->pcg_mode == 2) {
            TcDistK Kp = d->K;
            Kp.rot = d->rot_sr;
            void* pargs[] = {(void*)&Kp};
            TC_CUDA(cudaLaunchCooperativeKernel(dist_sr_fn(d->pipe_shape, d->l2_hint), dim3(d->sr_grid),
                                                dim3(d->pipe_threads), pargs, 0, d->stream));
        } else
            TC_CUDA(cudaLaunchCooperativeKernel(dist_fn(), dim3(d->grid), dim3(dist_threads()), args, 0, d->stream));
    }
    return 0;
}

// PCG formulation of tc_dist_theta_steps: 0 = classic (two reductions and a barrier per iteration), 1 = pipelined
// (one fused sweep and one reduction per iteration), 2 = single-reduction classic (sweep, ONE all-reduce of seven sums,
// merged vector phase, local barrier).
// Must be the same on every rank.  Returns -2 when the pipelined kernel cannot serve this body.
extern "C" int tc_dist_set_pcg(tc_dist* d, int32_t mode) {
    TC_CHECK(mode >= 0 && mode <= 2, "pcg mode");
    TC_CHECK(mode == 0 || (d->K.A.val && d->pipe_ok), "single-reduction / pipelined PCG: operator not set or a send row lies outside the boundary slices");
    d->pcg_mode = mode;
    return 0;
}
extern "C" int32_t tc_dist_get_pcg(tc_dist* d) { return d->pcg_mode; }

// stage solve of a W-method on the partitioned body: z = (I + coef * h A)^-1 rhs on the local rows, h from the device
extern "C" int tc_dist_solve(tc_dist* d, double coef, const double* h_dev, const double* rhs_local_dev, double* z_local_dev,
                             double rtol, int32_t maxit, const int32_t* skip_flag_dev, const int32_t* done_flag_dev) {
    TC_CHECK(d->K.A.val && d->grid > 0, "operator not set");
    TC_CHECK(d->world == 1 || d->peers_open, "peers not open");
    TcDistK K = d->K;
    K.theta = coef; K.h = 0.0; K.h_dev = h_dev; K.rtol = rtol; K.maxit = maxit;
    K.rhs_in = rhs_local_dev; K.x_out = z_local_dev; K.done = done_flag_dev; K.skip = skip_flag_dev;
    void* fn = dist_threads() == 512 ? (void*)tc_dist_theta_kernel<4, 512, true>
               : dist_threads() == 256 ? (void*)tc_dist_theta_kernel<8, 256, true> : (void*)tc_dist_theta_kernel<2, 1024, true>;
    void* args[] = {(void*)&K};
    TC_CUDA(cudaLaunchCooperativeKernel(fn, dim3(d->grid), dim3(dist_threads()), args, 0, d->stream));
    return 0;
}

// stats: iters_last, iters_total, comm error flag, solves (+ grid * 1e6), failed solves, late ghosts (TC_DIST_DIAG)
extern "C" int tc_dist_stats(tc_dist* d, int64_t* stats5) {
    int st[16];
    if (dist_read_state(d, st)) return -1;
    stats5[0] = st[2]; stats5[1] = st[3]; stats5[2] = st[4]; stats5[3] = st[5] + 1000000LL * (d->pcg_mode == 1 ? d->pipe_grid : d->pcg_mode == 2 ? d->sr_grid : d->grid); stats5[4] = st[8];
    stats5[5] = st[10];
    return 0;
}

// per-block phase profile (TC_DIST_DIAG=3): out[block * 10 + phase] accumulated ns (slot 8: SM id), cleared on read; returns the grid
extern "C" int tc_dist_block_profile(tc_dist* d, uint64_t* out, int32_t max_blocks) {
    TC_CHECK(d->blkprof, "per-block profile not enabled (TC_DIST_DIAG=3)");
    TC_CUDA(cudaStreamSynchronize(d->stream));
    const int g = d->pcg_mode == 1 ? d->pipe_grid : d->pcg_mode == 2 ? d->sr_grid : d->grid;
    const int nb = g < max_blocks ? g : max_blocks;
    TC_CUDA(cudaMemcpy(out, d->blkprof, sizeof(unsigned long long) * 10 * nb, cudaMemcpyDeviceToHost));
    TC_CUDA(cudaMemset(d->blkprof, 0, sizeof(unsigned long long) * 10 * TC_MAX_PARTIALS));
    return g;
}

// phase profile: 32 accumulated nanosecond counters (block 0: [0..15], last block: [16..31]); cleared on read
extern "C" int tc_dist_profile(tc_dist* d, uint64_t* out32) {
    TC_CUDA(cudaStreamSynchronize(d->stream));
    TC_CUDA(cudaMemcpy(out32, d->prof, sizeof(unsigned long long) * 32, cudaMemcpyDeviceToHost));
    TC_CUDA(cudaMemset(d->prof, 0, sizeof(unsigned long long) * 32));
    return 0;
}

// Teardown in two halves with a host barrier between them (CUDA IPC: the exporter must not free its
// buffer while an importer still has it mapped): every rank first drains its stream and closes its
// mappings of the peers' buffers, the caller runs a barrier, then tc_dist_destroy frees.
extern "C" int tc_dist_close_peers(tc_dist* d) {
    if (!d) return 0;
    cudaStreamSynchronize(d->stream);
    for (int w = 0; w < d->world; ++w)
        if (w != d->rank && d->peer[w]) { if (!d->peers_local) cudaIpcCloseMemHandle(d->peer[w]); d->peer[w] = nullptr; }
    d->peers_open = false;
    return 0;
}
extern "C" int tc_dist_destroy(tc_dist* d) {
    if (!d) return 0;
    tc_dist_close_peers(d);
    cudaFree(d->comm); cudaFree(d->y); cudaFree(d->p0); cudaFree(d->p1); cudaFree(d->x); cudaFree(d->r); cudaFree(d->q); cudaFree(d->dinv);
    cudaFree(d->state); cudaFree(d->prof); cudaFree(d->blkprof);
    cudaFree(d->K.bar.ctr); cudaFree(d->K.bar.part); cudaFree(d->K.bar.res);
    cudaFree(d->send_row); cudaFree(d->send_peer); cudaFree(d->send_dst);
    cudaFree(d->zh); cudaFree(d->pv); cudaFree(d->cc); cudaFree(d->bstash);
    cudaFree(d->sptr); cudaFree(d->s_lane); cudaFree(d->s_peer); cudaFree(d->s_dst);
    delete d;
    return 0;
}
