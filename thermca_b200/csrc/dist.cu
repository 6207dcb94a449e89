// Row-partitioned implicit step for ONE large FE body across GPUs (BASELINE config #5):
// fused compute + collective.  Each rank owns a block of operator rows and runs ONE cooperative
// persistent kernel per theta step.  Inside that kernel
//   * ghost values of the gathered vector are written straight into the peers' HBM with peer
//     stores over NVLink (no NCCL call, no host round trip), driven by send lists, so any
//     partition works (contiguous slabs of a banded ordering give rank +-1 neighbours, a general
//     partition any neighbour set);
//   * every PCG reduction is folded into a grid barrier (gridbar.cuh) whose last-arriving block
//     also trades the local sums with all peers through per-rank mailboxes and adds them in rank
//     order: one barrier == one global all-reduce, bit-identical scalars on all ranks ->
//     identical convergence decisions, deterministic results;
//   * the search direction is double buffered, so the rows a peer needs are recomputed from
//     stable inputs by the pushing threads (no intra-phase race, no extra barrier) and the push
//     overlaps the vector phase; the product runs interior slices first and waits for the
//     neighbours' flags only before the few boundary slices.
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
#include "../../include/thermca_b200.h"
#include "parts.cuh"
#include "gridbar.cuh"

#define TC_MAX_WORLD 16
#define TC_DIST_K 3                     // values per all-reduce

struct TcDistK {
    TcSell A;                             // local rows; col16 relative to the row (interior slices), col = index
                                          // into the extended vector [n local | n_ghost ghosts] (boundary slices)
    int n, n_ghost, rank, world;
    int n_bnd;                            // boundary slices (reference at least one ghost column)
    const int* order;                     // nslices: interior slices first, boundary slices last
    const double *b, *mass, *adiag;
    double *y_ext, *p_ext[2];             // [n | n_ghost] each, inside the exported buffer
    double *x, *r, *q, *dinv;
    long long n_send;                     // halo send list: local row -> (peer, slot in the peer's extended vector)
    const int* send_row; const int* send_peer; const long long* send_dst;
    double* peer_y[TC_MAX_WORLD]; double* peer_p[2][TC_MAX_WORLD];
    int* peer_hflag[TC_MAX_WORLD];        // peers' halo flag arrays (I write entry [rank])
    int n_nbr; int nbr[TC_MAX_WORLD];     // ranks I trade ghosts with
    int* my_hflag;                        // [TC_MAX_WORLD], entry w written by rank w
    double* peer_mb[TC_MAX_WORLD]; int* peer_mbf[TC_MAX_WORLD];
    double* my_mb; int* my_mbf;           // mb[2][TC_DIST_K][MAXW], mbf[2][MAXW]
    TcGridBar bar;
    int* state;                           // [0] reduction seq, [1] halo seq, [2] iters last, [3] iters total,
                                          // [4] comm error, [5] solves, [7] status of the last failed solve, [8] failed solves
    double theta, h, rtol;
    int maxit;
    long long spin_limit;                 // clock64 ticks
    unsigned long long* prof;             // [2][16] phase times in ns: block 0 and the last block
};

__device__ __forceinline__ int ld_relaxed_sys(const int* p) {
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded wait; on timeout (or when another thread of this rank already gave up) raises the error flag.
// Relaxed polls + ONE acquire fence at the end: an acquire load per poll would invalidate this SM's L1 on every
// iteration while co-resident warps are still gathering through it.
__device__ __forceinline__ void spin_until(const int* flag, int seq, int* err, long long limit) {
    if (ld_relaxed_sys(flag) < seq) {
        const long long t0 = clock64();
        while (ld_relaxed_sys(flag) < seq) {
            if (*((volatile int*)err)) return;
            if (clock64() - t0 > limit) { *((volatile int*)err) = 1; __threadfence(); return; }
        }
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
}

// All-reduce of the local totals, run by the last block to reach the grid barrier (gridbar.cuh hook).
// Thread w < world writes my values into peer w's mailbox (parity double buffered, sequence tagged) and
// waits for peer w's values in mine; thread 0 then adds the world's values in rank order.
struct TcDistExchange {
    const TcDistK& S;
    int seq;
    __device__ __forceinline__ void operator()(double* tot, int K) const {
        if (S.world == 1) return;
        const int par = seq & 1;
        int* err = S.state + 4;
        if (threadIdx.x < S.world) {
            const int peer = threadIdx.x;
            const bool bad = *((volatile int*)err) != 0;
            volatile double* mb = S.peer_mb[peer];
            for (int k = 0; k < K; ++k) mb[(par * TC_DIST_K + k) * TC_MAX_WORLD + S.rank] = bad ? nan("") : tot[k];
            st_release_sys(S.peer_mbf[peer] + par * TC_MAX_WORLD + S.rank, seq);     // orders the stores above
            spin_until(S.my_mbf + par * TC_MAX_WORLD + peer, seq, err, S.spin_limit);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const bool bad = *((volatile int*)err) != 0;
            volatile double* mine = S.my_mb;
            for (int k = 0; k < K; ++k) {
                double s = 0.0;
                for (int w = 0; w < S.world; ++w) s += mine[(par * TC_DIST_K + k) * TC_MAX_WORLD + w];
                tot[k] = bad ? nan("") : s;
            }
        }
        __syncthreads();
    }
};

#define TC_DIST_THREADS 256

// Local rows times the extended vector, interior slices first (16-bit relative or 32-bit local columns, gathered
// through L1) and the boundary slices LAST in every warp's grid-stride sequence: a warp waits for the neighbours'
// halo flags only when it reaches its first boundary slice (by then the interior part has hidden the NVLink
// latency) and reads the gathered vector of boundary slices through L2 (ld.global.cg: ghosts were written by
// another GPU, the local L1 may hold stale lines).
template <int MODE>
__device__ __forceinline__ double slab_spmv(const TcDistK& S, const double* xe, double* y, const double* b,
                                            double shift, int hseq) {
    const TcSell& A = S.A;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int warp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_block;
    const int nsl = A.nslices, n_int = nsl - S.n_bnd;
    bool waited = false;
    double dot = 0.0;
    for (int idx = warp; idx < nsl; idx += nwarps) {
        const bool boundary = idx >= n_int;
        const int s = S.order ? S.order[idx] : idx;
        if (boundary && !waited) {
            if (lane == 0)
                for (int k = 0; k < S.n_nbr; ++k) spin_until(S.my_hflag + S.nbr[k], hseq, S.state + 4, S.spin_limit);
            __syncwarp();
            waited = true;
        }
        const long long base = A.slice_ptr[s];
        const int width = (int)((A.slice_ptr[s + 1] - base) >> 5);
        const double* v = A.val + base + lane;
        double sum = 0.0;
        if (boundary) {
            const int* c = A.col + base + lane;
            for (int k = 0; k < width; ++k)
                sum = __dadd_rn(sum, __dmul_rn(ld_stream_f64(v + 32 * k), tc_ld_x<2>(xe + ld_stream_s32(c + 32 * k))));
        } else if (A.col16) {
            const short* c = A.col16 + base + lane;
            const double* xr0 = xe + (s * 32 + lane);
            int k = 0;
            for (; k + 4 <= width; k += 4) {
                const double v0 = ld_stream_f64(v + 32 * (k + 0)), v1 = ld_stream_f64(v + 32 * (k + 1));
                const double v2 = ld_stream_f64(v + 32 * (k + 2)), v3 = ld_stream_f64(v + 32 * (k + 3));
                const int c0 = ld_stream_s16(c + 32 * (k + 0)), c1 = ld_stream_s16(c + 32 * (k + 1));
                const int c2 = ld_stream_s16(c + 32 * (k + 2)), c3 = ld_stream_s16(c + 32 * (k + 3));
                const double x0 = tc_ld_x<0>(xr0 + c0), x1 = tc_ld_x<0>(xr0 + c1), x2 = tc_ld_x<0>(xr0 + c2), x3 = tc_ld_x<0>(xr0 + c3);
                sum = __dadd_rn(sum, __dmul_rn(v0, x0));
                sum = __dadd_rn(sum, __dmul_rn(v1, x1));
                sum = __dadd_rn(sum, __dmul_rn(v2, x2));
                sum = __dadd_rn(sum, __dmul_rn(v3, x3));
            }
            for (; k < width; ++k)
                sum = __dadd_rn(sum, __dmul_rn(ld_stream_f64(v + 32 * k), tc_ld_x<0>(xr0 + ld_stream_s16(c + 32 * k))));
        } else {
            const int* c = A.col + base + lane;
            int k = 0;
            for (; k + 4 <= width; k += 4) {
                const double v0 = ld_stream_f64(v + 32 * (k + 0)), v1 = ld_stream_f64(v + 32 * (k + 1));
                const double v2 = ld_stream_f64(v + 32 * (k + 2)), v3 = ld_stream_f64(v + 32 * (k + 3));
                const int c0 = ld_stream_s32(c + 32 * (k + 0)), c1 = ld_stream_s32(c + 32 * (k + 1));
                const int c2 = ld_stream_s32(c + 32 * (k + 2)), c3 = ld_stream_s32(c + 32 * (k + 3));
                const double x0 = tc_ld_x<0>(xe + c0), x1 = tc_ld_x<0>(xe + c1), x2 = tc_ld_x<0>(xe + c2), x3 = tc_ld_x<0>(xe + c3);
                sum = __dadd_rn(sum, __dmul_rn(v0, x0));
                sum = __dadd_rn(sum, __dmul_rn(v1, x1));
                sum = __dadd_rn(sum, __dmul_rn(v2, x2));
                sum = __dadd_rn(sum, __dmul_rn(v3, x3));
            }
            for (; k < width; ++k)
                sum = __dadd_rn(sum, __dmul_rn(ld_stream_f64(v + 32 * k), tc_ld_x<0>(xe + ld_stream_s32(c + 32 * k))));
        }
        const int row = s * 32 + lane;
        if (row < A.n) {
            if (MODE == SPMV_B) {
                y[row] = __dsub_rn(b[row], sum);
            } else {
                const double xr = tc_ld_x<0>(xe + row);
                const double q = S.mass[row] * (xr + shift * sum);
                y[row] = q;
                dot += xr * q;
            }
        }
    }
    return dot;
}

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// phase profiler: thread 0 of block 0 and of the last block accumulate wall time per phase
#define PROF_INIT() const int prof_row = blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x - 1 ? 1 : -1); \
    unsigned long long prof_t = (prof_row >= 0 && threadIdx.x == 0) ? gtime() : 0ULL
#define PROF(k) do { if (prof_row >= 0 && threadIdx.x == 0) { const unsigned long long t_ = gtime(); \
    S.prof[prof_row * 16 + (k)] += t_ - prof_t; prof_t = t_; } } while (0)

// after a barrier that ordered this rank's ghost pushes: tell every neighbour that halo `hseq` has landed
__device__ __forceinline__ void raise_halo_flags(const TcDistK& S, int hseq) {
    if (blockIdx.x == 0 && (int)threadIdx.x < S.n_nbr)
        st_release_sys(S.peer_hflag[S.nbr[threadIdx.x]] + S.rank, hseq);
}

// scaled right-hand side and Jacobi factor of one row (explicit roundings: the owner's loop and the halo push
// must produce the same bits)
__device__ __forceinline__ void init_row(const TcDistK& S, int i, double shift, double& bsc, double& d) {
    const double m = S.mass[i];
    bsc = __dmul_rn(m, S.q[i]);
    d = __ddiv_rn(1.0, __dmul_rn(m, __fma_rn(shift, S.adiag[i], 1.0)));
}

template <int MB, int THREADS = TC_DIST_THREADS>
__global__ void __launch_bounds__(THREADS, MB)
tc_dist_theta_kernel(TcDistK S) {
    __shared__ double sh[32];
    const int n = S.n;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    unsigned int gen = 0u, tgen = tc_bar_generation(S.bar);
    int rseq = S.state[0], hseq = S.state[1];
    PROF_INIT();
    const double shift = S.theta * S.h;
    double* y_loc = S.y_ext;
    int pb = 0;

    // ---- right-hand side f = b - A y (thermca/solver.py:281) on the local rows -----------------
    if (S.n_nbr > 0) {
        for (long long e = gtid; e < S.n_send; e += gsz) S.peer_y[S.send_peer[e]][S.send_dst[e]] = y_loc[S.send_row[e]];
        __threadfence_system();
        tc_grid_barrier(S.bar, gen);
        raise_halo_flags(S, ++hseq);
    }
    slab_spmv<SPMV_B>(S, S.y_ext, S.q, S.b, 0.0, hseq);
    tc_grid_barrier(S.bar, gen);
    double red[TC_DIST_K];
    {
        double* p0 = S.p_ext[0];
        double rho_l = 0.0, bb_l = 0.0;
        for (int i = gtid; i < n; i += gsz) {
            double bsc, d;
            init_row(S, i, shift, bsc, d);
            const double z = __dmul_rn(d, bsc);
            S.r[i] = bsc; S.x[i] = 0.0; S.dinv[i] = d; p0[i] = z;
            rho_l += bsc * z; bb_l += bsc * bsc;
        }
        // rows a peer needs: recomputed from the same stable inputs (bit-identical), pushed over NVLink
        for (long long e = gtid; e < S.n_send; e += gsz) {
            double bsc, d;
            init_row(S, S.send_row[e], shift, bsc, d);
            S.peer_p[0][S.send_peer[e]][S.send_dst[e]] = __dmul_rn(d, bsc);
        }
        if (S.n_send > 0) __threadfence_system();
        red[0] = block_sum(rho_l, sh); red[1] = block_sum(bb_l, sh); red[2] = 0.0;
    }
    tc_grid_reduce_last<2>(S.bar, tgen, red, TcDistExchange{S, ++rseq});
    double rho = red[0];
    const double bb = red[1];
    const double tol2 = S.rtol * S.rtol * bb;
    int it = 0, status = 0;
    if (bb > 0.0) {
        status = 1;
        for (; it < S.maxit;) {
            PROF(0);
            raise_halo_flags(S, ++hseq);          // ghost pushes of p[pb] were fenced before the last barrier
            const double* pe = S.p_ext[pb];
            const double dot = slab_spmv<SPMV_SHIFT>(S, pe, S.q, nullptr, shift, hseq);
            PROF(1);
            red[0] = block_sum(dot, sh);
            PROF(2);
            tc_grid_reduce_last<1>(S.bar, tgen, red, TcDistExchange{S, ++rseq});
            PROF(3);
            const double pq = red[0];
            if (!(pq > 0.0) || !isfinite(pq)) { status = 2; break; }       // uniform over the grid AND over the ranks
            const double alpha = rho / pq;
            double rn = 0.0, rr = 0.0;
            for (int i = gtid; i < n; i += gsz) {
                S.x[i] += alpha * pe[i];
                const double ri = S.r[i] - alpha * S.q[i];
                S.r[i] = ri;
                rn += ri * (S.dinv[i] * ri);
                rr += ri * ri;
            }
            red[0] = block_sum(rn, sh); red[1] = block_sum(rr, sh);
            PROF(4);
            tc_grid_reduce_last<2>(S.bar, tgen, red, TcDistExchange{S, ++rseq});
            PROF(5);
            ++it;
            if (!(red[1] > tol2)) { status = isfinite(red[1]) ? 0 : 2; break; }
            const double beta = red[0] / rho;
            double* pn = S.p_ext[pb ^ 1];
            for (int i = gtid; i < n; i += gsz) pn[i] = __fma_rn(beta, pe[i], __dmul_rn(S.dinv[i], S.r[i]));
            for (long long e = gtid; e < S.n_send; e += gsz) {
                const int i = S.send_row[e];
                S.peer_p[pb ^ 1][S.send_peer[e]][S.send_dst[e]] = __fma_rn(beta, pe[i], __dmul_rn(S.dinv[i], S.r[i]));
            }
            if (S.n_send > 0) __threadfence_system();
            rho = red[0];
            pb ^= 1;
            PROF(6);
            tc_grid_barrier(S.bar, gen);
            PROF(7);
        }
    }
    // ---- y += h * z (not committed after a communication failure / breakdown) ------------------
    if (status != 2)
        for (int i = gtid; i < n; i += gsz) y_loc[i] += S.h * S.x[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S.state[0] = rseq; S.state[1] = hseq; S.state[2] = it; S.state[3] += it; S.state[5] += 1;
        if (status) { S.state[7] = status; S.state[8] += 1; }
    }
}

// ---- host ------------------------------------------------------------------------------------
static int dist_threads() {
    static int t = -1;
    if (t < 0) {
        const char* e = getenv("TC_PCG_THREADS");
        t = e ? atoi(e) : TC_DIST_THREADS;
        if (t != 256 && t != 512 && t != 1024) t = TC_DIST_THREADS;
    }
    return t;
}
static void* dist_fn() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TC_PCG_OCC");
        v = e ? atoi(e) : 8;
    }
    if (dist_threads() == 1024) return (void*)tc_dist_theta_kernel<2, 1024>;
    if (dist_threads() == 512) return (void*)tc_dist_theta_kernel<4, 512>;
    switch (v) {
        case 5: return (void*)tc_dist_theta_kernel<5>;
        case 6: return (void*)tc_dist_theta_kernel<6>;
        default: return (void*)tc_dist_theta_kernel<8>;
    }
}

struct tc_dist {
    cudaStream_t stream;
    int rank, world;
    long long n, n_ghost;
    char* comm = nullptr;             // my exported buffer
    size_t comm_bytes = 0;
    char* peer[TC_MAX_WORLD];
    long long peer_ext[TC_MAX_WORLD]; // n + n_ghost of every rank (fixes the layout of its buffer)
    double *x = nullptr, *r = nullptr, *q = nullptr, *dinv = nullptr;
    int* state = nullptr;
    unsigned long long* prof = nullptr;
    int* send_row = nullptr; int* send_peer = nullptr; long long* send_dst = nullptr;
    TcDistK K;
    int grid = 0;
    bool peers_open = false;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
// layout of a rank's exported buffer: [y_ext | p_ext0 | p_ext1 | mailbox | mailbox flags | halo flags]
static size_t ext_bytes(long long ext) { return align_up((size_t)ext * sizeof(double), 256); }
static size_t off_p(long long ext, int k) { return (size_t)(1 + k) * ext_bytes(ext); }
static size_t off_mb(long long ext) { return 3 * ext_bytes(ext); }
static size_t off_mbf(long long ext) { return off_mb(ext) + align_up(sizeof(double) * 2 * TC_DIST_K * TC_MAX_WORLD, 256); }
static size_t off_hf(long long ext) { return off_mbf(ext) + align_up(sizeof(int) * 2 * TC_MAX_WORLD, 256); }

extern "C" int tc_dist_create(tc_dist** out, void* stream, int32_t rank, int32_t world, int64_t n_loc,
                              int64_t n_ghost) {
    TC_CHECK(world >= 1 && world <= TC_MAX_WORLD, "world size");
    TC_CHECK(rank >= 0 && rank < world && n_loc > 0 && n_ghost >= 0, "rank / sizes");
    tc_dist* d = new tc_dist();
    memset(d->peer, 0, sizeof d->peer);
    memset(&d->K, 0, sizeof d->K);
    d->stream = (cudaStream_t)stream; d->rank = rank; d->world = world;
    d->n = n_loc; d->n_ghost = n_ghost;
    const long long ext = n_loc + n_ghost;
    d->comm_bytes = off_hf(ext) + align_up(sizeof(int) * TC_MAX_WORLD, 256);
    TC_CUDA(cudaMalloc(&d->comm, d->comm_bytes));
    TC_CUDA(cudaMemset(d->comm, 0, d->comm_bytes));
    const size_t nb = sizeof(double) * (size_t)n_loc;
    TC_CUDA(cudaMalloc(&d->x, nb)); TC_CUDA(cudaMalloc(&d->r, nb));
    TC_CUDA(cudaMalloc(&d->q, nb)); TC_CUDA(cudaMalloc(&d->dinv, nb));
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
    d->peer_ext[rank] = ext;
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

// handles: world * 64 bytes; peer_ext: n + n_ghost of all ranks
extern "C" int tc_dist_open_peers(tc_dist* d, const void* handles, const int64_t* peer_ext) {
    for (int w = 0; w < d->world; ++w) {
        d->peer_ext[w] = peer_ext[w];
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

// Halo plan: send list (local row -> peer, slot in the peer's extended vector) and the neighbour set.
// All arrays are HOST pointers; they are copied.
extern "C" int tc_dist_set_halo(tc_dist* d, int64_t n_send, const int32_t* send_row, const int32_t* send_peer,
                                const int64_t* send_dst, int32_t n_nbr, const int32_t* nbr) {
    TC_CHECK(n_nbr >= 0 && n_nbr < TC_MAX_WORLD, "neighbour count");
    for (int64_t e = 0; e < n_send; ++e) {
        const int w = send_peer[e];
        TC_CHECK(w >= 0 && w < d->world && w != d->rank && d->peer[w], "send list names a peer that is not open");
        TC_CHECK(send_row[e] >= 0 && send_row[e] < d->n && send_dst[e] >= 0 && send_dst[e] < d->peer_ext[w],
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
    TcDistK& K = d->K;
    K.n_send = n_send; K.send_row = d->send_row; K.send_peer = d->send_peer; K.send_dst = d->send_dst;
    K.n_nbr = n_nbr;
    for (int k = 0; k < n_nbr; ++k) {
        TC_CHECK(nbr[k] >= 0 && nbr[k] < d->world && nbr[k] != d->rank, "neighbour rank");
        K.nbr[k] = nbr[k];
    }
    return 0;
}

// Operator of the local rows (device pointers, kept by reference).  col: index into the extended vector
// [n local | n_ghost]; col16 (optional): column - row for the interior slices; order (optional): slice
// permutation with the n_bnd boundary slices last.
extern "C" int tc_dist_set_operator(tc_dist* d, int32_t nslices, const int64_t* slice_ptr, const int32_t* col,
                                    const int16_t* col16, const double* val, const double* mass,
                                    const double* adiag, const double* b, int32_t n_bnd, const int32_t* order) {
    TC_CHECK(n_bnd >= 0 && n_bnd <= nslices && (n_bnd == 0 || order), "boundary slices need the slice order");
    TcDistK& K = d->K;
    K.A = TcSell{(int)d->n, nslices, (const long long*)slice_ptr, col, val, (const short*)col16};
    K.n = (int)d->n; K.n_ghost = (int)d->n_ghost; K.rank = d->rank; K.world = d->world;
    K.n_bnd = n_bnd; K.order = order;
    K.b = b; K.mass = mass; K.adiag = adiag;
    const long long ext = d->n + d->n_ghost;
    K.y_ext = (double*)d->comm;
    K.p_ext[0] = (double*)(d->comm + off_p(ext, 0)); K.p_ext[1] = (double*)(d->comm + off_p(ext, 1));
    K.x = d->x; K.r = d->r; K.q = d->q; K.dinv = d->dinv;
    for (int w = 0; w < TC_MAX_WORLD; ++w) {
        K.peer_y[w] = nullptr; K.peer_p[0][w] = nullptr; K.peer_p[1][w] = nullptr;
        K.peer_hflag[w] = nullptr; K.peer_mb[w] = nullptr; K.peer_mbf[w] = nullptr;
    }
    for (int w = 0; w < d->world; ++w) {
        if (!d->peer[w]) continue;
        const long long e = d->peer_ext[w];
        K.peer_y[w] = (double*)d->peer[w];
        K.peer_p[0][w] = (double*)(d->peer[w] + off_p(e, 0)); K.peer_p[1][w] = (double*)(d->peer[w] + off_p(e, 1));
        K.peer_mb[w] = (double*)(d->peer[w] + off_mb(e)); K.peer_mbf[w] = (int*)(d->peer[w] + off_mbf(e));
        K.peer_hflag[w] = (int*)(d->peer[w] + off_hf(e));
    }
    K.my_mb = (double*)(d->comm + off_mb(ext)); K.my_mbf = (int*)(d->comm + off_mbf(ext));
    K.my_hflag = (int*)(d->comm + off_hf(ext));
    K.state = d->state;
    K.prof = d->prof;
    int dev = 0, sms = 148, occ = 0, khz = 1965000;
    TC_CUDA(cudaGetDevice(&dev));
    TC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TC_CUDA(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    const char* to = getenv("TC_DIST_TIMEOUT_S");
    double secs = to ? atof(to) : 20.0;
    if (!(secs > 0.0)) secs = 20.0;
    K.spin_limit = (long long)(secs * 1e3 * (double)khz);
    TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)dist_fn(), dist_threads(), 0));
    if (occ > 2048 / dist_threads()) occ = 2048 / dist_threads();
    if (occ < 1) occ = 1;
    d->grid = occ * sms;
    if (d->grid > TC_MAX_PARTIALS) d->grid = TC_MAX_PARTIALS;
    const int need = (nslices + dist_threads() / 32 - 1) / (dist_threads() / 32);
    if (need < d->grid) d->grid = need > 0 ? need : 1;
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

extern "C" int tc_dist_set_state(tc_dist* d, const double* y_local_dev) {
    TC_CUDA(cudaMemcpyAsync(d->K.y_ext, y_local_dev, sizeof(double) * d->n, cudaMemcpyDeviceToDevice, d->stream));
    TC_CUDA(cudaStreamSynchronize(d->stream));
    return 0;
}
extern "C" int tc_dist_get_state(tc_dist* d, double* y_local_dev) {
    TC_CUDA(cudaMemcpyAsync(y_local_dev, d->K.y_ext, sizeof(double) * d->n, cudaMemcpyDeviceToDevice, d->stream));
    return tc_dist_check(d);
}
// enqueue-only copies for pipelined host<->device streaming: dir 0 = set (caller -> state), 1 = get
extern "C" int tc_dist_copy_state_async(tc_dist* d, double* y_local_dev, int32_t dir) {
    if (dir == 0)
        TC_CUDA(cudaMemcpyAsync(d->K.y_ext, y_local_dev, sizeof(double) * d->n, cudaMemcpyDeviceToDevice, d->stream));
    else
        TC_CUDA(cudaMemcpyAsync(y_local_dev, d->K.y_ext, sizeof(double) * d->n, cudaMemcpyDeviceToDevice, d->stream));
    return 0;
}
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
        TC_CUDA(cudaLaunchCooperativeKernel(dist_fn(), dim3(d->grid), dim3(dist_threads()), args, 0, d->stream));
    }
    return 0;
}

// stats: iters_last, iters_total, comm error flag, solves (+ grid * 1e6), failed solves
extern "C" int tc_dist_stats(tc_dist* d, int64_t* stats5) {
    int st[16];
    if (dist_read_state(d, st)) return -1;
    stats5[0] = st[2]; stats5[1] = st[3]; stats5[2] = st[4]; stats5[3] = st[5] + 1000000LL * d->grid; stats5[4] = st[8];
    return 0;
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
        if (w != d->rank && d->peer[w]) { cudaIpcCloseMemHandle(d->peer[w]); d->peer[w] = nullptr; }
    d->peers_open = false;
    return 0;
}
extern "C" int tc_dist_destroy(tc_dist* d) {
    if (!d) return 0;
    tc_dist_close_peers(d);
    cudaFree(d->comm); cudaFree(d->x); cudaFree(d->r); cudaFree(d->q); cudaFree(d->dinv);
    cudaFree(d->state); cudaFree(d->prof);
    cudaFree(d->K.bar.ctr); cudaFree(d->K.bar.part); cudaFree(d->K.bar.res);
    cudaFree(d->send_row); cudaFree(d->send_peer); cudaFree(d->send_dst);
    delete d;
    return 0;
}
