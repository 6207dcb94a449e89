// Row-partitioned implicit step for ONE large FE body across GPUs (BASELINE config #5):
// fused compute + collective.  Each rank owns a contiguous slab of operator rows and runs
// ONE cooperative persistent kernel per theta step.  Inside that kernel
//   * the halo layers of the gathered vector are written straight into the neighbours'
//     HBM with peer stores over NVLink (no NCCL call, no host round trip),
//   * the 2-3 scalar reductions of each PCG iteration go through per-rank mailboxes that
//     every rank writes with peer stores and sums in rank order (bit-identical scalars on
//     all ranks -> identical convergence decisions, deterministic),
// so communication overlaps the slab's own SpMV / vector phases block by block.
// Peer memory is exchanged once at setup as CUDA IPC handles (plumbing via
// torch.distributed in thermca_b200/dist.py).  The reference has nothing comparable (no
// multi-process code); per rank the arithmetic is the single-GPU PCG of
// pcg.cuh on rows [row_begin, row_end).
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/thermca_b200.h"
#include "parts.cuh"
namespace cg = cooperative_groups;

#define TC_MAX_WORLD 16
#define TC_SPIN_LIMIT (6000000000LL)     // ~3 s at 2 GHz: never hang the GPU on a lost peer

struct TcDistK {
    TcSell A;
    long long row_begin;
    int n, H, HP, rank, world;            // HP: halo slot rounded up to 16 doubles (128-byte aligned local part)
    int n_lo, n_hi;                       // local row counts of the neighbours
    const double *b, *mass, *adiag;
    double *y_ext, *p_ext;                // [H | n | H]
    double *x, *r, *q, *dinv;
    double *pdir, *sdir;                  // single-reduction CG only: search direction p and s = P p
    double* partial;                      // 4 banks
    double* peer_y[2]; double* peer_p[2]; // neighbours' extended vectors (peer memory)
    int* peer_hflag[2];                   // neighbours' halo flags
    int* my_hflag;                        // [0] from lower, [1] from upper neighbour
    double* peer_mb[TC_MAX_WORLD]; int* peer_mbf[TC_MAX_WORLD];
    double* my_mb; int* my_mbf;           // mb[2][3][MAXW], mbf[2][MAXW]
    int* state;                           // [0] reduction seq, [1] halo seq, [2] iters last, [3] iters total, [4] error, [5] solves, [6] halo post counter
    double theta, h, rtol;
    int maxit;
    unsigned long long* prof;             // [2][16] phase times in ns: block 0 and the last block
    double* lres;                         // [2][4] locally published reduction results
};

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void spin_until(const int* flag, int seq, int* err) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < seq) {
        if (*((volatile int*)err)) return;
        if (clock64() - t0 > TC_SPIN_LIMIT) { *((volatile int*)err) = 1; return; }
    }
}

__device__ __forceinline__ double bank_sum(const double* bank, int G, double* sh) {
    __shared__ double bc;
    const double s = ordered_partial_sum(bank, G, sh);
    __syncthreads();
    if (threadIdx.x == 0) bc = s;
    __syncthreads();
    return bc;
}

#define TC_HALO_BLOCKS 8      // blocks that push the halo layers (112 KB per side on the 10M slab)

// Post my boundary rows into the neighbours' halo regions.  Only the first TC_HALO_BLOCKS blocks
// take part; the last of them to finish raises the neighbours' flags.  No grid-wide barrier.
__device__ void halo_post(const TcDistK& S, const double* vec, double* const* peer_vec, int hseq) {
    double* lo = peer_vec[0];
    double* hi = peer_vec[1];
    if (!lo && !hi) return;
    const int nb = min((int)gridDim.x, TC_HALO_BLOCKS);
    if ((int)blockIdx.x >= nb) return;
    const int H = S.H, n = S.n;
    const int HP = S.HP;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H; i += nb * blockDim.x) {
        if (lo) lo[(long long)HP + S.n_lo + i] = vec[HP + i];               // -> their upper halo
        if (hi) hi[HP - H + i] = vec[HP + n - H + i];                       // -> their lower halo
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(S.state + 6, 1);
        if (done == nb - 1) {
            S.state[6] = 0;
            __threadfence_system();
            if (lo) st_release_sys(S.peer_hflag[0] + 1, hseq);
            if (hi) st_release_sys(S.peer_hflag[1] + 0, hseq);
        }
    }
}
// Wait until both neighbours have posted halo `hseq` (called before the boundary slices only).
// Only blocks that own boundary slices wait; they read the gathered vector through L2
// (ld.global.cg), so no L1 invalidation is needed and the interior pass of co-resident blocks
// keeps its cached lines.
__device__ void halo_wait(const TcDistK& S, int hseq, int nb_max) {
    if (!S.peer_y[0] && !S.peer_y[1]) return;
    if ((int)(blockIdx.x * (blockDim.x >> 5)) >= nb_max) return;      // this block has no boundary slice
    if (threadIdx.x == 0) {
        if (S.peer_y[0]) spin_until(S.my_hflag + 0, hseq, S.state + 4);
        if (S.peer_y[1]) spin_until(S.my_hflag + 1, hseq, S.state + 4);
    }
    __syncthreads();
}

// every block holds the same local values; result = sum over ranks in rank order.
// Only block 0 talks to the peers (8 remote stores + 8 pollers on the NVLink-written mailbox); it then
// publishes the sum in a local slot that the other ~1200 blocks poll, so the mailbox lines are not
// hammered by ten thousand system-scope loads while the peers' stores are in flight.
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ void allreduce3(const TcDistK& S, double* v, int seq) {
    const int par = seq & 1;
    if (S.world == 1) return;
    volatile double* res = S.lres + par * 4;      // local result slot [2][4]
    int* lflag = S.state + 8 + par;               // local flags
    if (blockIdx.x == 0) {
        if (threadIdx.x < S.world) {
            const int peer = threadIdx.x;
            volatile double* mb = S.peer_mb[peer];
            for (int k = 0; k < 3; ++k) mb[(par * 3 + k) * TC_MAX_WORLD + S.rank] = v[k];
            st_release_sys(S.peer_mbf[peer] + par * TC_MAX_WORLD + S.rank, seq);   // orders the stores above
            spin_until(S.my_mbf + par * TC_MAX_WORLD + threadIdx.x, seq, S.state + 4);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            volatile double* mine = S.my_mb;
            double out[3] = {0.0, 0.0, 0.0};
            for (int w = 0; w < S.world; ++w)
                for (int k = 0; k < 3; ++k) out[k] += mine[(par * 3 + k) * TC_MAX_WORLD + w];
            res[0] = out[0]; res[1] = out[1]; res[2] = out[2];
            st_release_gpu(lflag, seq);
        }
    }
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while (ld_acquire_gpu(lflag) < seq) {
            if (*((volatile int*)(S.state + 4))) break;
            if (clock64() - t0 > TC_SPIN_LIMIT) { *((volatile int*)(S.state + 4)) = 1; break; }
        }
    }
    __syncthreads();
    v[0] = res[0]; v[1] = res[1]; v[2] = res[2];
    __syncthreads();
}

#define TC_DIST_THREADS 256

// Slab SpMV with the boundary slices LAST in every warp's grid-stride sequence: one balanced loop
// over the permuted slice order [interior | lower boundary | upper boundary]; a warp waits for the
// neighbours' halo flags only when it reaches its first boundary slice (by then the interior part,
// ~99.7 % of the work, has hidden the NVLink latency), and reads the gathered vector of boundary
// slices through L2 (ld.global.cg).
template <int MODE>
__device__ __forceinline__ double slab_spmv(const TcDistK& S, const double* xg, double* y, const double* b,
                                            double shift, int nb_lo, int nb_hi, int hseq) {
    const TcSell& A = S.A;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int warp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_block;
    const int nsl = A.nslices, n_int = nsl - nb_lo - nb_hi;
    bool waited = false;
    double dot = 0.0;
    for (int idx = warp; idx < nsl; idx += nwarps) {
        int s;
        const bool boundary = idx >= n_int;
        if (!boundary) s = nb_lo + idx;
        else { const int k = idx - n_int; s = k < nb_lo ? k : nsl - nb_hi + (k - nb_lo); }
        if (boundary && !waited) {
            if (lane == 0) {
                if (S.peer_y[0]) spin_until(S.my_hflag + 0, hseq, S.state + 4);
                if (S.peer_y[1]) spin_until(S.my_hflag + 1, hseq, S.state + 4);
            }
            __syncwarp();
            waited = true;
        }
        const long long base = A.slice_ptr[s];
        const int width = (int)((A.slice_ptr[s + 1] - base) >> 5);
        const double* v = A.val + base + lane;
        double sum = 0.0;
        if (A.col16) {
            const short* c = A.col16 + base + lane;
            const double* xr0 = xg + (S.row_begin + s * 32 + lane);
            if (!boundary) {
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
                for (int k = 0; k < width; ++k)
                    sum = __dadd_rn(sum, __dmul_rn(ld_stream_f64(v + 32 * k), tc_ld_x<2>(xr0 + ld_stream_s16(c + 32 * k))));
            }
        } else {
        const int* c = A.col + base + lane;
        if (!boundary) {
            int k = 0;
            for (; k + 4 <= width; k += 4) {
                const double v0 = ld_stream_f64(v + 32 * (k + 0)), v1 = ld_stream_f64(v + 32 * (k + 1));
                const double v2 = ld_stream_f64(v + 32 * (k + 2)), v3 = ld_stream_f64(v + 32 * (k + 3));
                const int c0 = ld_stream_s32(c + 32 * (k + 0)), c1 = ld_stream_s32(c + 32 * (k + 1));
                const int c2 = ld_stream_s32(c + 32 * (k + 2)), c3 = ld_stream_s32(c + 32 * (k + 3));
                const double x0 = tc_ld_x<0>(xg + c0), x1 = tc_ld_x<0>(xg + c1), x2 = tc_ld_x<0>(xg + c2), x3 = tc_ld_x<0>(xg + c3);
                sum = __dadd_rn(sum, __dmul_rn(v0, x0));
                sum = __dadd_rn(sum, __dmul_rn(v1, x1));
                sum = __dadd_rn(sum, __dmul_rn(v2, x2));
                sum = __dadd_rn(sum, __dmul_rn(v3, x3));
            }
            for (; k < width; ++k)
                sum = __dadd_rn(sum, __dmul_rn(ld_stream_f64(v + 32 * k), tc_ld_x<0>(xg + ld_stream_s32(c + 32 * k))));
        } else {
            for (int k = 0; k < width; ++k)
                sum = __dadd_rn(sum, __dmul_rn(ld_stream_f64(v + 32 * k), tc_ld_x<2>(xg + ld_stream_s32(c + 32 * k))));
        }
        }
        const int row = s * 32 + lane;
        if (row < A.n) {
            if (MODE == SPMV_B) {
                y[row] = __dsub_rn(b[row], sum);
            } else {
                const double xr = tc_ld_x<0>(xg + S.row_begin + row);
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

template <int MB>
__global__ void __launch_bounds__(TC_DIST_THREADS, MB)
tc_dist_theta_kernel(TcDistK S) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[32];
    const int G = gridDim.x, n = S.n, H = S.H;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    int rseq = S.state[0], hseq = S.state[1];
    PROF_INIT();
    const double shift = S.theta * S.h;
    // vectors are gathered by GLOBAL column: element of global row g sits at ext[H + g - row_begin]
    const int HP = S.HP;
    const double* y_g = S.y_ext + HP - S.row_begin;
    const double* p_g = S.p_ext + HP - S.row_begin;
    double* y_loc = S.y_ext + HP;
    double* p_loc = S.p_ext + HP;
    double* const p_lo = S.peer_p[0];      // neighbours' p_ext (peer memory) or null
    double* const p_hi = S.peer_p[1];

    // ---- right-hand side f = b - A y (thermca/solver.py:281) on the slab ---------------------
    // slices whose rows reach into a neighbour's rows: done last, after the halo has landed
    const int nsl = S.A.nslices;
    const int nb_lo = S.peer_y[0] ? min(nsl, (H + 31) / 32) : 0;
    const int nb_hi = S.peer_y[1] ? min(nsl - nb_lo, (H + 31) / 32 + 1) : 0;
    halo_post(S, S.y_ext, S.peer_y, ++hseq);
    slab_spmv<SPMV_B>(S, y_g, S.q, S.b, 0.0, nb_lo, nb_hi, hseq);
    grid.sync();
    double rho_l = 0.0, bb_l = 0.0;
    for (int i = gtid; i < n; i += gsz) {
        const double m = S.mass[i];
        const double bsc = m * S.q[i];
        const double d = 1.0 / (m * (1.0 + shift * S.adiag[i]));
        const double z = d * bsc;
        S.r[i] = bsc; S.x[i] = 0.0; S.dinv[i] = d; p_loc[i] = z;
        rho_l += bsc * z; bb_l += bsc * bsc;
        // boundary rows go straight into the neighbours' halo slots (peer stores over NVLink)
        bool remote = false;
        if (p_lo && i < H) { p_lo[(long long)HP + S.n_lo + i] = z; remote = true; }
        if (p_hi && i >= n - H) { p_hi[HP - H + (i - (n - H))] = z; remote = true; }
        if (remote) __threadfence_system();
    }
    {
        const double r0 = block_sum(rho_l, sh), r1 = block_sum(bb_l, sh);
        if (threadIdx.x == 0) { S.partial[blockIdx.x] = r0; S.partial[TC_MAX_PARTIALS + blockIdx.x] = r1; }
    }
    grid.sync();
    double red[3];
    red[0] = bank_sum(S.partial, G, sh);
    red[1] = bank_sum(S.partial + TC_MAX_PARTIALS, G, sh);
    red[2] = 0.0;
    allreduce3(S, red, ++rseq);
    double rho = red[0];
    const double tol2 = S.rtol * S.rtol * red[1];
    int it = 0;
    if (red[1] > 0.0) {
        for (; it < S.maxit;) {
            PROF(0);
            ++hseq;
            if (blockIdx.x == 0 && threadIdx.x == 0) {      // halo stores were fenced before the last grid.sync
                if (p_lo) st_release_sys(S.peer_hflag[0] + 1, hseq);
                if (p_hi) st_release_sys(S.peer_hflag[1] + 0, hseq);
            }
            PROF(1);
            const double dot = slab_spmv<SPMV_SHIFT>(S, p_g, S.q, nullptr, shift, nb_lo, nb_hi, hseq);
            PROF(2);
            const double bs = block_sum(dot, sh);
            if (threadIdx.x == 0) S.partial[blockIdx.x] = bs;
            PROF(4);
            grid.sync();
            PROF(5);
            red[0] = bank_sum(S.partial, G, sh); red[1] = 0.0; red[2] = 0.0;
            allreduce3(S, red, ++rseq);
            PROF(6);
            const double alpha = rho / red[0];
            double rn = 0.0, rr = 0.0;
            for (int i = gtid; i < n; i += gsz) {
                S.x[i] += alpha * p_loc[i];
                const double ri = S.r[i] - alpha * S.q[i];
                S.r[i] = ri;
                rn += ri * (S.dinv[i] * ri);
                rr += ri * ri;
            }
            const double r0 = block_sum(rn, sh), r1 = block_sum(rr, sh);
            if (threadIdx.x == 0) {
                S.partial[2 * TC_MAX_PARTIALS + blockIdx.x] = r0;
                S.partial[3 * TC_MAX_PARTIALS + blockIdx.x] = r1;
            }
            PROF(7);
            grid.sync();
            PROF(8);
            red[0] = bank_sum(S.partial + 2 * TC_MAX_PARTIALS, G, sh);
            red[1] = bank_sum(S.partial + 3 * TC_MAX_PARTIALS, G, sh);
            red[2] = 0.0;
            allreduce3(S, red, ++rseq);
            PROF(9);
            ++it;
            if (!(red[1] > tol2) || *((volatile int*)(S.state + 4))) break;
            const double beta = red[0] / rho;
            for (int i = gtid; i < n; i += gsz) {
                const double pn = S.dinv[i] * S.r[i] + beta * p_loc[i];
                p_loc[i] = pn;
                bool remote = false;
                if (p_lo && i < H) { p_lo[(long long)HP + S.n_lo + i] = pn; remote = true; }
                if (p_hi && i >= n - H) { p_hi[HP - H + (i - (n - H))] = pn; remote = true; }
                if (remote) __threadfence_system();
            }
            rho = red[0];
            PROF(10);
            grid.sync();
            PROF(11);
        }
    }
    // ---- y += h * z --------------------------------------------------------------------------
    for (int i = gtid; i < n; i += gsz) y_loc[i] += S.h * S.x[i];
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S.state[0] = rseq; S.state[1] = hseq; S.state[2] = it; S.state[3] += it; S.state[5] += 1;
    }
}


// ---- single-reduction variant (Chronopoulos & Gear 1989) ----------------------------------------
// Classic PCG needs two global reductions per iteration (p.Pp, then r.z); across GPUs each one is a
// NVLink round trip plus the skew between ranks.  With s = P p kept by recurrence,
//   u = D^-1 r,  w = P u,  gamma = r.u,  delta = w.u           (ONE reduction: gamma, delta, r.r)
//   beta = gamma/gamma_old,  alpha = gamma / (delta - beta*gamma/alpha_old)
//   p = u + beta p,  s = w + beta s,  x += alpha p,  r -= alpha s
// one allreduce and two grid barriers per iteration remain (classic: two and three), for about the same
// HBM traffic.  The gathered vector (with halo) is u; it is pushed to the neighbours from the vector phase.
template <int MB>
__global__ void __launch_bounds__(TC_DIST_THREADS, MB)
tc_dist_theta_cgcg_kernel(TcDistK S) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh[32];
    const int G = gridDim.x, n = S.n, H = S.H;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    int rseq = S.state[0], hseq = S.state[1];
    PROF_INIT();
    const double shift = S.theta * S.h;
    const int HP = S.HP;
    const double* y_g = S.y_ext + HP - S.row_begin;
    const double* u_g = S.p_ext + HP - S.row_begin;
    double* y_loc = S.y_ext + HP;
    double* u_loc = S.p_ext + HP;
    double* const u_lo = S.peer_p[0];
    double* const u_hi = S.peer_p[1];
    double* w = S.q;
    const int nsl = S.A.nslices;
    const int nb_lo = S.peer_y[0] ? min(nsl, (H + 31) / 32) : 0;
    const int nb_hi = S.peer_y[1] ? min(nsl - nb_lo, (H + 31) / 32 + 1) : 0;

    halo_post(S, S.y_ext, S.peer_y, ++hseq);
    slab_spmv<SPMV_B>(S, y_g, S.q, S.b, 0.0, nb_lo, nb_hi, hseq);
    grid.sync();
    double g_l = 0.0, bb_l = 0.0;
    for (int i = gtid; i < n; i += gsz) {
        const double m = S.mass[i];
        const double bsc = m * S.q[i];
        const double d = 1.0 / (m * (1.0 + shift * S.adiag[i]));
        const double u = d * bsc;
        S.r[i] = bsc; S.x[i] = 0.0; S.dinv[i] = d; u_loc[i] = u; S.pdir[i] = 0.0; S.sdir[i] = 0.0;
        g_l += bsc * u; bb_l += bsc * bsc;
        bool remote = false;
        if (u_lo && i < H) { u_lo[(long long)HP + S.n_lo + i] = u; remote = true; }
        if (u_hi && i >= n - H) { u_hi[HP - H + (i - (n - H))] = u; remote = true; }
        if (remote) __threadfence_system();
    }
    {
        const double r0 = block_sum(g_l, sh), r1 = block_sum(bb_l, sh);
        if (threadIdx.x == 0) { S.partial[blockIdx.x] = r0; S.partial[3 * TC_MAX_PARTIALS + blockIdx.x] = r1; }
    }
    grid.sync();
    // partial banks: gamma -> bank rp (0/1), r.r -> bank 3+rp (3/4), delta -> bank 2.  gamma / r.r
    // alternate by iteration parity: a fast block may already publish iteration i+1 while a slow one
    // is still summing iteration i.
    int rp = 0;
    double red[3];
    double gamma = 0.0, alpha = 0.0, beta = 0.0, tol2 = 0.0;
    int it = 0;
    bool first = true;
    for (;;) {
        PROF(0);
        // ---- w = P u (halo of u was fenced before the barrier above) -------------------------
        ++hseq;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (u_lo) st_release_sys(S.peer_hflag[0] + 1, hseq);
            if (u_hi) st_release_sys(S.peer_hflag[1] + 0, hseq);
        }
        const double dot = slab_spmv<SPMV_SHIFT>(S, u_g, w, nullptr, shift, nb_lo, nb_hi, hseq);
        PROF(2);
        const double bs = block_sum(dot, sh);
        if (threadIdx.x == 0) S.partial[2 * TC_MAX_PARTIALS + blockIdx.x] = bs;
        PROF(4);
        grid.sync();
        PROF(5);
        red[0] = bank_sum(S.partial + rp * TC_MAX_PARTIALS, G, sh);          // gamma = r.u
        red[1] = bank_sum(S.partial + 2 * TC_MAX_PARTIALS, G, sh);           // delta = w.u
        red[2] = bank_sum(S.partial + (3 + rp) * TC_MAX_PARTIALS, G, sh);    // r.r
        allreduce3(S, red, ++rseq);
        PROF(6);
        if (first) {
            tol2 = S.rtol * S.rtol * red[2];
            if (!(red[2] > 0.0)) break;
            beta = 0.0; alpha = red[0] / red[1];
            first = false;
        } else {
            if (!(red[2] > tol2) || it >= S.maxit || *((volatile int*)(S.state + 4))) break;
            beta = red[0] / gamma;
            alpha = red[0] / (red[1] - beta * red[0] / alpha);
        }
        gamma = red[0];
        // ---- vector phase -----------------------------------------------------------------------
        double gn = 0.0, rr = 0.0;
        for (int i = gtid; i < n; i += gsz) {
            const double pn = u_loc[i] + beta * S.pdir[i];
            const double sn = w[i] + beta * S.sdir[i];
            S.pdir[i] = pn; S.sdir[i] = sn;
            S.x[i] += alpha * pn;
            const double ri = S.r[i] - alpha * sn;
            S.r[i] = ri;
            const double un = S.dinv[i] * ri;
            u_loc[i] = un;
            gn += ri * un; rr += ri * ri;
            bool remote = false;
            if (u_lo && i < H) { u_lo[(long long)HP + S.n_lo + i] = un; remote = true; }
            if (u_hi && i >= n - H) { u_hi[HP - H + (i - (n - H))] = un; remote = true; }
            if (remote) __threadfence_system();
        }
        const double r0 = block_sum(gn, sh), r1 = block_sum(rr, sh);
        rp ^= 1;
        if (threadIdx.x == 0) {
            S.partial[rp * TC_MAX_PARTIALS + blockIdx.x] = r0;
            S.partial[(3 + rp) * TC_MAX_PARTIALS + blockIdx.x] = r1;
        }
        ++it;
        PROF(7);
        grid.sync();
        PROF(8);
    }
    for (int i = gtid; i < n; i += gsz) y_loc[i] += S.h * S.x[i];
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S.state[0] = rseq; S.state[1] = hseq; S.state[2] = it; S.state[3] += it; S.state[5] += 1;
    }
}

// ---- host ------------------------------------------------------------------------------------
static void* dist_fn() {
    static int v = -1;
    static int classic = -1;
    if (v < 0) {
        const char* e = getenv("TC_PCG_OCC");
        v = e ? atoi(e) : 8;
        // measured on 2 and 8 B200 (profiles/README.md): the single-reduction variant saves ~40 % of the
        // reduction wait but its fused vector phase streams 12 arrays and runs ~25 % slower, net -3..-8 %;
        // two-reduction PCG stays the default, TC_DIST_CG=cgcg selects the other one
        const char* a = getenv("TC_DIST_CG");
        classic = (a && a[0] == 'c' && a[1] == 'g') ? 0 : 1;
    }
    if (!classic) {
        switch (v) {
            case 5: return (void*)tc_dist_theta_cgcg_kernel<5>;
            case 6: return (void*)tc_dist_theta_cgcg_kernel<6>;
            default: return (void*)tc_dist_theta_cgcg_kernel<8>;
        }
    }
    switch (v) {
        case 5: return (void*)tc_dist_theta_kernel<5>;
        case 6: return (void*)tc_dist_theta_kernel<6>;
        default: return (void*)tc_dist_theta_kernel<8>;
    }
}

struct tc_dist {
    cudaStream_t stream;
    int rank, world;
    long long n, H, HP, row_begin;
    char* comm = nullptr;             // my exported buffer
    size_t comm_bytes = 0;
    size_t off_y, off_p, off_mb, off_mbf, off_hflag;
    char* peer[TC_MAX_WORLD];
    long long peer_n[TC_MAX_WORLD];
    double *x = nullptr, *r = nullptr, *q = nullptr, *dinv = nullptr, *partial = nullptr;
    double *pdir = nullptr, *sdir = nullptr;
    int* state = nullptr;
    unsigned long long* prof = nullptr;
    double* lres = nullptr;
    TcDistK K;
    int grid = 0;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int tc_dist_create(tc_dist** out, void* stream, int32_t rank, int32_t world, int64_t n_loc,
                              int64_t halo, int64_t row_begin) {
    TC_CHECK(world >= 1 && world <= TC_MAX_WORLD, "world size");
    TC_CHECK(halo <= n_loc, "halo wider than the slab");
    tc_dist* d = new tc_dist();
    memset(d->peer, 0, sizeof d->peer);
    d->stream = (cudaStream_t)stream; d->rank = rank; d->world = world;
    d->n = n_loc; d->H = halo; d->row_begin = row_begin;
    d->HP = (halo + 15) / 16 * 16;
    const size_t ext = (size_t)(n_loc + 2 * d->HP) * sizeof(double);
    d->off_y = 0;
    d->off_p = align_up(ext, 256);
    d->off_mb = d->off_p + align_up(ext, 256);
    d->off_mbf = d->off_mb + align_up(sizeof(double) * 2 * 3 * TC_MAX_WORLD, 256);
    d->off_hflag = d->off_mbf + align_up(sizeof(int) * 2 * TC_MAX_WORLD, 256);
    d->comm_bytes = d->off_hflag + 256;
    TC_CUDA(cudaMalloc(&d->comm, d->comm_bytes));
    TC_CUDA(cudaMemset(d->comm, 0, d->comm_bytes));
    const size_t nb = sizeof(double) * (size_t)n_loc;
    TC_CUDA(cudaMalloc(&d->x, nb)); TC_CUDA(cudaMalloc(&d->r, nb));
    TC_CUDA(cudaMalloc(&d->q, nb)); TC_CUDA(cudaMalloc(&d->dinv, nb));
    TC_CUDA(cudaMalloc(&d->pdir, nb)); TC_CUDA(cudaMalloc(&d->sdir, nb));
    TC_CUDA(cudaMalloc(&d->partial, sizeof(double) * 6 * TC_MAX_PARTIALS));
    TC_CUDA(cudaMalloc(&d->state, sizeof(int) * 16));
    TC_CUDA(cudaMemset(d->state, 0, sizeof(int) * 16));
    TC_CUDA(cudaMalloc(&d->lres, sizeof(double) * 8));
    TC_CUDA(cudaMemset(d->lres, 0, sizeof(double) * 8));
    TC_CUDA(cudaMalloc(&d->prof, sizeof(unsigned long long) * 32));
    TC_CUDA(cudaMemset(d->prof, 0, sizeof(unsigned long long) * 32));
    d->peer[rank] = d->comm;
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

// handles: world * 64 bytes; peer_n: local row counts of all ranks
extern "C" int tc_dist_open_peers(tc_dist* d, const void* handles, const int64_t* peer_n) {
    for (int w = 0; w < d->world; ++w) {
        d->peer_n[w] = peer_n[w];
        if (w == d->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * w, 64);
        void* p = nullptr;
        TC_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        d->peer[w] = (char*)p;
    }
    return 0;
}

extern "C" int tc_dist_set_operator(tc_dist* d, int32_t nslices, const int64_t* slice_ptr, const int32_t* col,
                                    const int16_t* col16, const double* val, const double* mass,
                                    const double* adiag, const double* b) {
    TcDistK& K = d->K;
    K.A = TcSell{(int)d->n, nslices, (const long long*)slice_ptr, col, val, (const short*)col16};
    K.row_begin = d->row_begin; K.n = (int)d->n; K.H = (int)d->H; K.HP = (int)d->HP; K.rank = d->rank; K.world = d->world;
    K.b = b; K.mass = mass; K.adiag = adiag;
    K.y_ext = (double*)(d->comm + d->off_y); K.p_ext = (double*)(d->comm + d->off_p);
    if (getenv("TC_DIST_EXP") && d->world == 1) {      // experiment: vectors outside the IPC-exported allocation
        double* alt = nullptr;
        TC_CUDA(cudaMalloc(&alt, 2 * align_up((size_t)(d->n + 2 * d->HP) * sizeof(double), 256)));
        TC_CUDA(cudaMemset(alt, 0, 2 * align_up((size_t)(d->n + 2 * d->HP) * sizeof(double), 256)));
        K.y_ext = alt; K.p_ext = (double*)((char*)alt + align_up((size_t)(d->n + 2 * d->HP) * sizeof(double), 256));
    }
    K.x = d->x; K.r = d->r; K.q = d->q; K.dinv = d->dinv; K.partial = d->partial;
    K.pdir = d->pdir; K.sdir = d->sdir;
    const int lo = d->rank - 1, hi = d->rank + 1;
    // peers share my buffer layout except for their own n: offsets depend on n_peer
    auto ext_bytes = [&](int w) { return align_up((size_t)(d->peer_n[w] + 2 * d->HP) * sizeof(double), 256); };
    auto off_p = [&](int w) { return ext_bytes(w); };
    auto off_mb = [&](int w) { return 2 * ext_bytes(w); };
    auto off_mbf = [&](int w) { return off_mb(w) + align_up(sizeof(double) * 2 * 3 * TC_MAX_WORLD, 256); };
    auto off_hf = [&](int w) { return off_mbf(w) + align_up(sizeof(int) * 2 * TC_MAX_WORLD, 256); };
    for (int s = 0; s < 2; ++s) { K.peer_y[s] = nullptr; K.peer_p[s] = nullptr; K.peer_hflag[s] = nullptr; }
    K.n_lo = K.n_hi = 0;
    if (lo >= 0) {
        K.peer_y[0] = (double*)(d->peer[lo]); K.peer_p[0] = (double*)(d->peer[lo] + off_p(lo));
        K.peer_hflag[0] = (int*)(d->peer[lo] + off_hf(lo)); K.n_lo = (int)d->peer_n[lo];
    }
    if (hi < d->world) {
        K.peer_y[1] = (double*)(d->peer[hi]); K.peer_p[1] = (double*)(d->peer[hi] + off_p(hi));
        K.peer_hflag[1] = (int*)(d->peer[hi] + off_hf(hi)); K.n_hi = (int)d->peer_n[hi];
    }
    for (int w = 0; w < TC_MAX_WORLD; ++w) { K.peer_mb[w] = nullptr; K.peer_mbf[w] = nullptr; }
    for (int w = 0; w < d->world; ++w) {
        K.peer_mb[w] = (double*)(d->peer[w] + off_mb(w));
        K.peer_mbf[w] = (int*)(d->peer[w] + off_mbf(w));
    }
    K.my_mb = (double*)(d->comm + d->off_mb); K.my_mbf = (int*)(d->comm + d->off_mbf);
    K.my_hflag = (int*)(d->comm + d->off_hflag);
    K.state = d->state;
    K.prof = d->prof;
    K.lres = d->lres;
    int dev = 0, sms = 148, occ = 0;
    TC_CUDA(cudaGetDevice(&dev));
    TC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)dist_fn(), TC_DIST_THREADS, 0));
    if (occ > 8) occ = 8;
    d->grid = occ * sms;
    if (d->grid > TC_MAX_PARTIALS) d->grid = TC_MAX_PARTIALS;
    const int need = (nslices + 7) / 8;
    if (need < d->grid) d->grid = need > 0 ? need : 1;
    return 0;
}

extern "C" int tc_dist_set_state(tc_dist* d, const double* y_local_dev) {
    TC_CUDA(cudaMemcpyAsync(d->K.y_ext + d->HP, y_local_dev, sizeof(double) * d->n, cudaMemcpyDeviceToDevice, d->stream));
    TC_CUDA(cudaStreamSynchronize(d->stream));
    return 0;
}
extern "C" int tc_dist_get_state(tc_dist* d, double* y_local_dev) {
    TC_CUDA(cudaMemcpyAsync(y_local_dev, d->K.y_ext + d->HP, sizeof(double) * d->n, cudaMemcpyDeviceToDevice, d->stream));
    TC_CUDA(cudaStreamSynchronize(d->stream));
    return 0;
}

extern "C" int tc_dist_theta_steps(tc_dist* d, int32_t nsteps, double theta, double h, double rtol, int32_t maxit) {
    d->K.theta = theta; d->K.h = h; d->K.rtol = rtol; d->K.maxit = maxit;
    for (int k = 0; k < nsteps; ++k) {
        void* args[] = {(void*)&d->K};
        TC_CUDA(cudaLaunchCooperativeKernel(dist_fn(), dim3(d->grid), dim3(TC_DIST_THREADS), args, 0, d->stream));
    }
    return 0;
}

// stats: iters_last, iters_total, error flag, solves
extern "C" int tc_dist_stats(tc_dist* d, int64_t* stats4) {
    int st[16];
    TC_CUDA(cudaStreamSynchronize(d->stream));
    TC_CUDA(cudaMemcpy(st, d->state, sizeof st, cudaMemcpyDeviceToHost));
    stats4[0] = st[2]; stats4[1] = st[3]; stats4[2] = st[4]; stats4[3] = st[5] + 1000000LL * d->grid;
    return 0;
}

// phase profile: 32 accumulated nanosecond counters (block 0: [0..15], last block: [16..31]); cleared on read
extern "C" int tc_dist_profile(tc_dist* d, uint64_t* out32) {
    TC_CUDA(cudaStreamSynchronize(d->stream));
    TC_CUDA(cudaMemcpy(out32, d->prof, sizeof(unsigned long long) * 32, cudaMemcpyDeviceToHost));
    TC_CUDA(cudaMemset(d->prof, 0, sizeof(unsigned long long) * 32));
    return 0;
}

extern "C" int tc_dist_destroy(tc_dist* d) {
    if (!d) return 0;
    for (int w = 0; w < d->world; ++w)
        if (w != d->rank && d->peer[w]) cudaIpcCloseMemHandle(d->peer[w]);
    cudaFree(d->comm); cudaFree(d->x); cudaFree(d->r); cudaFree(d->q); cudaFree(d->dinv);
    cudaFree(d->pdir); cudaFree(d->sdir);
    cudaFree(d->partial); cudaFree(d->state); cudaFree(d->prof); cudaFree(d->lres);
    delete d;
    return 0;
}
