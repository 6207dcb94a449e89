// Host side of the C ABI (include/thermca_b200.h): solver object, right-hand side
// pipeline, scipy-equivalent adaptive Runge-Kutta with on-device step control, and the
// implicit (theta / ROS2) steps on top of the Jacobi-PCG.
//
// What it replaces in the reference: the body of thermca/solver.py:transient_solution
// (solver.py:61-345) including simple_solve_ivp (solver.py:348-390) and, through scipy,
// RungeKutta._step_impl / rk_step / select_initial_step (scipy/integrate/_ivp/rk.py,
// common.py — third-party, restated from the installed 1.18.1 sources the reference pins
// as 1.14.1).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/thermca_b200.h"
#include "network.cuh"
#include "pcg.cuh"

// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
extern "C" void tc_set_error(const char* file, int line, const char* msg) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s:%d: %s", file, line, msg);
    g_err = buf;
}
extern "C" const char* tc_last_error(void) { return g_err.c_str(); }
extern "C" int tc_version(void) { return 100; }

#define TC_PCG_DEFAULT_OCC 8
#define TC_PCG_COOP_THREADS_DEFAULT 1024      // 2 blocks of 1024 threads per SM (measured: profiles/README.md)
static int pcg_coop_grid(int sm_count);

// ---- control block layout (device) ------------------------------------------------------
enum { CT_T = 0, CT_H, CT_HABS, CT_ERR, CT_TEND, CT_RTOL, CT_ATOL, CT_MAXSTEP, CT_TNEW,
       CT_D0, CT_D1, CT_D2, CT_H0, CT_EXP, CT_INTERVAL, CT_N };
enum { FL_DONE = 0, FL_ACC, FL_REJ, FL_STATUS, FL_STEP_REJECTED, FL_ACC_THIS, FL_STORED,
       FL_ACC_LAST, FL_RESULT_STEP, FL_NEXT_COLL, FL_STORE_SLOT, FL_COUNTER, FL_NUM_SKIP,
       FL_CAPACITY, FL_N };

struct TcLin { int nk; const double* k[7]; double c[7]; };

// out = y + h * sum_j c_j k_j          (scipy rk.py: dy = dot(K[:s].T, a[:s]) * h ; y + dy)
__global__ void tc_lin_kernel(long long n, const double* y, TcLin L,
                              const double* __restrict__ ctrl, double* out,
                              const int* __restrict__ done) {
    if (done && *done) return;
    const double h = ctrl[CT_H];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 7; ++j)
            if (j < L.nk) acc += L.k[j][i] * L.c[j];
        out[i] = y[i] + acc * h;
    }
}

// partial sums of (a_i/scale_i)^2 [, (b_i/scale_i)^2], scale = atol + |y_i| rtol   (select_initial_step)
__global__ void tc_norm2_kernel(long long n, const double* __restrict__ y, const double* __restrict__ a,
                                const double* __restrict__ b, const double* __restrict__ bsub,
                                const double* __restrict__ ctrl, double* __restrict__ partial,
                                const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ double sh[32];
    const double rtol = ctrl[CT_RTOL], atol = ctrl[CT_ATOL];
    double s0 = 0.0, s1 = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double sc = atol + fabs(y[i]) * rtol;
        const double va = a[i] / sc;
        s0 += va * va;
        if (b) {
            const double vb = (bsub ? b[i] - bsub[i] : b[i]) / sc;
            s1 += vb * vb;
        }
    }
    const double r0 = block_sum(s0, sh);
    const double r1 = block_sum(s1, sh);
    if (threadIdx.x == 0) { partial[blockIdx.x] = r0; partial[TC_MAX_PARTIALS + blockIdx.x] = r1; }
}

// scipy common.py:select_initial_step, first half: h0 from d0, d1
__global__ void tc_init_h0_kernel(long long n, int G, const double* __restrict__ partial, double* ctrl) {
    __shared__ double sh[32];
    const double a = ordered_partial_sum(partial, G, sh);
    __syncthreads();
    const double b = ordered_partial_sum(partial + TC_MAX_PARTIALS, G, sh);
    if (threadIdx.x == 0) {
        const double d0 = sqrt(a) / sqrt((double)n), d1 = sqrt(b) / sqrt((double)n);
        double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        h0 = fmin(h0, ctrl[CT_INTERVAL]);
        ctrl[CT_D0] = d0; ctrl[CT_D1] = d1; ctrl[CT_H0] = h0;
        ctrl[CT_H] = h0;            // stage time / Euler probe use h0
    }
}
// second half: d2 -> h1 -> h_abs
__global__ void tc_init_h1_kernel(long long n, int G, const double* __restrict__ partial, double* ctrl,
                                  double order) {
    __shared__ double sh[32];
    const double b = ordered_partial_sum(partial + TC_MAX_PARTIALS, G, sh);
    if (threadIdx.x == 0) {
        const double h0 = ctrl[CT_H0], d1 = ctrl[CT_D1];
        const double d2 = (sqrt(b) / sqrt((double)n)) / h0;
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
        else h1 = pow(0.01 / fmax(d1, d2), 1.0 / (order + 1.0));
        ctrl[CT_D2] = d2;
        ctrl[CT_HABS] = fmin(fmin(100.0 * h0, h1), fmin(ctrl[CT_INTERVAL], ctrl[CT_MAXSTEP]));
    }
}

// scipy rk.py:_step_impl prologue (per attempt)
__global__ void tc_prepare_kernel(double* ctrl, int* fl) {
    if (fl[FL_DONE]) return;
    const double t = ctrl[CT_T];
    double h_abs = ctrl[CT_HABS];
    const double min_step = 10.0 * fabs(::nextafter(t, (double)INFINITY) - t);
    if (fl[FL_ACC_LAST]) {
        if (h_abs > ctrl[CT_MAXSTEP]) h_abs = ctrl[CT_MAXSTEP];
        else if (h_abs < min_step) h_abs = min_step;
        fl[FL_STEP_REJECTED] = 0;
    }
    if (h_abs < min_step) { fl[FL_STATUS] = -1; fl[FL_DONE] = 1; return; }
    double h = h_abs;
    double t_new = t + h;
    if (t_new - ctrl[CT_TEND] > 0.0) t_new = ctrl[CT_TEND];
    h = t_new - t;
    ctrl[CT_H] = h; ctrl[CT_HABS] = fabs(h); ctrl[CT_TNEW] = t_new;
}

// accept / reject + next step size from the squared error sum `tot` over n entries (scipy rk.py:145-165); one thread
__device__ void tc_rk_control(double tot, double n, double* ctrl, int* fl) {
    const double err = sqrt(tot) / sqrt(n);
    ctrl[CT_ERR] = err;
    double h_abs = ctrl[CT_HABS];
    const double ex = ctrl[CT_EXP];
    if (err < 1.0) {
        double factor = (err == 0.0) ? 10.0 : fmin(10.0, 0.9 * pow(err, ex));
        if (fl[FL_STEP_REJECTED]) factor = fmin(1.0, factor);
        h_abs *= factor;
        ctrl[CT_HABS] = h_abs;
        ctrl[CT_T] = ctrl[CT_TNEW];
        fl[FL_ACC] += 1; fl[FL_ACC_THIS] = 1; fl[FL_ACC_LAST] = 1;
        const double t = ctrl[CT_T], te = ctrl[CT_TEND];
        // store_results condition (thermca/solver.py:313): skip counter or numpy.isclose(t, t_end)
        const bool close = fabs(t - te) <= (1e-8 + 1e-5 * fabs(te));
        if (fl[FL_NEXT_COLL] == fl[FL_RESULT_STEP] || close) {
            fl[FL_STORE_SLOT] = fl[FL_STORED];
            fl[FL_STORED] += 1;
            fl[FL_NEXT_COLL] += fl[FL_NUM_SKIP] + 1;
        } else {
            fl[FL_STORE_SLOT] = -1;
        }
        fl[FL_RESULT_STEP] += 1;
        if (t - te >= 0.0) fl[FL_DONE] = 2;    // finished; raised to 1 by the accept kernel
    } else {
        h_abs *= fmax(0.2, 0.9 * pow(err, ex));
        ctrl[CT_HABS] = h_abs;
        fl[FL_REJ] += 1; fl[FL_STEP_REJECTED] = 1; fl[FL_ACC_THIS] = 0; fl[FL_ACC_LAST] = 0;
        fl[FL_STORE_SLOT] = -1;
    }
}

// error norm + accept / reject + next step size   (scipy rk.py:145-165, common.py:63-65)
// err_i = h * sum_j E_j K_j[i]; scale = atol + max(|y|,|ynew|) rtol
__global__ void tc_error_control_kernel(long long n, const double* __restrict__ y,
                                        const double* __restrict__ ynew, TcLin E, double* ctrl,
                                        int* fl, double* __restrict__ partial) {
    if (fl[FL_DONE]) return;
    __shared__ double sh[32];
    const double h = ctrl[CT_H], rtol = ctrl[CT_RTOL], atol = ctrl[CT_ATOL];
    double s = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 7; ++j)
            if (j < E.nk) acc += E.k[j][i] * E.c[j];
        const double sc = atol + fmax(fabs(y[i]), fabs(ynew[i])) * rtol;
        const double e = acc * h / sc;
        s += e * e;
    }
    const double bs = block_sum(s, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = bs;
    if (!last_block_arrives(reinterpret_cast<unsigned int*>(fl + FL_COUNTER))) return;
    const double tot = ordered_partial_sum(partial, gridDim.x, sh);
    if (threadIdx.x != 0) return;
    tc_rk_control(tot, (double)n, ctrl, fl);
}

// on accept: y <- ynew, K0 <- K_last (FSAL), history snapshot
__global__ void tc_accept_kernel(long long n, long long u, double* __restrict__ y,
                                 const double* __restrict__ ynew, double* __restrict__ k0,
                                 const double* __restrict__ klast, const int* __restrict__ fl,
                                 double* __restrict__ hist_io, const long long* __restrict__ probe_idx,
                                 int n_probe) {
    if (fl[FL_DONE] == 1 || !fl[FL_ACC_THIS]) return;
    const int slot = fl[FL_STORE_SLOT];
    const long long nio = n - u;
    const bool full_rows = probe_idx == nullptr;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const double v = ynew[i];
        y[i] = v;
        if (k0) k0[i] = klast[i];
        if (full_rows && slot >= 0 && hist_io && i >= u) hist_io[(long long)slot * nio + (i - u)] = v;
    }
    // probe mode: a history row holds only the selected part DOFs (indices into the part block of the state)
    if (!full_rows && slot >= 0 && hist_io)
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_probe;
             i += (long long)gridDim.x * blockDim.x)
            hist_io[(long long)slot * n_probe + i] = ynew[u + probe_idx[i]];
}
// ---- the same norms on a row-partitioned network (explicit pairs under torchrun) -------------------------------------
// The state of a rank is [replicated entries | LOCAL rows of the partitioned body [lo, hi) | replicated entries]; the
// reference's norms are means over the GLOBAL state.  Partial sums are kept per class; rank 0 alone contributes the
// replicated class (its grouping of those entries differs between ranks because hi does), the sums are traded with
// tc_dist_sum_exchange (rank order), so every rank takes bit-identical accept / reject and step-size decisions.
//   MODE 0: error norm, bank 0/1 = replicated/body;  MODE 1: the two select_initial_step norms, banks 0/1 and 2/3
template <int MODE>
__global__ void tc_dist_norm_partial_kernel(long long n, long long lo, long long hi, const double* __restrict__ y,
                                            const double* __restrict__ ynew, TcLin E, const double* __restrict__ a,
                                            const double* __restrict__ b, const double* __restrict__ bsub,
                                            const double* __restrict__ ctrl, double* __restrict__ partial,
                                            const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ double sh[32];
    const double h = ctrl[CT_H], rtol = ctrl[CT_RTOL], atol = ctrl[CT_ATOL];
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int cls = (i >= lo && i < hi) ? 1 : 0;
        if (MODE == 0) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < 7; ++j)
                if (j < E.nk) acc += E.k[j][i] * E.c[j];
            const double sc = atol + fmax(fabs(y[i]), fabs(ynew[i])) * rtol;
            const double e = acc * h / sc;
            s[cls] += e * e;
        } else {
            const double sc = atol + fabs(y[i]) * rtol;
            const double va = a[i] / sc;
            s[cls] += va * va;
            const double vb = (bsub ? b[i] - bsub[i] : b[i]) / sc;
            s[2 + cls] += vb * vb;
        }
    }
    for (int q = 0; q < (MODE == 0 ? 2 : 4); ++q) {
        const double r = block_sum(s[q], sh);
        if (threadIdx.x == 0) partial[q * TC_MAX_PARTIALS + blockIdx.x] = r;
        __syncthreads();
    }
}
// ex[j] = body sum j (+ replicated sum j on rank 0), ex[nv] = local body rows (+ replicated entries on rank 0)
__global__ void tc_dist_norm_fold_kernel(int G, const double* __restrict__ partial, int nv, int rank, long long n_repl,
                                         long long n_body, double* __restrict__ ex, const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ double sh[32];
    for (int j = 0; j < nv; ++j) {
        const double rep = ordered_partial_sum(partial + (2 * j) * TC_MAX_PARTIALS, G, sh);
        __syncthreads();
        const double bod = ordered_partial_sum(partial + (2 * j + 1) * TC_MAX_PARTIALS, G, sh);
        __syncthreads();
        if (threadIdx.x == 0) ex[j] = rank == 0 ? rep + bod : bod;
    }
    if (threadIdx.x == 0) ex[nv] = (double)(rank == 0 ? n_repl + n_body : n_body);
}
__global__ void tc_dist_control_kernel(const double* __restrict__ ex, double* ctrl, int* fl) {
    if (fl[FL_DONE]) return;
    tc_rk_control(ex[0], ex[1], ctrl, fl);
}
// totals -> the layout tc_init_h0_kernel / tc_init_h1_kernel read with G = 1
__global__ void tc_dist_norm_unfold_kernel(const double* __restrict__ ex, double* __restrict__ partial) {
    partial[0] = ex[0];
    partial[TC_MAX_PARTIALS] = ex[1];
}

// network snapshot + flag finalisation (single block)
__global__ void tc_snapshot_kernel(TcNetParams P, const double* __restrict__ ctrl, int* fl,
                                   double* __restrict__ hist_net, const double* __restrict__ y = nullptr) {
    if (fl[FL_DONE] == 1) return;
    if (fl[FL_ACC_THIS]) {
        const int slot = fl[FL_STORE_SLOT];
        if (slot >= 0 && hist_net) {
            const long long w = 1 + 3LL * P.N + 2LL * P.nnz;
            double* row = hist_net + slot * w;
            const double* D = P.darena;
            if (threadIdx.x == 0) row[0] = ctrl[CT_T];
            for (int i = threadIdx.x; i < P.N; i += blockDim.x) {
                // unknown nodes: the accepted state itself (temp_u[:] = solve_temp[:u], thermca/solver.py:314)
                row[1 + i] = (y && i < P.u) ? y[i] : D[P.o_T + i];
                row[1 + P.N + i] = D[P.o_capy + i];
                row[1 + 2 * P.N + i] = D[P.o_heat + i];
            }
            for (int i = threadIdx.x; i < P.nnz; i += blockDim.x) {
                row[1 + 3 * P.N + i] = D[P.o_cond + i];
                row[1 + 3 * P.N + P.nnz + i] = D[P.o_clean + i];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (fl[FL_DONE] == 2) fl[FL_DONE] = 1;
        else if (fl[FL_ACC_THIS] && fl[FL_STORED] >= fl[FL_CAPACITY]) { fl[FL_STATUS] = 2; fl[FL_DONE] = 1; }
    }
}

// generic helpers for the implicit path
__global__ void tc_copy_kernel(long long n, const double* __restrict__ a, double* __restrict__ out,
                               const int* __restrict__ done) {
    if (done && *done) return;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) out[i] = a[i];
}
// out = a + s*b
__global__ void tc_axpby_kernel(long long n, const double* __restrict__ a, double s,
                                const double* __restrict__ b, double* __restrict__ out,
                                const int* __restrict__ done) {
    if (done && *done) return;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) out[i] = a[i] + s * b[i];
}
__global__ void tc_advance_time_kernel(double* ctrl, int* fl) {
    if (fl[FL_DONE]) return;
    ctrl[CT_T] = ctrl[CT_T] + ctrl[CT_H];
    fl[FL_ACC] += 1;
}

// ------------------------------------------------------------------------------------------
// Blocks stay explicit inside the W-method while gamma*h times the Gershgorin bound G of their operator is below
// this value: h*rho(A) <= h*G < 2/1.707 = 1.17, inside the stability interval [-2, 0] of the explicit limit of ROS2
// (Heun's method) with a margin of 40 %.
#define TC_W_EXPLICIT_BOUND 2.0

struct FfePart {
    tc_ffe_desc d;
    TcSell A;
    TcSell J;                          // tdep: frozen operator used as the W-method Jacobian (n == 0: none)
    TcSurfLoad load;
    TcFilmDiag fdiag;
    double *r, *p, *q, *dinv;          // PCG workspace
    TcPcg pcg;
    const float* binv = nullptr;       // block-Jacobi: inverted 32x32 slice blocks of M + shift K (tc_set_block_precond)
    double bj_shift = 0.0;             // the shift they were built for
    tc_dist* dist = nullptr;           // row-partitioned body (tc_attach_dist): this part holds the LOCAL rows only
    int job0 = 0, njobs = 0;           // its surface-mean jobs
    bool has_jac = false;              // a film leads to a stationary node (tc_set_jac_films): always implicit in the W-methods
};
struct LpPart { tc_lp_desc d; TcSell A; const double* b; double *r, *p, *q, *dinv; TcPcg pcg; bool has_jac = false; };
struct MorPart { tc_mor_desc d; TcMor M; bool has_jac = false; };
#define TC_JAC_MAX 4

struct tc_solver {
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    TcNetParams net{};
    TcMeanJobs jobs{};
    int mean_grid = 0;
    double* mean_partials = nullptr;
    std::vector<FfePart> ffe;
    std::vector<LpPart> lp;
    std::vector<MorPart> mor;
    long long n_state = 0;
    bool finalized = false;
    double* ctrl = nullptr;
    int* flags = nullptr;
    double* y = nullptr; double* ynew = nullptr; double* ystage = nullptr;
    double* K[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double* red_partial = nullptr;     // 4 banks
    double* pcg_sc = nullptr; int* pcg_isc = nullptr; unsigned long long* pcg_prof = nullptr;
    TcGridBar bar{};                   // reduce-barrier state of the cooperative PCG kernel (gridbar.cuh)
    int* pin_flags = nullptr;          // pinned host mirror
    double* pin_ctrl = nullptr;
    // history
    double* hist_io = nullptr; double* hist_net = nullptr; int capacity = 0; int num_skip = 0;
    const long long* probe_idx = nullptr; int n_probe = 0;     // probe mode of the history (tc_set_probes)
    double* net_vec = nullptr;                                 // 4*u doubles: CG vectors of the network block solve
    int* gate_flags = nullptr; double* gate_partial = nullptr; unsigned int* gate_counter = nullptr;   // stiffness gates
    double w_bound = TC_W_EXPLICIT_BOUND;                      // env TC_W_EXPLICIT_BOUND overrides (0: always implicit)
    // RK
    int method = TC_RK45; int n_stages = 6;
    cudaGraphExec_t rk_graph = nullptr;
    // one W-method attempt (ROS2 / ROW3w) as a CUDA graph: captured at the first advance after a begin, dropped whenever
    // something baked into its kernel arguments changes (history, probes, Jacobian terms, preconditioner, tolerances)
    cudaGraphExec_t w_graph = nullptr;
    int w_graph_kind = 0;              // 2: ROS2, 3: ROW3w
    bool w_graph_off = false;          // capture refused by the runtime (or TC_W_GRAPH=0): direct launches
    long long w_attempt_kernels = 0, w_attempt_nfev = 0;
    long long nfev = 0;            // host-side count of enqueued RHS pipelines (diagnostic)
    long long nfev_base = 0;       // evaluations outside step attempts (initial snapshot, f0, step probe)
    int evals_per_attempt = 0;
    long long launches = 0;        // kernels launched by this solver (host count)
    std::vector<cudaEvent_t> ev;   // (start, stop) pairs around PCG launches when timing is on
    bool timing = false;
    int vec_grid = 148;
    long long attempt_kernels = 0;
    // implicit
    double pcg_rtol = 1e-10; int pcg_maxit = 1000; int pcg_mode = TC_PCG_COOP; int coop_grid = 0;
    double ros_gamma = 1.0 + 0.70710678118654752440;
    bool has_dist = false;             // a part is row-partitioned (tc_attach_dist)
    long long dist_rows_global = 0;    // rows of that body over all ranks (tc_set_dist_global_rows): enables the explicit pairs
    bool dist_full_rhs = false;        // enqueue_rhs forms f = b - A y for the partitioned body (explicit pairs); false:
                                       // only b (the fused theta kernel of dist.cu does the product itself)
    double* dist_ex = nullptr;         // 4 doubles traded by tc_dist_sum_exchange (norm sums + row count)
    // rank-one Jacobian terms of films that lead to stationary nodes (tc_set_jac_films)
    int jac_n = 0;
    int jac_rec[TC_JAC_MAX][8] = {};
    double* jac_Y[TC_JAC_MAX] = {nullptr, nullptr, nullptr, nullptr};   // B^-1 u_j (state length)
    double* jac_U = nullptr;           // u_j (state length)
    double* jac_part = nullptr;        // (TC_JAC_MAX + 1) x mean_grid partial sums of c^T Y_j and c^T z
    double* jac_ginv = nullptr;        // (I - s C^T Y)^-1, K x K
};

static void drop_graphs(tc_solver* s) {
    if (s->rk_graph) { cudaGraphExecDestroy(s->rk_graph); s->rk_graph = nullptr; }
    if (s->w_graph) { cudaGraphExecDestroy(s->w_graph); s->w_graph = nullptr; }
}

static int grid_for(long long n, int threads, int cap) {
    long long g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

extern "C" int tc_create(tc_handle* out, void* cuda_stream) {
    tc_solver* s = new tc_solver();
    s->stream = (cudaStream_t)cuda_stream;
    int dev = 0;
    TC_CUDA(cudaGetDevice(&dev));
    TC_CUDA(cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, dev));
    TC_CUDA(cudaMalloc(&s->ctrl, sizeof(double) * 32));
    TC_CUDA(cudaMalloc(&s->flags, sizeof(int) * 32));
    TC_CUDA(cudaMemset(s->ctrl, 0, sizeof(double) * 32));
    TC_CUDA(cudaMemset(s->flags, 0, sizeof(int) * 32));
    TC_CUDA(cudaMalloc(&s->red_partial, sizeof(double) * 4 * TC_MAX_PARTIALS));
    TC_CUDA(cudaMalloc(&s->dist_ex, sizeof(double) * 8));
    TC_CUDA(cudaMalloc(&s->pcg_sc, sizeof(double) * 16));
    TC_CUDA(cudaMalloc(&s->pcg_isc, sizeof(int) * 16));
    TC_CUDA(cudaMemset(s->pcg_sc, 0, sizeof(double) * 16));
    TC_CUDA(cudaMemset(s->pcg_isc, 0, sizeof(int) * 16));
    TC_CUDA(cudaMalloc(&s->pcg_prof, sizeof(unsigned long long) * 16));
    TC_CUDA(cudaMemset(s->pcg_prof, 0, sizeof(unsigned long long) * 16));
    TC_CUDA(cudaMalloc(&s->bar.ctr, sizeof(unsigned int) * TC_BAR_CTR_WORDS));
    TC_CUDA(cudaMemset(s->bar.ctr, 0, sizeof(unsigned int) * TC_BAR_CTR_WORDS));
    TC_CUDA(cudaMalloc(&s->bar.part, sizeof(double) * TC_BAR_PART_DOUBLES));
    TC_CUDA(cudaMalloc(&s->bar.res, sizeof(double) * 2 * TC_BAR_MAXK));
    TC_CUDA(cudaMemset(s->bar.res, 0, sizeof(double) * 2 * TC_BAR_MAXK));
    TC_CUDA(cudaMallocHost(&s->pin_flags, sizeof(int) * 32));
    TC_CUDA(cudaMallocHost(&s->pin_ctrl, sizeof(double) * 32));
    *out = s;
    return 0;
}

extern "C" int tc_destroy(tc_handle s) {
    if (!s) return 0;
    drop_graphs(s);
    cudaFree(s->ctrl); cudaFree(s->flags); cudaFree(s->red_partial); cudaFree(s->dist_ex); cudaFree(s->pcg_sc);
    cudaFree(s->pcg_isc); cudaFree(s->pcg_prof); cudaFree(s->mean_partials);
    cudaFree(s->bar.ctr); cudaFree(s->bar.part); cudaFree(s->bar.res);
    cudaFree(s->y); cudaFree(s->ynew); cudaFree(s->ystage);
    for (int i = 0; i < 7; ++i) cudaFree(s->K[i]);
    for (int j = 0; j < TC_JAC_MAX; ++j) cudaFree(s->jac_Y[j]);
    cudaFree(s->jac_U); cudaFree(s->jac_part); cudaFree(s->jac_ginv);
    cudaFree(s->net_vec); cudaFree(s->gate_flags); cudaFree(s->gate_partial); cudaFree(s->gate_counter);
    for (auto& p : s->ffe) { cudaFree(p.r); cudaFree(p.p); cudaFree(p.q); cudaFree(p.dinv); }
    for (auto& p : s->lp) { cudaFree(p.r); cudaFree(p.p); cudaFree(p.q); cudaFree(p.dinv); }
    cudaFreeHost(s->pin_flags); cudaFreeHost(s->pin_ctrl);
    delete s;
    return 0;
}

extern "C" int tc_set_network(tc_handle s, const tc_net_desc* d) {
    TC_CHECK(s && d, "null argument");
    TcNetParams& P = s->net;
    P.iarena = d->iarena; P.darena = d->darena;
    P.cmds = d->iarena + d->cmds_off;
    P.prog_dir = d->iarena + d->prog_dir_off;
    P.tab_dir = d->iarena + d->tab_dir_off;
    P.n_iarena = d->n_iarena; P.n_darena = d->n_darena;
    P.N = d->N; P.u = d->u; P.nnz = d->nnz;
    P.o_T = d->o_T; P.o_heat = d->o_heat; P.o_capy = d->o_capy; P.o_vol = d->o_vol;
    P.o_volcapy = d->o_volcapy; P.o_condy = d->o_condy; P.o_cond = d->o_cond; P.o_clean = d->o_clean;
    P.o_calc = d->o_calc; P.o_inputs = d->o_inputs;
    P.o_indptr = d->o_indptr; P.o_indices = d->o_indices; P.o_diag = d->o_diag;
    s->jobs.njobs = d->n_mean_jobs;
    s->jobs.job_of_block = d->job_of_block; s->jobs.first_block = d->first_block; s->jobs.nblk = d->nblk;
    s->jobs.ent_start = (const long long*)d->ent_start; s->jobs.idx = d->ent_idx; s->jobs.w = d->ent_w;
    s->mean_grid = d->n_mean_blocks;
    if (s->mean_grid > 0) TC_CUDA(cudaMalloc(&s->mean_partials, sizeof(double) * s->mean_grid));
    return 0;
}

extern "C" int tc_add_ffe_part(tc_handle s, const tc_ffe_desc* d) {
    TC_CHECK(s && d && !s->finalized, "bad state");
    FfePart p{};
    p.d = *d;
    p.A = TcSell{d->n, d->nslices, (const long long*)d->slice_ptr, d->col, d->val, (const short*)d->col16};
    if (d->tdep && d->j_val)
        p.J = TcSell{d->n, d->j_nslices, (const long long*)d->j_slice_ptr, d->j_col, d->j_val, nullptr};
    p.load = TcSurfLoad{d->load_nrows, d->load_row, d->load_start, d->load_surf, d->load_bval};
    p.fdiag = TcFilmDiag{d->fd_nrows, (const long long*)d->fd_pos, d->fd_row, d->fd_abody, d->fd_start,
                         d->fd_film, d->fd_adata};
    s->ffe.push_back(p);
    return 0;
}
extern "C" int tc_add_lp_part(tc_handle s, const tc_lp_desc* d) {
    TC_CHECK(s && d && !s->finalized, "bad state");
    LpPart p{};
    p.d = *d;
    p.A = TcSell{d->n, d->nslices, (const long long*)(s->net.iarena + d->slice_ptr_ioff64),
                 s->net.iarena + d->col_ioff, s->net.darena + d->val_doff, nullptr};
    p.b = s->net.darena + d->b_doff;
    s->lp.push_back(p);
    return 0;
}
extern "C" int tc_add_mor_part(tc_handle s, const tc_mor_desc* d) {
    TC_CHECK(s && d && !s->finalized, "bad state");
    MorPart p{};
    p.d = *d;
    p.M = TcMor{d->S, d->r, d->F, d->A_body, d->A_films, d->B_T};
    s->mor.push_back(p);
    return 0;
}

extern "C" int tc_finalize(tc_handle s, int64_t n_state) {
    TC_CHECK(s && !s->finalized, "bad state");
    s->n_state = n_state;
    const size_t bytes = sizeof(double) * (size_t)(n_state > 0 ? n_state : 1);
    TC_CUDA(cudaMalloc(&s->y, bytes));
    TC_CUDA(cudaMalloc(&s->ynew, bytes));
    TC_CUDA(cudaMalloc(&s->ystage, bytes));
    for (int i = 0; i < 7; ++i) TC_CUDA(cudaMalloc(&s->K[i], bytes));
    s->vec_grid = grid_for(n_state, 256, s->sm_count * 8);
    for (auto& p : s->ffe) {
        const size_t b = sizeof(double) * (size_t)p.d.n;
        TC_CUDA(cudaMalloc(&p.r, b)); TC_CUDA(cudaMalloc(&p.p, b));
        TC_CUDA(cudaMalloc(&p.q, b)); TC_CUDA(cudaMalloc(&p.dinv, b));
    }
    for (auto& p : s->lp) {
        const size_t b = sizeof(double) * (size_t)p.d.n;
        TC_CUDA(cudaMalloc(&p.r, b)); TC_CUDA(cudaMalloc(&p.p, b));
        TC_CUDA(cudaMalloc(&p.q, b)); TC_CUDA(cudaMalloc(&p.dinv, b));
    }
    if (s->net.u > 0) TC_CUDA(cudaMalloc(&s->net_vec, sizeof(double) * 4 * (size_t)s->net.u));
    {
        const size_t nparts = s->ffe.size() + s->lp.size() + 1;
        TC_CUDA(cudaMalloc(&s->gate_flags, sizeof(int) * nparts));
        TC_CUDA(cudaMemset(s->gate_flags, 0, sizeof(int) * nparts));
        TC_CUDA(cudaMalloc(&s->gate_partial, sizeof(double) * 256));
        TC_CUDA(cudaMalloc(&s->gate_counter, sizeof(unsigned int)));
        TC_CUDA(cudaMemset(s->gate_counter, 0, sizeof(unsigned int)));
    }
    s->coop_grid = pcg_coop_grid(s->sm_count);
    s->finalized = true;
    return 0;
}

// ---- right-hand side pipeline (thermca/solver.py:193-302) ---------------------------------
#ifndef TC_RHS_COL16_DEFAULT
#define TC_RHS_COL16_DEFAULT 1
#endif
static int enqueue_rhs(tc_solver* s, double c_stage, const double* yin, double* out) {
    cudaStream_t st = s->stream;
    const int* done = s->flags + FL_DONE;
    if (s->mean_grid > 0)
        tc_surf_mean_kernel<<<s->mean_grid, TC_MEAN_THREADS, 0, st>>>(s->jobs, yin, s->mean_partials, done);
    for (auto& p : s->ffe)               // row-partitioned bodies: add the other ranks' rows to the surface means
        if (p.dist && tc_dist_mean_exchange(p.dist, s->mean_partials, s->jobs.first_block, s->jobs.nblk, p.job0, p.njobs))
            return -1;
    tc_network_kernel<<<1, TC_NET_THREADS, 0, st>>>(s->net, s->ctrl, c_stage, yin, out, s->mean_partials, done);
    for (auto& p : s->lp) {
        const int g = grid_for(p.A.nslices, TC_SPMV_THREADS / 32, s->sm_count * 8);
        tc_sell_spmv_kernel<SPMV_B><<<g, TC_SPMV_THREADS, 0, st>>>(
            p.A, yin + p.d.y_off, out + p.d.y_off, p.b, nullptr, s->ctrl, 0.0, nullptr, done);
    }
    for (auto& p : s->ffe) {
        if (p.d.tdep) {
            TcTdep T{p.A, p.d.mass, s->net.tab_dir, s->net.darena, p.d.condy_tab, p.d.volcapy_tab};
            const int g = grid_for(p.A.nslices, TC_SPMV_THREADS / 32, s->sm_count * 8);
            tc_ffe_tdep_spmv_kernel<<<g, TC_SPMV_THREADS, 0, st>>>(T, yin + p.d.y_off, out + p.d.y_off, done);
            if (p.d.ts_nrows > 0) {
                TcTdepSurf L{p.d.ts_nrows, p.d.ts_row, p.d.ts_start, p.d.ts_surf, p.d.ts_qval, p.d.ts_fstart,
                             p.d.ts_film, p.d.ts_adata};
                tc_ffe_tdep_surf_kernel<<<grid_for(p.d.ts_nrows, 256, s->sm_count * 4), 256, 0, st>>>(
                    L, T, s->net.darena + p.d.flux_doff, s->net.darena + p.d.films_doff, yin + p.d.y_off,
                    out + p.d.y_off, done);
            }
            continue;
        }
        if (p.fdiag.nrows > 0 && p.d.fd_rhs)
            tc_ffe_film_diag_kernel<<<grid_for(p.fdiag.nrows, 256, s->sm_count * 4), 256, 0, st>>>(
                p.fdiag, s->net.darena + p.d.films_doff, p.d.val, p.d.adiag, done);
        if (p.dist) {
            // row-partitioned body: the product needs the peers' ghost rows and happens inside the fused kernel of
            // dist.cu (tc_theta_steps); here only the load vector b = B flux of the local rows is formed
            TC_CUDA(cudaMemsetAsync(out + p.d.y_off, 0, sizeof(double) * p.d.n, st));
            if (p.load.nrows > 0)
                tc_ffe_surf_load_kernel<<<grid_for(p.load.nrows, 256, s->sm_count * 4), 256, 0, st>>>(
                    p.load, s->net.darena + p.d.flux_doff, out + p.d.y_off, done);
            // explicit pairs: f = b - A y with the peers' ghost rows of THIS stage vector (halo inside the kernel)
            if (s->dist_full_rhs) {
                if (tc_dist_residual(p.dist, yin + p.d.y_off, out + p.d.y_off, done)) return -1;
                s->launches += 2;
            }
            continue;
        }
        const int g = grid_for(p.A.nslices, TC_SPMV_THREADS / 32, s->sm_count * 8);
        TcSell Ar = p.A;
        // 16-bit relative columns in the right-hand side product too (TC_RHS_COL16=0 switches back): with the per-entry
        // tail loop of round 1 they lost 3.6 %, with the grouped tail and the indices one group ahead (parts.cuh) an
        // explicit RK45 attempt on the 10 M mesh takes 2.56 instead of 2.74 ms (same bits: the columns are the same)
        static const int rhs_col16 = [] { const char* e = getenv("TC_RHS_COL16"); return e ? atoi(e) : TC_RHS_COL16_DEFAULT; }();
        if (!rhs_col16) Ar.col16 = nullptr;
        tc_sell_spmv_kernel<SPMV_NEG><<<g, TC_SPMV_THREADS, 0, st>>>(
            Ar, yin + p.d.y_off, out + p.d.y_off, nullptr, nullptr, s->ctrl, 0.0, nullptr, done);
        if (p.load.nrows > 0)
            tc_ffe_surf_load_kernel<<<grid_for(p.load.nrows, 256, s->sm_count * 4), 256, 0, st>>>(
                p.load, s->net.darena + p.d.flux_doff, out + p.d.y_off, done);
    }
    for (auto& p : s->mor) {
        const int rows = p.M.S * p.M.r;
        tc_mor_rhs_kernel<<<grid_for(rows, 8, s->sm_count * 4), 256, 0, st>>>(
            p.M, s->net.darena + p.d.films_doff, s->net.darena + p.d.flux_doff, yin + p.d.y_off,
            out + p.d.y_off, done);
    }
    s->nfev += 1;
    s->launches += (s->mean_grid > 0 ? 1 : 0) + 1 + (long long)s->lp.size() + (long long)s->mor.size();
    for (auto& p : s->ffe)
        s->launches += p.d.tdep ? 1 + (p.d.ts_nrows > 0 ? 1 : 0) : 1 + (p.fdiag.nrows > 0 && p.d.fd_rhs ? 1 : 0) + (p.load.nrows > 0 ? 1 : 0);
    return 0;
}

static int set_ctrl(tc_solver* s, int idx, double v) {
    TC_CUDA(cudaMemcpyAsync(s->ctrl + idx, &v, sizeof(double), cudaMemcpyHostToDevice, s->stream));
    TC_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int tc_set_time(tc_handle s, double t) { return set_ctrl(s, CT_T, t); }

extern "C" int tc_rhs(tc_handle s, double t, const double* y_dev, double* ydot_dev) {
    TC_CHECK(s && s->finalized, "solver not finalized");
    double c2[2] = {t, 0.0};
    TC_CUDA(cudaMemcpyAsync(s->ctrl + CT_T, c2, sizeof c2, cudaMemcpyHostToDevice, s->stream));
    TC_CUDA(cudaStreamSynchronize(s->stream));
    // row-partitioned body: the complete f = b - A y of the local rows once the explicit pairs are enabled
    // (tc_set_dist_global_rows), otherwise only the load vector b (what the fused theta kernel consumes); collective
    const bool keep = s->dist_full_rhs;
    s->dist_full_rhs = s->has_dist && s->dist_rows_global > 0;
    const int rc = enqueue_rhs(s, 0.0, y_dev, ydot_dev);
    s->dist_full_rhs = keep;
    if (rc) return -1;
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int tc_set_state(tc_handle s, const double* y_dev) {
    TC_CHECK(s && s->finalized, "solver not finalized");
    TC_CUDA(cudaMemcpyAsync(s->y, y_dev, sizeof(double) * s->n_state, cudaMemcpyDeviceToDevice, s->stream));
    return 0;
}
extern "C" int tc_get_state(tc_handle s, double* y_dev) {
    TC_CHECK(s && s->finalized, "solver not finalized");
    TC_CUDA(cudaMemcpyAsync(y_dev, s->y, sizeof(double) * s->n_state, cudaMemcpyDeviceToDevice, s->stream));
    TC_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}
// enqueue-only variant for pipelined host<->device streaming (no synchronisation; ordered on the solver stream)
extern "C" int tc_get_state_async(tc_handle s, double* y_dev) {
    TC_CHECK(s && s->finalized, "solver not finalized");
    TC_CUDA(cudaMemcpyAsync(y_dev, s->y, sizeof(double) * s->n_state, cudaMemcpyDeviceToDevice, s->stream));
    return 0;
}
extern "C" int tc_set_probes(tc_handle s, const int64_t* probe_idx_dev, int32_t n_probe) {
    TC_CHECK(s, "null handle");
    TC_CHECK(n_probe >= 0, "negative probe count");
    s->probe_idx = n_probe > 0 ? (const long long*)probe_idx_dev : nullptr;
    s->n_probe = n_probe > 0 ? n_probe : 0;
    drop_graphs(s);
    return 0;
}

extern "C" int tc_set_history(tc_handle s, double* hist_io, double* hist_net, int32_t capacity, int32_t num_skip) {
    TC_CHECK(s, "null");
    s->hist_io = hist_io; s->hist_net = hist_net; s->capacity = capacity; s->num_skip = num_skip;
    drop_graphs(s);
    return 0;
}
extern "C" int tc_reset_history(tc_handle s) {
    int zero = 0;
    TC_CUDA(cudaMemcpyAsync(s->flags + FL_STORED, &zero, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    // a "history full" stop is resumable
    TC_CUDA(cudaMemcpyAsync(s->pin_flags, s->flags, sizeof(int) * FL_N, cudaMemcpyDeviceToHost, s->stream));
    TC_CUDA(cudaStreamSynchronize(s->stream));
    if (s->pin_flags[FL_STATUS] == 2) {
        int two[1] = {0};
        TC_CUDA(cudaMemcpyAsync(s->flags + FL_STATUS, two, sizeof(int), cudaMemcpyHostToDevice, s->stream));
        TC_CUDA(cudaMemcpyAsync(s->flags + FL_DONE, two, sizeof(int), cudaMemcpyHostToDevice, s->stream));
        TC_CUDA(cudaStreamSynchronize(s->stream));
    }
    return 0;
}

// ---- Runge-Kutta tableaus (scipy/integrate/_ivp/rk.py: RK23, RK45) ----------------------
struct Tableau { int n_stages; double C[7]; double A[7][6]; double B[7]; double E[7]; double order; };
static Tableau tableau(int method) {
    Tableau T{};
    if (method == TC_RK45) {
        T.n_stages = 6; T.order = 4;
        const double C[6] = {0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1};
        const double A[6][5] = {{0, 0, 0, 0, 0},
                                {1.0 / 5, 0, 0, 0, 0},
                                {3.0 / 40, 9.0 / 40, 0, 0, 0},
                                {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
                                {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
                                {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
        const double B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
        const double E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
        for (int i = 0; i < 6; ++i) { T.C[i] = C[i]; T.B[i] = B[i]; for (int j = 0; j < 5; ++j) T.A[i][j] = A[i][j]; }
        for (int i = 0; i < 7; ++i) T.E[i] = E[i];
    } else {
        T.n_stages = 3; T.order = 2;
        const double C[3] = {0, 1.0 / 2, 3.0 / 4};
        const double A[3][2] = {{0, 0}, {1.0 / 2, 0}, {0, 3.0 / 4}};
        const double B[3] = {2.0 / 9, 1.0 / 3, 4.0 / 9};
        const double E[4] = {5.0 / 72, -1.0 / 12, -1.0 / 9, 1.0 / 8};
        for (int i = 0; i < 3; ++i) { T.C[i] = C[i]; T.B[i] = B[i]; for (int j = 0; j < 2; ++j) T.A[i][j] = A[i][j]; }
        for (int i = 0; i < 4; ++i) T.E[i] = E[i];
    }
    return T;
}

// norms of the adaptive pairs on a row-partitioned network: class-split partial sums -> fold -> exchange (dist.cu)
static const FfePart* dist_part(const tc_solver* s) {
    for (auto& p : s->ffe) if (p.dist) return &p;
    return nullptr;
}
template <int MODE>
static int enqueue_dist_norm(tc_solver* s, const double* ynew, const TcLin& E, const double* a, const double* b,
                             const double* bsub, const int* done) {
    const FfePart* p = dist_part(s);
    const long long n = s->n_state, lo = p->d.y_off, hi = p->d.y_off + p->d.n;
    const int G = s->vec_grid, nv = MODE == 0 ? 1 : 2;
    tc_dist_norm_partial_kernel<MODE><<<G, 256, 0, s->stream>>>(n, lo, hi, s->y, ynew, E, a, b, bsub, s->ctrl, s->red_partial, done);
    tc_dist_norm_fold_kernel<<<1, 256, 0, s->stream>>>(G, s->red_partial, nv, tc_dist_rank(p->dist), n - p->d.n, p->d.n,
                                                       s->dist_ex, done);
    if (tc_dist_sum_exchange(p->dist, s->dist_ex, nv + 1, done)) return -1;
    s->launches += 3;
    return 0;
}

static int enqueue_rk_attempt(tc_solver* s) {
    cudaStream_t st = s->stream;
    const Tableau T = tableau(s->method);
    const int ns = T.n_stages;
    const int* done = s->flags + FL_DONE;
    const long long n = s->n_state;
    tc_prepare_kernel<<<1, 1, 0, st>>>(s->ctrl, s->flags);
    for (int sg = 1; sg < ns; ++sg) {
        TcLin L{};
        L.nk = sg;
        for (int j = 0; j < sg; ++j) { L.k[j] = s->K[j]; L.c[j] = T.A[sg][j]; }
        tc_lin_kernel<<<s->vec_grid, 256, 0, st>>>(n, s->y, L, s->ctrl, s->ystage, done);
        if (enqueue_rhs(s, T.C[sg], s->ystage, s->K[sg])) return -1;
    }
    TcLin Lb{};
    Lb.nk = ns;
    for (int j = 0; j < ns; ++j) { Lb.k[j] = s->K[j]; Lb.c[j] = T.B[j]; }
    tc_lin_kernel<<<s->vec_grid, 256, 0, st>>>(n, s->y, Lb, s->ctrl, s->ynew, done);
    if (enqueue_rhs(s, 1.0, s->ynew, s->K[ns])) return -1;
    TcLin Le{};
    Le.nk = ns + 1;
    for (int j = 0; j <= ns; ++j) { Le.k[j] = s->K[j]; Le.c[j] = T.E[j]; }
    if (s->has_dist) {
        if (enqueue_dist_norm<0>(s, s->ynew, Le, nullptr, nullptr, nullptr, done)) return -1;
        tc_dist_control_kernel<<<1, 1, 0, st>>>(s->dist_ex, s->ctrl, s->flags);
    } else {
        tc_error_control_kernel<<<s->vec_grid, 256, 0, st>>>(n, s->y, s->ynew, Le, s->ctrl, s->flags, s->red_partial);
    }
    tc_accept_kernel<<<s->vec_grid, 256, 0, st>>>(n, s->net.u, s->y, s->ynew, s->K[0], s->K[ns], s->flags, s->hist_io, s->probe_idx, s->n_probe);
    tc_snapshot_kernel<<<1, 256, 0, st>>>(s->net, s->ctrl, s->flags, s->hist_net);
    s->launches += ns + 4;   // stage combinations, prepare, error control, accept, snapshot
    return 0;
}

static int fetch_flags(tc_solver* s) {
    TC_CUDA(cudaMemcpyAsync(s->pin_flags, s->flags, sizeof(int) * FL_N, cudaMemcpyDeviceToHost, s->stream));
    TC_CUDA(cudaMemcpyAsync(s->pin_ctrl, s->ctrl, sizeof(double) * CT_N, cudaMemcpyDeviceToHost, s->stream));
    TC_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

static int begin_common(tc_solver* s, double t0, double t_end, double rtol, double atol, double max_step,
                        double exponent) {
    double c[CT_N];
    memset(c, 0, sizeof c);
    c[CT_T] = t0; c[CT_TEND] = t_end; c[CT_RTOL] = rtol; c[CT_ATOL] = atol;
    c[CT_MAXSTEP] = max_step > 0 ? max_step : INFINITY;
    c[CT_EXP] = exponent; c[CT_INTERVAL] = fabs(t_end - t0);
    int f[FL_N];
    memset(f, 0, sizeof f);
    f[FL_ACC_LAST] = 1; f[FL_NEXT_COLL] = s->num_skip; f[FL_NUM_SKIP] = s->num_skip;
    f[FL_CAPACITY] = s->capacity > 0 ? s->capacity : 0x7fffffff; f[FL_STORE_SLOT] = -1;
    TC_CUDA(cudaMemcpyAsync(s->ctrl, c, sizeof c, cudaMemcpyHostToDevice, s->stream));
    TC_CUDA(cudaMemcpyAsync(s->flags, f, sizeof f, cudaMemcpyHostToDevice, s->stream));
    TC_CUDA(cudaStreamSynchronize(s->stream));
    s->nfev = 0;
    return 0;
}

// initial snapshot (thermca/solver.py:329-330): rhs at (t0, y0), then store slot 0
static int store_initial(tc_solver* s) {
    if (!s->hist_io && !s->hist_net) return 0;
    int f[3] = {1, 0, 0};   // ACC_THIS = 1, STORED(after) handled below
    int slot0 = 0, one = 1;
    TC_CUDA(cudaMemcpyAsync(s->flags + FL_ACC_THIS, f, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    TC_CUDA(cudaMemcpyAsync(s->flags + FL_STORE_SLOT, &slot0, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    TC_CUDA(cudaMemcpyAsync(s->flags + FL_STORED, &one, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    tc_accept_kernel<<<s->vec_grid, 256, 0, s->stream>>>(s->n_state, s->net.u, s->ynew, s->y, nullptr, nullptr,
                                                          s->flags, s->hist_io, s->probe_idx, s->n_probe);
    tc_snapshot_kernel<<<1, 256, 0, s->stream>>>(s->net, s->ctrl, s->flags, s->hist_net);
    int zero = 0, m1 = -1;
    TC_CUDA(cudaMemcpyAsync(s->flags + FL_ACC_THIS, &zero, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    TC_CUDA(cudaMemcpyAsync(s->flags + FL_STORE_SLOT, &m1, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    TC_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int tc_rk_begin(tc_handle s, int32_t method, double t0, double t_end, double rtol, double atol,
                           double max_step, double first_step) {
    TC_CHECK(s && s->finalized, "solver not finalized");
    TC_CHECK(!s->has_dist || s->dist_rows_global > 0, "explicit pairs on a row-partitioned body need tc_set_dist_global_rows");
    TC_CHECK(method == TC_RK45 || method == TC_RK23, "unknown method");
    s->dist_full_rhs = s->has_dist;
    if (s->rk_graph && s->method != method) { cudaGraphExecDestroy(s->rk_graph); s->rk_graph = nullptr; }
    s->method = method;
    const Tableau T = tableau(method);
    s->n_stages = T.n_stages;
    if (begin_common(s, t0, t_end, rtol, atol, max_step, -1.0 / (T.order + 1.0))) return -1;
    s->evals_per_attempt = T.n_stages;
    s->nfev_base = 1;
    cudaStream_t st = s->stream;
    const long long n = s->n_state;
    const int G = s->vec_grid;
    // transient_solution's own first evaluation + initial snapshot (solver.py:329-330)
    if (enqueue_rhs(s, 0.0, s->y, s->K[0])) return -1;
    if (store_initial(s)) return -1;
    if (!(t_end > t0)) {
        int one = 1;
        TC_CUDA(cudaMemcpyAsync(s->flags + FL_DONE, &one, sizeof(int), cudaMemcpyHostToDevice, st));
        TC_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    // scipy RungeKutta.__init__: f = fun(t0, y0); h_abs = select_initial_step(...)
    if (enqueue_rhs(s, 0.0, s->y, s->K[0])) return -1;
    s->nfev_base += 1;
    if (first_step > 0) {
        if (set_ctrl(s, CT_HABS, first_step)) return -1;
    } else if (n == 0) {
        if (set_ctrl(s, CT_HABS, INFINITY)) return -1;
    } else {
        // partitioned network: global sums through the exchange, then the same two kernels on the totals (G = 1)
        const long long ng = s->has_dist ? n - dist_part(s)->d.n + s->dist_rows_global : n;
        const int Gn = s->has_dist ? 1 : G;
        const TcLin none{};
        if (s->has_dist) {
            if (enqueue_dist_norm<1>(s, nullptr, none, s->y, s->K[0], nullptr, nullptr)) return -1;
            tc_dist_norm_unfold_kernel<<<1, 1, 0, st>>>(s->dist_ex, s->red_partial);
        } else {
            tc_norm2_kernel<<<G, 256, 0, st>>>(n, s->y, s->y, s->K[0], nullptr, s->ctrl, s->red_partial, nullptr);
        }
        tc_init_h0_kernel<<<1, 256, 0, st>>>(ng, Gn, s->red_partial, s->ctrl);
        TcLin L{};
        L.nk = 1; L.k[0] = s->K[0]; L.c[0] = 1.0;
        tc_lin_kernel<<<G, 256, 0, st>>>(n, s->y, L, s->ctrl, s->ystage, nullptr);   // y1 = y0 + h0 f0
        if (enqueue_rhs(s, 1.0, s->ystage, s->K[1])) return -1;                        // f1 at t0 + h0
        s->nfev_base += 1;
        if (s->has_dist) {
            if (enqueue_dist_norm<1>(s, nullptr, none, s->y, s->K[1], s->K[0], nullptr)) return -1;
            tc_dist_norm_unfold_kernel<<<1, 1, 0, st>>>(s->dist_ex, s->red_partial);
        } else {
            tc_norm2_kernel<<<G, 256, 0, st>>>(n, s->y, s->y, s->K[1], s->K[0], s->ctrl, s->red_partial, nullptr);
        }
        tc_init_h1_kernel<<<1, 256, 0, st>>>(ng, Gn, s->red_partial, s->ctrl, T.order);
    }
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int tc_rk_advance(tc_handle s, int32_t max_attempts, int32_t poll_every) {
    TC_CHECK(s && s->finalized, "solver not finalized");
    cudaStream_t st = s->stream;
    if (!s->rk_graph) {
        cudaGraph_t graph;
        const long long nfev0 = s->nfev, l0 = s->launches;
        TC_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_rk_attempt(s);
        TC_CUDA(cudaStreamEndCapture(st, &graph));
        s->nfev = nfev0;
        s->attempt_kernels = s->launches - l0;
        s->launches = l0;
        TC_CHECK(rc == 0, "capture failed");
        TC_CUDA(cudaGraphInstantiate(&s->rk_graph, graph, 0));
        TC_CUDA(cudaGraphDestroy(graph));
    }
    if (poll_every <= 0) poll_every = 16;
    int launched = 0;
    while (launched < max_attempts) {
        int burst = poll_every < max_attempts - launched ? poll_every : max_attempts - launched;
        // never outrun the history ring: each attempt stores at most one snapshot
        for (int i = 0; i < burst; ++i) TC_CUDA(cudaGraphLaunch(s->rk_graph, st));
        launched += burst;
        s->nfev += (long long)burst * s->n_stages;
        s->launches += (long long)burst * s->attempt_kernels;
        if (fetch_flags(s)) return -1;
        if (s->pin_flags[FL_DONE]) break;
    }
    TC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tc_get_info(tc_handle s, int64_t* info8, double* t_now, double* h_next) {
    TC_CHECK(s, "null");
    if (fetch_flags(s)) return -1;
    info8[0] = s->pin_flags[FL_DONE]; info8[1] = s->pin_flags[FL_ACC]; info8[2] = s->pin_flags[FL_REJ];
    info8[3] = s->pin_flags[FL_STATUS]; info8[4] = s->pin_flags[FL_STORED];
    info8[5] = s->nfev_base + (long long)s->evals_per_attempt * (s->pin_flags[FL_ACC] + s->pin_flags[FL_REJ]);
    info8[6] = s->pin_flags[FL_RESULT_STEP]; info8[7] = s->launches;
    if (t_now) *t_now = s->pin_ctrl[CT_T];
    if (h_next) *h_next = s->pin_ctrl[CT_HABS];
    return 0;
}

extern "C" int tc_graph_info(tc_handle s, int32_t* out2) {
    TC_CHECK(s && out2, "null");
    out2[0] = s->rk_graph ? 1 : 0;
    out2[1] = s->w_graph ? s->w_graph_kind : (s->w_graph_off ? -1 : 0);
    return 0;
}

// ---- PCG drivers ---------------------------------------------------------------------------------
static int pcg_occ_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TC_PCG_OCC");
        v = e ? atoi(e) : TC_PCG_DEFAULT_OCC;
        if (v != 5 && v != 6 && v != 8) v = TC_PCG_DEFAULT_OCC;
    }
    return v;
}
// block shape of the cooperative kernel: 1024 threads x 2 blocks/SM by default; TC_PCG_THREADS = 512 / 256 selects
// 4 / 8 blocks per SM (same warps per SM; fewer participants per grid barrier are cheaper: 1 M DOF 716 -> 820 M
// DOF*steps/s, 10 M DOF 412 -> 419, profiles/README.md)
static int pcg_coop_threads() {
    static int t = -1;
    if (t < 0) {
        const char* e = getenv("TC_PCG_THREADS");
        t = e ? atoi(e) : TC_PCG_COOP_THREADS_DEFAULT;
        if (t != 256 && t != 512 && t != 768 && t != 1024) t = TC_PCG_COOP_THREADS_DEFAULT;
    }
    return t;
}
static void* pcg_coop_fn() {
    const int t = pcg_coop_threads();
    if (t == 1024) return (void*)tc_pcg_coop_kernel<2, 1024>;
    if (t == 768) return (void*)tc_pcg_coop_kernel<2, 768>;          // 42 registers, 48 warps per SM
    if (t == 512) return (void*)tc_pcg_coop_kernel<4, 512>;
    switch (pcg_occ_variant()) {
        case 5: return (void*)tc_pcg_coop_kernel<5>;
        case 6: return (void*)tc_pcg_coop_kernel<6>;
        default: return (void*)tc_pcg_coop_kernel<8>;
    }
}
static int pcg_coop_grid(int sm_count) {
    int occ = 0;
    const int t = pcg_coop_threads();
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)pcg_coop_fn(), t, 0);
    if (occ > 2048 / t) occ = 2048 / t;
    if (occ < 1) occ = 1;
    int g = occ * sm_count;
    if (g > TC_MAX_PARTIALS) g = TC_MAX_PARTIALS;
    return g;
}

static int launch_pcg(tc_solver* s, TcPcg& P, int mode, int coop_grid, const int* done, const float* binv = nullptr) {
    cudaStream_t st = s ? s->stream : nullptr;
    // CG terminates in at most n steps in exact arithmetic: never spin on a tolerance below the
    // attainable residual of a tiny system
    if ((long long)P.maxit > 2LL * P.A.n + 50) P.maxit = 2 * P.A.n + 50;
    if (mode == TC_PCG_COOP_BJ) {
        TC_CHECK(binv, "block-Jacobi PCG needs the inverted blocks (tc_set_block_precond)");
        double* zvec = P.dinv;                      // the point-Jacobi factor array is free in this variant: holds z
        void* args[] = {(void*)&P, (void*)&binv, (void*)&zvec, (void*)&done};
        const int cthreads = pcg_coop_threads();
        int g = tc_balanced_grid(P.A.nslices, cthreads / 32, coop_grid);
        void* fn = cthreads == 1024 ? (void*)tc_pcg_coop_bj_kernel<2, 1024>
                   : cthreads == 512 ? (void*)tc_pcg_coop_bj_kernel<4, 512> : (void*)tc_pcg_coop_bj_kernel<8, 256>;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (s && s->timing) {
            TC_CUDA(cudaEventCreate(&e0)); TC_CUDA(cudaEventCreate(&e1));
            TC_CUDA(cudaEventRecord(e0, st));
        }
        TC_CUDA(cudaLaunchCooperativeKernel(fn, dim3(g), dim3(cthreads), args, 0, st));
        if (e0) { TC_CUDA(cudaEventRecord(e1, st)); s->ev.push_back(e0); s->ev.push_back(e1); }
        if (s) s->launches += 1;
        return 0;
    }
    if (mode == TC_PCG_COOP) {
        void* args[] = {(void*)&P, (void*)&done};
        int g = coop_grid;
        const int cthreads = pcg_coop_threads();
        g = tc_balanced_grid(P.A.nslices, cthreads / 32, g);
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (s && s->timing) {
            TC_CUDA(cudaEventCreate(&e0)); TC_CUDA(cudaEventCreate(&e1));
            TC_CUDA(cudaEventRecord(e0, st));
        }
        TC_CUDA(cudaLaunchCooperativeKernel(pcg_coop_fn(), dim3(g), dim3(cthreads), args, 0, st));
        if (e0) { TC_CUDA(cudaEventRecord(e1, st)); s->ev.push_back(e0); s->ev.push_back(e1); }
        if (s) s->launches += 1;
        return 0;
    }
    const int gv = grid_for(P.A.n, TC_PCG_THREADS, TC_MAX_PARTIALS);
    const int gs = grid_for(P.A.nslices, TC_PCG_THREADS / 32, TC_MAX_PARTIALS);
    tc_pcg_init_kernel<<<gv, TC_PCG_THREADS, 0, st>>>(P, done);
    tc_pcg_scalar_kernel<<<1, TC_PCG_THREADS, 0, st>>>(P, 0, gv, done);
    int flag = 0;
    for (int it = 0; it < P.maxit; ++it) {
        tc_pcg_k1<<<gs, TC_PCG_THREADS, 0, st>>>(P, it, done);
        tc_pcg_k2<<<gv, TC_PCG_THREADS, 0, st>>>(P, it, gs, done);
        tc_pcg_k3<<<gv, TC_PCG_THREADS, 0, st>>>(P, it, gv, done);
        tc_pcg_scalar_kernel<<<1, TC_PCG_THREADS, 0, st>>>(P, 1, gv, done);
        if ((it % 8) == 7) {
            TC_CUDA(cudaMemcpyAsync(&flag, P.isc, sizeof(int), cudaMemcpyDeviceToHost, st));
            TC_CUDA(cudaStreamSynchronize(st));
            if (flag) break;
        }
    }
    return 0;
}

// z = (I + coef*h*A_p)^-1 rhs_p on FE parts, z = rhs elsewhere
// Unknown network nodes inside the W-method: z_u = (I + s C_uu^-1 L_uu)^-1 rhs_u, i.e. (C + s L_uu) z = C rhs with
// the conductance matrix in the differential-equation form of the latest right-hand side (calc, rows < u:
// diagonal = sum of the row's conductances, off-diagonal = -G; thermca/solver.py:45-58,252-255).  C + s L_uu is
// symmetric positive definite; u <= ~10^3, so one CTA runs Jacobi-CG over the CSR rows (columns >= u are the
// coupling to bound nodes and part surfaces, which stays explicit).  vec: 4*u doubles of scratch.
__global__ void __launch_bounds__(256)
tc_net_block_solve_kernel(TcNetParams P, const double* __restrict__ ctrl, double coef,
                          const double* __restrict__ rhs, double* __restrict__ z, double* __restrict__ vec,
                          double rtol, const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ double sh[32];
    __shared__ double bc[2];
    const int u = P.u, tid = threadIdx.x, nt = blockDim.x;
    const int* indptr = P.iarena + P.o_indptr;
    const int* indices = P.iarena + P.o_indices;
    const double* calc = P.darena + P.o_calc;
    const double* capy = P.darena + P.o_capy;
    const int* dg = P.iarena + P.o_diag;
    const double sft = coef * ctrl[1];
    double *r = vec, *p = vec + u, *q = vec + 2 * u, *dinv = vec + 3 * u;
    double rho_l = 0.0, bb_l = 0.0;
    for (int i = tid; i < u; i += nt) {
        const double b = capy[i] * rhs[i];
        const double d = 1.0 / (capy[i] + sft * calc[dg[i]]);
        r[i] = b; z[i] = 0.0; dinv[i] = d; p[i] = d * b;
        rho_l += b * d * b; bb_l += b * b;
    }
    auto bsum = [&](double v, int slot) {
        const double t = block_sum(v, sh);
        if (tid == 0) bc[slot] = t;
        __syncthreads();
        const double out = bc[slot];
        __syncthreads();
        return out;
    };
    double rho = bsum(rho_l, 0);
    const double bb = bsum(bb_l, 1);
    if (!(bb > 0.0)) return;
    const double tol2 = rtol * rtol * bb;
    const int maxit = 2 * u + 20;
    for (int it = 0; it < maxit; ++it) {
        double pq_l = 0.0;
        for (int i = tid; i < u; i += nt) {
            double acc = 0.0;
            for (int j = indptr[i]; j < indptr[i + 1]; ++j) {
                const int c = indices[j];
                if (c < u) acc += calc[j] * p[c];
            }
            const double qi = capy[i] * p[i] + sft * acc;
            q[i] = qi;
            pq_l += p[i] * qi;
        }
        const double pq = bsum(pq_l, 0);
        const double alpha = rho / pq;
        double rn_l = 0.0, rr_l = 0.0;
        for (int i = tid; i < u; i += nt) {
            z[i] += alpha * p[i];
            const double ri = r[i] - alpha * q[i];
            r[i] = ri;
            rn_l += ri * dinv[i] * ri; rr_l += ri * ri;
        }
        const double rho_new = bsum(rn_l, 0);
        const double rr = bsum(rr_l, 1);
        if (!(rr > tol2)) break;
        const double beta = rho_new / rho;
        rho = rho_new;
        for (int i = tid; i < u; i += nt) p[i] = dinv[i] * r[i] + beta * p[i];
        __syncthreads();
    }
}

// MOR blocks inside the W-method: (I + s A_q) z_q = rhs_q for every surface q of a reduced body, with
// A_q = A_body[q] + sum_f film_f A_films[f][q] (r x r, symmetric positive definite; thermca/fem/mor_io.py:90-92) applied on
// the fly.  One CTA per surface runs Jacobi-CG with its vectors in shared memory (5 r doubles).
__global__ void __launch_bounds__(256)
tc_mor_block_solve_kernel(TcMor M, const double* __restrict__ films, const double* __restrict__ ctrl, double coef,
                          double bound, const double* __restrict__ rhs, double* __restrict__ z, double rtol,
                          const int* __restrict__ done) {
    if (done && *done) return;
    extern __shared__ double smv[];
    __shared__ double sh[32];
    __shared__ double bc[2];
    const int r = M.r, q_s = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const size_t blk = (size_t)M.S * r * r;
    const double* ab = M.A_body + (size_t)q_s * r * r;
    const double* af = M.A_films + (size_t)q_s * r * r;
    const double* b = rhs + (size_t)q_s * r;
    double* x = z + (size_t)q_s * r;
    double *rv = smv, *p = smv + r, *q = smv + 2 * r, *dinv = smv + 3 * r, *xs = smv + 4 * r;
    const double sft = coef * ctrl[1];
    auto a_at = [&](int i, int j) {
        double v = ab[(size_t)i * r + j];
        for (int f = 0; f < M.F; ++f) v += films[f] * af[f * blk + (size_t)i * r + j];
        return v;
    };
    auto bsum = [&](double v, int slot) {
        const double t = block_sum(v, sh);
        if (tid == 0) bc[slot] = t;
        __syncthreads();
        const double out = bc[slot];
        __syncthreads();
        return out;
    };
    // A W-method may use any Jacobian approximation, also none: when the step is well inside the explicit
    // stability region of this block (Gershgorin bound of s*A below TC_W_EXPLICIT_BOUND) the copy made by the caller (z = rhs)
    // stands and the block is explicit for this attempt - measured: fewer steps than damping non-stiff modes.
    {
        double g_l = 0.0;
        for (int i = warp; i < r; i += nw) {
            double acc = 0.0;
            for (int j = lane; j < r; j += 32) acc += fabs(a_at(i, j));
            acc = warp_sum(acc);
            if (lane == 0) g_l = fmax(g_l, acc);
        }
        g_l = warp_max(g_l);            // lane 0 of every warp holds its maximum; reduce over warps
        __syncthreads();
        if (lane == 0) sh[warp] = g_l;
        __syncthreads();
        double g = 0.0;
        for (int w = 0; w < nw; ++w) g = fmax(g, sh[w]);
        __syncthreads();
        if (sft * g < bound) return;      // uniform over the CTA
    }
    double rho_l = 0.0, bb_l = 0.0;
    for (int i = tid; i < r; i += nt) {
        const double bi = b[i];
        const double d = 1.0 / (1.0 + sft * a_at(i, i));
        rv[i] = bi; xs[i] = 0.0; dinv[i] = d; p[i] = d * bi;
        rho_l += bi * d * bi; bb_l += bi * bi;
    }
    double rho = bsum(rho_l, 0);
    const double bb = bsum(bb_l, 1);
    if (bb > 0.0) {
        const double tol2 = rtol * rtol * bb;
        const int maxit = 2 * r + 20;
        for (int it = 0; it < maxit; ++it) {
            for (int i = warp; i < r; i += nw) {            // q = p + s A p, one warp per row
                double acc = 0.0;
                for (int j = lane; j < r; j += 32) acc += a_at(i, j) * p[j];
                acc = warp_sum(acc);
                if (lane == 0) q[i] = p[i] + sft * acc;
            }
            __syncthreads();
            double pq_l = 0.0;
            for (int i = tid; i < r; i += nt) pq_l += p[i] * q[i];
            const double alpha = rho / bsum(pq_l, 0);
            double rn_l = 0.0, rr_l = 0.0;
            for (int i = tid; i < r; i += nt) {
                xs[i] += alpha * p[i];
                const double ri = rv[i] - alpha * q[i];
                rv[i] = ri;
                rn_l += ri * dinv[i] * ri; rr_l += ri * ri;
            }
            const double rho_new = bsum(rn_l, 0);
            const double rr = bsum(rr_l, 1);
            if (!(rr > tol2)) break;
            const double beta = rho_new / rho;
            rho = rho_new;
            for (int i = tid; i < r; i += nt) p[i] = dinv[i] * rv[i] + beta * p[i];
            __syncthreads();
        }
    }
    __syncthreads();
    for (int i = tid; i < r; i += nt) x[i] = xs[i];
}

// Stiffness gate of one PCG block inside the W-method: with A = M^-1 K diagonally dominant, 2 max_i A_ii bounds
// its spectral radius.  While s * 2 max A_ii < TC_W_EXPLICIT_BOUND the step is inside the explicit stability region and the
// block is left explicit for this attempt (z = rhs, the copy the caller made): damping non-stiff modes with an
// inexact Jacobian only costs accuracy (measured on the temperature dependent LP golden: error 7e-3 -> 1e-6 at
// the same tolerance).  This is what LSODA's Adams/BDF switch does for the reference.
__global__ void __launch_bounds__(256)
tc_pcg_gate_kernel(long long n, const double* __restrict__ adiag, const double* __restrict__ ctrl, double coef,
                   double bound, double* __restrict__ partial, unsigned int* __restrict__ counter,
                   int* __restrict__ skip, const int* __restrict__ done, double* __restrict__ gmax_out = nullptr) {
    if (done && *done) return;
    __shared__ double sh[32];
    double m = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmax(m, adiag[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, sh[w]);
        partial[blockIdx.x] = m;
    }
    if (last_block_arrives(counter)) {
        if (threadIdx.x == 0) {
            double g = 0.0;
            for (unsigned int b = 0; b < gridDim.x; ++b) g = fmax(g, ((volatile double*)partial)[b]);
            if (gmax_out) *gmax_out = g;               // row-partitioned body: the decision needs the max over the ranks
            else *skip = (coef * ctrl[1] * 2.0 * g < bound) ? 1 : 0;
        }
    }
}
__global__ void tc_gate_decide_kernel(const double* __restrict__ gmax, const double* __restrict__ ctrl, double coef,
                                      double bound, int* __restrict__ skip, const int* __restrict__ done) {
    if (done && *done) return;
    *skip = (coef * ctrl[1] * 2.0 * gmax[0] < bound) ? 1 : 0;
}

// w_method: the caller tolerates an inexact Jacobian (ROS2), so temperature dependent bodies are solved with
// their operator frozen at the initial temperatures; otherwise they are left explicit
static int enqueue_block_solve(tc_solver* s, double coef, const double* rhs, double* z, bool w_method = false) {
    cudaStream_t st = s->stream;
    const int* done = s->flags + FL_DONE;
    tc_copy_kernel<<<s->vec_grid, 256, 0, st>>>(s->n_state, rhs, z, done);
    int gate_ix = 0;
    if (w_method && s->net.u > 0 && s->net_vec)
        tc_net_block_solve_kernel<<<1, 256, 0, st>>>(s->net, s->ctrl, coef, rhs, z, s->net_vec, 1e-12, done);
    if (w_method)
        for (auto& p : s->mor) {
            tc_mor_block_solve_kernel<<<p.M.S, 256, sizeof(double) * 5 * p.M.r, st>>>(
                p.M, s->net.darena + p.d.films_doff, s->ctrl, coef, p.has_jac ? 0.0 : s->w_bound, rhs + p.d.y_off,
                z + p.d.y_off, 1e-12, done);
        }
    // LP parts: same SPD structure (M A = L_body + film/margin diagonal, thermca/lpm/lp_io.py:106-128)
    for (auto& p : s->lp) {
        TcPcg& P = p.pcg;
        P.A = p.A; P.mass = s->net.darena + p.d.mass_doff; P.adiag = s->net.darena + p.d.adiag_doff;
        P.rhs = rhs + p.d.y_off; P.x = z + p.d.y_off;
        P.r = p.r; P.p = p.p; P.q = p.q; P.dinv = p.dinv;
        P.partial = s->red_partial; P.sc = s->pcg_sc; P.isc = s->pcg_isc;
        P.rtol = s->pcg_rtol; P.maxit = s->pcg_maxit; P.ctrl = s->ctrl; P.coef = coef;
        P.prof = nullptr;
        P.skip = nullptr;
        P.bar = s->bar;
        if (w_method && s->pcg_mode == TC_PCG_COOP && !p.has_jac) {
            int* flag = s->gate_flags + gate_ix++;
            tc_pcg_gate_kernel<<<grid_for(p.d.n, 256, 256), 256, 0, st>>>(p.d.n, P.adiag, s->ctrl, coef, s->w_bound,
                                                                          s->gate_partial, s->gate_counter, flag, done);
            P.skip = flag;
        }
        if (launch_pcg(s, P, s->pcg_mode == TC_PCG_COOP_BJ ? TC_PCG_COOP : s->pcg_mode, s->coop_grid, done)) return -1;
    }
    for (auto& p : s->ffe) {
        if (p.dist) {
            // row-partitioned body.  theta steps: advanced by the fused kernel of dist.cu (tc_theta_steps).  W-methods:
            // the stage solve runs in the solve-only instantiation of that kernel across the GPUs, always implicit
            if (w_method) {
                // stiffness gate as for the other blocks, on the largest diagonal entry over ALL ranks (rank-uniform flag)
                int* flag = s->gate_flags + gate_ix++;
                tc_pcg_gate_kernel<<<grid_for(p.d.n, 256, 256), 256, 0, st>>>(p.d.n, p.d.adiag, s->ctrl, coef, s->w_bound,
                                                                              s->gate_partial, s->gate_counter, flag, done,
                                                                              s->dist_ex + 4);
                if (tc_dist_max_exchange(p.dist, s->dist_ex + 4, 1, done)) return -1;
                tc_gate_decide_kernel<<<1, 1, 0, st>>>(s->dist_ex + 4, s->ctrl, coef, s->w_bound, flag, done);
                if (tc_dist_solve(p.dist, coef, s->ctrl + CT_H, rhs + p.d.y_off, z + p.d.y_off, s->pcg_rtol, s->pcg_maxit,
                                  flag, done))
                    return -1;
                s->launches += 4;
            }
            continue;
        }
        const bool frozen = p.d.tdep != 0;
        if (frozen && !(w_method && p.J.n > 0)) continue;      // operator depends on T: explicit (copy above)
        TcPcg& P = p.pcg;
        P.A = frozen ? p.J : p.A; P.mass = frozen ? p.d.j_mass : p.d.mass; P.adiag = frozen ? p.d.j_adiag : p.d.adiag;
        P.rhs = rhs + p.d.y_off; P.x = z + p.d.y_off;
        P.r = p.r; P.p = p.p; P.q = p.q; P.dinv = p.dinv;
        P.partial = s->red_partial; P.sc = s->pcg_sc; P.isc = s->pcg_isc;
        P.rtol = s->pcg_rtol; P.maxit = s->pcg_maxit; P.ctrl = s->ctrl; P.coef = coef;
        P.prof = s->timing ? s->pcg_prof : nullptr;
        P.skip = nullptr;
        P.bar = s->bar;
        if (w_method && s->pcg_mode == TC_PCG_COOP && !p.has_jac) {
            int* flag = s->gate_flags + gate_ix++;
            tc_pcg_gate_kernel<<<grid_for(p.d.n, 256, 256), 256, 0, st>>>(p.d.n, P.adiag, s->ctrl, coef, s->w_bound,
                                                                          s->gate_partial, s->gate_counter, flag, done);
            P.skip = flag;
        }
        if (s->pcg_mode == TC_PCG_COOP_BJ && !frozen) {
            TC_CHECK(p.binv, "pcg_mode TC_PCG_COOP_BJ: call tc_set_block_precond for every full-order FE part first");
            if (launch_pcg(s, P, TC_PCG_COOP_BJ, s->coop_grid, done, p.binv)) return -1;
        } else if (launch_pcg(s, P, s->pcg_mode == TC_PCG_COOP_BJ ? TC_PCG_COOP : s->pcg_mode, s->coop_grid, done)) {
            return -1;
        }
    }
    return 0;
}

// ---- rank-one Jacobian terms of films that lead to stationary nodes (tc_set_jac_films) ------------------------------
// W = B - s sum_j u_j c_j^T with B = blockdiag(I + s A_p) (what enqueue_block_solve inverts), u_j the load vector of
// film j times theta_j (network.cuh C_JTHETA), c_j the surface-mean weights of the film's surface.  Woodbury:
//   W^-1 r = z + Y (I - s C^T Y)^-1 s C^T z,   z = B^-1 r,   Y = B^-1 U.
// Y and the K x K inverse are formed once per attempt (enqueue_jac_prepare); every block solve of the attempt is
// followed by the correction (enqueue_jac_correct).  c^T x is the surface-mean kernel of the right-hand side applied
// to x.  CPU restatement: oracle/ros2_oracle.py _jac_prepare.
struct TcJacRanges { int K; int first[TC_JAC_MAX]; int nblk[TC_JAC_MAX]; };
struct TcJacY { const double* p[TC_JAC_MAX]; };
__global__ void tc_jac_u_lp_kernel(int dof, int F, int f, const double* __restrict__ lfm, const double* __restrict__ minv,
                                   const double* __restrict__ theta, double* __restrict__ U,
                                   const int* __restrict__ done) {
    if (done && *done) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dof; i += gridDim.x * blockDim.x)
        U[i] = theta[0] * minv[i] * lfm[i * F + f];
}
__global__ void tc_jac_u_ffe_kernel(TcFilmDiag Fd, int f, const double* __restrict__ theta, const double* __restrict__ film,
                                    double* __restrict__ U, const int* __restrict__ done) {
    if (done && *done) return;
    const double tf = theta[0] * film[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Fd.nrows; i += gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int k = Fd.start[i]; k < Fd.start[i + 1]; ++k)
            if (Fd.film[k] == f) acc += tf * Fd.adata[k];
        U[Fd.row[i]] = acc;
    }
}
__global__ void tc_jac_u_vec_kernel(int n, const double* __restrict__ v, const double* __restrict__ theta,
                                    const double* __restrict__ film, double* __restrict__ U, const int* __restrict__ done) {
    if (done && *done) return;
    const double tf = theta[0] * film[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) U[i] = tf * v[i];
}
__global__ void tc_jac_gram_kernel(TcJacRanges R, const double* __restrict__ partY, int stride, const double* __restrict__ ctrl,
                                   double coef, double* __restrict__ ginv, const int* __restrict__ done) {
    if (done && *done) return;
    if (threadIdx.x != 0) return;
    const int K = R.K;
    const double sft = coef * ctrl[CT_H];
    double G[TC_JAC_MAX][2 * TC_JAC_MAX];
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) {
            double cy = 0.0;
            for (int b = 0; b < R.nblk[i]; ++b) cy += partY[(size_t)j * stride + R.first[i] + b];
            G[i][j] = (i == j ? 1.0 : 0.0) - sft * cy;
            G[i][K + j] = i == j ? 1.0 : 0.0;
        }
    for (int c = 0; c < K; ++c) {                       // Gauss-Jordan with partial pivoting
        int piv = c;
        for (int i = c + 1; i < K; ++i) if (fabs(G[i][c]) > fabs(G[piv][c])) piv = i;
        for (int j = 0; j < 2 * K; ++j) { const double t = G[c][j]; G[c][j] = G[piv][j]; G[piv][j] = t; }
        const double d = 1.0 / G[c][c];
        for (int j = 0; j < 2 * K; ++j) G[c][j] *= d;
        for (int i = 0; i < K; ++i) {
            if (i == c) continue;
            const double m = G[i][c];
            for (int j = 0; j < 2 * K; ++j) G[i][j] -= m * G[c][j];
        }
    }
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) ginv[i * K + j] = G[i][K + j];
}
__global__ void tc_jac_apply_kernel(long long n, double* __restrict__ z, TcJacY Y, TcJacRanges R,
                                    const double* __restrict__ partZ, const double* __restrict__ ginv,
                                    const double* __restrict__ ctrl, double coef, const int* __restrict__ done) {
    if (done && *done) return;
    __shared__ double m[TC_JAC_MAX];
    const int K = R.K;
    if (threadIdx.x == 0) {
        const double sft = coef * ctrl[CT_H];
        double t[TC_JAC_MAX];
        for (int i = 0; i < K; ++i) {
            double a = 0.0;
            for (int b = 0; b < R.nblk[i]; ++b) a += partZ[R.first[i] + b];
            t[i] = sft * a;
        }
        for (int i = 0; i < K; ++i) {
            double a = 0.0;
            for (int j = 0; j < K; ++j) a += ginv[i * K + j] * t[j];
            m[i] = a;
        }
    }
    __syncthreads();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v = z[i];
        for (int j = 0; j < K; ++j) v += m[j] * Y.p[j][i];
        z[i] = v;
    }
}

static TcJacRanges jac_ranges(const tc_solver* s) {
    TcJacRanges R{};
    R.K = s->jac_n;
    for (int j = 0; j < s->jac_n; ++j) { R.first[j] = s->jac_rec[j][6]; R.nblk[j] = s->jac_rec[j][7]; }
    return R;
}

static int enqueue_jac_prepare(tc_solver* s, double coef) {
    if (s->jac_n == 0) return 0;
    cudaStream_t st = s->stream;
    const int* done = s->flags + FL_DONE;
    double* D = s->net.darena;
    for (int j = 0; j < s->jac_n; ++j) {
        const int* q = s->jac_rec[j];       // kind, part, film, surface, theta_doff, film_doff, first_block, nblk
        TC_CUDA(cudaMemsetAsync(s->jac_U, 0, sizeof(double) * (size_t)s->n_state, st));
        if (q[0] == 0) {
            LpPart& p = s->lp[q[1]];
            tc_jac_u_lp_kernel<<<grid_for(p.d.n, 256, 64), 256, 0, st>>>(p.d.n, p.d.F, q[2], D + p.d.lfm_doff, D + p.d.minv_doff,
                                                                         D + q[4], s->jac_U + p.d.y_off, done);
        } else if (q[0] == 1) {
            FfePart& p = s->ffe[q[1]];
            tc_jac_u_ffe_kernel<<<grid_for(p.fdiag.nrows, 256, s->sm_count * 4), 256, 0, st>>>(
                p.fdiag, q[2], D + q[4], D + q[5], s->jac_U + p.d.y_off, done);
        } else {
            MorPart& p = s->mor[q[1]];
            tc_jac_u_vec_kernel<<<grid_for(p.M.r, 256, 8), 256, 0, st>>>(
                p.M.r, p.M.B_T + (size_t)q[3] * p.M.r, D + q[4], D + q[5], s->jac_U + p.d.y_off + (size_t)q[3] * p.M.r, done);
        }
        if (enqueue_block_solve(s, coef, s->jac_U, s->jac_Y[j], true)) return -1;
        tc_surf_mean_kernel<<<s->mean_grid, TC_MEAN_THREADS, 0, st>>>(s->jobs, s->jac_Y[j],
                                                                      s->jac_part + (size_t)j * s->mean_grid, done);
    }
    tc_jac_gram_kernel<<<1, 32, 0, st>>>(jac_ranges(s), s->jac_part, s->mean_grid, s->ctrl, coef, s->jac_ginv, done);
    return 0;
}

// z = B^-1 r  ->  z = W^-1 r
static int enqueue_jac_correct(tc_solver* s, double coef, double* z) {
    if (s->jac_n == 0) return 0;
    cudaStream_t st = s->stream;
    const int* done = s->flags + FL_DONE;
    double* partZ = s->jac_part + (size_t)TC_JAC_MAX * s->mean_grid;
    tc_surf_mean_kernel<<<s->mean_grid, TC_MEAN_THREADS, 0, st>>>(s->jobs, z, partZ, done);
    TcJacY Y{};
    for (int j = 0; j < s->jac_n; ++j) Y.p[j] = s->jac_Y[j];
    tc_jac_apply_kernel<<<s->vec_grid, 256, 0, st>>>(s->n_state, z, Y, jac_ranges(s), partZ, s->jac_ginv, s->ctrl, coef, done);
    return 0;
}

extern "C" int tc_set_jac_films(tc_handle s, int32_t n, const int32_t* rec) {
    TC_CHECK(s && s->finalized, "solver not finalized");
    TC_CHECK(n >= 0 && n <= TC_JAC_MAX, "at most 4 films between part surfaces and stationary nodes are supported");
    TC_CHECK(n == 0 || s->mean_grid > 0, "no surface-mean jobs");
    drop_graphs(s);
    for (int j = 0; j < n; ++j) {
        const int32_t* q = rec + 8 * j;
        for (int k = 0; k < 8; ++k) s->jac_rec[j][k] = q[k];
        if (q[0] == 0) {
            TC_CHECK(q[1] >= 0 && q[1] < (int)s->lp.size() && q[2] >= 0 && q[2] < s->lp[q[1]].d.F, "LP film index");
            s->lp[q[1]].has_jac = true;
        } else if (q[0] == 1) {
            TC_CHECK(q[1] >= 0 && q[1] < (int)s->ffe.size(), "FE part index");
            FfePart& p = s->ffe[q[1]];
            TC_CHECK(p.fdiag.nrows > 0 && !p.d.tdep && !p.dist, "FE part without film lists (fd_*), temperature dependent "
                                                                "or row-partitioned");
            p.has_jac = true;
        } else {
            TC_CHECK(q[0] == 2 && q[1] >= 0 && q[1] < (int)s->mor.size() && q[3] >= 0 && q[3] < s->mor[q[1]].M.S,
                     "MOR film index");
            s->mor[q[1]].has_jac = true;
        }
        TC_CHECK(q[6] >= 0 && q[7] >= 1 && q[6] + q[7] <= s->mean_grid, "mean job range");
    }
    if (n > 0 && !s->jac_U) {
        const size_t bytes = sizeof(double) * (size_t)(s->n_state > 0 ? s->n_state : 1);
        TC_CUDA(cudaMalloc(&s->jac_U, bytes));
        for (int j = 0; j < TC_JAC_MAX; ++j) TC_CUDA(cudaMalloc(&s->jac_Y[j], bytes));
        TC_CUDA(cudaMalloc(&s->jac_part, sizeof(double) * (size_t)(TC_JAC_MAX + 1) * s->mean_grid));
        TC_CUDA(cudaMalloc(&s->jac_ginv, sizeof(double) * TC_JAC_MAX * TC_JAC_MAX));
    }
    s->jac_n = n;
    return 0;
}

// Inverted diagonal blocks for the block-Jacobi variant of FE part `part`: nslices x 32 x 32 floats, [slice][j][i],
// symmetric, built for M + shift*K (device pointer, kept by reference; the caller rebuilds them when h or a film changes)
extern "C" int tc_set_block_precond(tc_handle s, int32_t part, const float* binv_dev, double shift) {
    TC_CHECK(s && part >= 0 && part < (int)s->ffe.size(), "part index");
    s->ffe[part].binv = binv_dev;
    s->ffe[part].bj_shift = shift;
    drop_graphs(s);
    return 0;
}

// Mark full-order FE part `part` as row-partitioned: the part of this solver holds the LOCAL rows of the body (local
// surface lists, local film-diagonal lists; its SELL arrays are the ones given to tc_dist_set_operator), `d` owns the
// halo plan.  first_job / n_jobs: the part's surface-mean jobs.  Supported by tc_rhs (network rows and load vector),
// tc_theta_steps; the explicit and W-method paths refuse such a solver.
extern "C" int tc_attach_dist(tc_handle s, int32_t part, tc_dist* d, int32_t first_job, int32_t n_jobs) {
    TC_CHECK(s && part >= 0 && part < (int)s->ffe.size() && d, "part index");
    TC_CHECK(!s->ffe[part].d.tdep, "temperature dependent bodies cannot be row-partitioned");
    TC_CHECK(tc_dist_local_rows(d) == s->ffe[part].d.n, "local row count of the part and of the partition differ");
    TC_CHECK(tc_dist_stream(d) == (void*)s->stream, "the partition must use the solver's stream");
    s->ffe[part].dist = d;
    // rank-one Jacobian terms of films between this body and stationary nodes are dropped: their Woodbury correction
    // uses surface means of the LOCAL rows only; a W-method admits any Jacobian approximation (it costs steps, not order)
    if (s->ffe[part].has_jac) {
        int m = 0;
        for (int j = 0; j < s->jac_n; ++j)
            if (!(s->jac_rec[j][0] == 1 && s->jac_rec[j][1] == part)) {
                if (m != j) for (int k = 0; k < 8; ++k) s->jac_rec[m][k] = s->jac_rec[j][k];
                ++m;
            }
        s->jac_n = m;
        s->ffe[part].has_jac = false;
    }
    drop_graphs(s);
    s->ffe[part].job0 = first_job; s->ffe[part].njobs = n_jobs;
    s->has_dist = true;
    return 0;
}

extern "C" int tc_set_dist_global_rows(tc_handle s, int64_t n_rows_global) {
    TC_CHECK(s && s->has_dist, "no row-partitioned part attached");
    TC_CHECK(n_rows_global >= dist_part(s)->d.n, "global row count below the local one");
    s->dist_rows_global = n_rows_global;
    return 0;
}

extern "C" int tc_theta_steps(tc_handle s, int32_t nsteps, double theta, double h_step, double pcg_rtol,
                              int32_t pcg_maxit, int32_t pcg_mode) {
    TC_CHECK(s && s->finalized, "solver not finalized");
    s->pcg_rtol = pcg_rtol; s->pcg_maxit = pcg_maxit; s->pcg_mode = pcg_mode;
    s->dist_full_rhs = false;                       // the fused kernel of dist.cu forms b - A y itself
    if (set_ctrl(s, CT_H, h_step)) return -1;
    cudaStream_t st = s->stream;
    const int* done = s->flags + FL_DONE;
    for (int k = 0; k < nsteps; ++k) {
        if (enqueue_rhs(s, theta, s->y, s->K[0])) return -1;                 // f(t + theta h, y)
        if (enqueue_block_solve(s, theta, s->K[0], s->K[1])) return -1;      // (I + theta h A) z = f
        for (auto& p : s->ffe)           // row-partitioned bodies are advanced by their own kernel below
            if (p.dist) TC_CUDA(cudaMemsetAsync(s->K[1] + p.d.y_off, 0, sizeof(double) * p.d.n, st));
        TcLin L{};
        L.nk = 1; L.k[0] = s->K[1]; L.c[0] = 1.0;
        tc_lin_kernel<<<s->vec_grid, 256, 0, st>>>(s->n_state, s->y, L, s->ctrl, s->y, done);  // y += h z
        for (auto& p : s->ffe)
            if (p.dist) {
                // K[0] holds the load vector b(t + theta h) of the local rows; the fused kernel forms f = b - A y with the
                // peers' ghost rows, solves (I + theta h A) z = f across the GPUs and commits y += h z in place
                if (tc_dist_set_load(p.dist, s->K[0] + p.d.y_off)) return -1;
                if (tc_dist_set_state_ptr(p.dist, s->y + p.d.y_off)) return -1;
                if (tc_dist_theta_steps(p.dist, 1, theta, h_step, pcg_rtol, pcg_maxit)) return -1;
                s->launches += 1;
            }
        tc_advance_time_kernel<<<1, 1, 0, st>>>(s->ctrl, s->flags);
        s->launches += 3;   // block-solve copy, y update, time advance (+ rhs and pcg counted inside)
    }
    TC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tc_ros2_begin(tc_handle s, double t0, double t_end, double rtol, double atol, double max_step,
                             double first_step, double pcg_rtol, int32_t pcg_maxit, int32_t pcg_mode) {
    TC_CHECK(s && s->finalized, "solver not finalized");
    TC_CHECK(!s->has_dist || s->dist_rows_global > 0, "W-methods on a row-partitioned body need tc_set_dist_global_rows");
    TC_CHECK(!s->has_dist || pcg_mode == TC_PCG_COOP, "a row-partitioned body needs pcg_mode TC_PCG_COOP");
    s->dist_full_rhs = s->has_dist;
    for (auto& p : s->mor)
        TC_CHECK(p.M.r <= 1024, "a MOR block with r > 1024 cannot be solved implicitly (its CG vectors do not fit shared "
                                "memory); use an explicit pair for this model");
    s->pcg_rtol = pcg_rtol; s->pcg_maxit = pcg_maxit; s->pcg_mode = pcg_mode;
    drop_graphs(s);
    s->w_bound = TC_W_EXPLICIT_BOUND;
    if (const char* e = getenv("TC_W_EXPLICIT_BOUND")) s->w_bound = atof(e);
    if (begin_common(s, t0, t_end, rtol, atol, max_step, -1.0 / 2.0)) return -1;
    s->evals_per_attempt = 2;
    s->nfev_base = 1;
    if (enqueue_rhs(s, 0.0, s->y, s->K[0])) return -1;
    tc_copy_kernel<<<s->vec_grid, 256, 0, s->stream>>>(s->n_state, s->K[0], s->K[5], nullptr);   // "previous f(t+h, y+)"
    if (store_initial(s)) return -1;
    double h0 = first_step;
    if (!(h0 > 0)) h0 = 1e-3 * fabs(t_end - t0);
    if (set_ctrl(s, CT_HABS, h0)) return -1;
    if (!(t_end > t0)) {
        int one = 1;
        TC_CUDA(cudaMemcpyAsync(s->flags + FL_DONE, &one, sizeof(int), cudaMemcpyHostToDevice, s->stream));
        TC_CUDA(cudaStreamSynchronize(s->stream));
    }
    return 0;
}

// One ROS2 attempt (Verwer et al. 1999): gamma = 1 + 1/sqrt(2), W-method with J ~ -blockdiag(A_p)
//   (I + g h A) k1 = f(t, y);  (I + g h A) k2 = f(t + h, y + h k1) - 2 k1
//   y+ = y + 1.5 h k1 + 0.5 h k2;  err = y+ - (y + h k1) = 0.5 h (k1 + k2)
// f(t, y) of an attempt is the f(t + h, y+) the previous accepted attempt evaluated last (kept in K[5]); after a
// rejection y is unchanged and K[0] still holds it
__global__ void tc_ros2_reuse_kernel(long long n, const double* __restrict__ fend, double* __restrict__ f1,
                                     const int* __restrict__ fl) {
    if (fl[FL_DONE] == 1 || !fl[FL_ACC_LAST]) return;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        f1[i] = fend[i];
}

static int enqueue_ros2_attempt(tc_solver* s) {
    cudaStream_t st = s->stream;
    const int* done = s->flags + FL_DONE;
    const long long n = s->n_state;
    const int G = s->vec_grid;
    tc_ros2_reuse_kernel<<<G, 256, 0, st>>>(n, s->K[5], s->K[0], s->flags);        // before prepare clears nothing it needs
    tc_prepare_kernel<<<1, 1, 0, st>>>(s->ctrl, s->flags);
    if (enqueue_jac_prepare(s, s->ros_gamma)) return -1;
    if (enqueue_block_solve(s, s->ros_gamma, s->K[0], s->K[1], true)) return -1;    // k1
    if (enqueue_jac_correct(s, s->ros_gamma, s->K[1])) return -1;
    TcLin L1{};
    L1.nk = 1; L1.k[0] = s->K[1]; L1.c[0] = 1.0;
    tc_lin_kernel<<<G, 256, 0, st>>>(n, s->y, L1, s->ctrl, s->ystage, done);       // y + h k1
    if (enqueue_rhs(s, 1.0, s->ystage, s->K[2])) return -1;
    tc_axpby_kernel<<<G, 256, 0, st>>>(n, s->K[2], -2.0, s->K[1], s->K[3], done);  // f2 - 2 k1
    if (enqueue_block_solve(s, s->ros_gamma, s->K[3], s->K[4], true)) return -1;    // k2
    if (enqueue_jac_correct(s, s->ros_gamma, s->K[4])) return -1;
    TcLin L2{};
    L2.nk = 2; L2.k[0] = s->K[1]; L2.c[0] = 1.5; L2.k[1] = s->K[4]; L2.c[1] = 0.5;
    tc_lin_kernel<<<G, 256, 0, st>>>(n, s->y, L2, s->ctrl, s->ynew, done);
    // f(t + h, y+): next attempt's f(t, y) if this one is accepted, and the network state the snapshot stores
    // (surface means, conductances, sources at the new state - as after the FSAL stage of the explicit pairs)
    if (enqueue_rhs(s, 1.0, s->ynew, s->K[5])) return -1;
    TcLin Le{};
    Le.nk = 2; Le.k[0] = s->K[1]; Le.c[0] = 0.5; Le.k[1] = s->K[4]; Le.c[1] = 0.5;
    if (s->has_dist) {
        if (enqueue_dist_norm<0>(s, s->ynew, Le, nullptr, nullptr, nullptr, done)) return -1;
        tc_dist_control_kernel<<<1, 1, 0, st>>>(s->dist_ex, s->ctrl, s->flags);
    } else {
        tc_error_control_kernel<<<G, 256, 0, st>>>(n, s->y, s->ynew, Le, s->ctrl, s->flags, s->red_partial);
    }
    tc_accept_kernel<<<G, 256, 0, st>>>(n, s->net.u, s->y, s->ynew, nullptr, nullptr, s->flags, s->hist_io, s->probe_idx,
                                        s->n_probe);
    tc_snapshot_kernel<<<1, 256, 0, st>>>(s->net, s->ctrl, s->flags, s->hist_net, s->y);
    return 0;
}

// W-method attempts are launch bound on small models (ROS2: ~25, ROW3w: ~45 kernels of a few microseconds each, several
// of them cooperative launches): one attempt is captured once per begin and replayed as a CUDA graph.  Not captured:
// the phase-per-launch PCG mode (it polls a device flag from the host), PCG timing (events per launch), and runtimes
// that refuse a cooperative launch inside a capture - those fall back to direct launches, same kernels, same order.
static int w_advance(tc_solver* s, int kind, int (*enqueue)(tc_solver*), int max_attempts, int poll_every) {
    TC_CHECK(s && s->finalized, "solver not finalized");
    cudaStream_t st = s->stream;
    if (poll_every <= 0) poll_every = 4;
    bool use_graph = !s->w_graph_off && !s->timing && s->pcg_mode != TC_PCG_MULTI;
    if (const char* e = getenv("TC_W_GRAPH")) use_graph = use_graph && atoi(e) != 0;   // A/B switch (scripts, tests)
    if (use_graph && s->w_graph && s->w_graph_kind != kind) { cudaGraphExecDestroy(s->w_graph); s->w_graph = nullptr; }
    if (use_graph && !s->w_graph) {
        cudaGraph_t graph = nullptr;
        const long long nfev0 = s->nfev, l0 = s->launches;
        TC_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue(s);
        const cudaError_t ec = cudaStreamEndCapture(st, &graph);
        s->w_attempt_nfev = s->nfev - nfev0; s->w_attempt_kernels = s->launches - l0;
        s->nfev = nfev0; s->launches = l0;
        cudaError_t ei = cudaErrorUnknown;
        if (rc == 0 && ec == cudaSuccess && graph) ei = cudaGraphInstantiate(&s->w_graph, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (ei != cudaSuccess) {
            cudaGetLastError();                                 // capture refused: nothing ran, fall back for good
            s->w_graph = nullptr; s->w_graph_off = true; use_graph = false;
        } else {
            s->w_graph_kind = kind;
        }
    }
    int launched = 0;
    while (launched < max_attempts) {
        int burst = poll_every < max_attempts - launched ? poll_every : max_attempts - launched;
        for (int i = 0; i < burst; ++i) {
            if (use_graph) {
                TC_CUDA(cudaGraphLaunch(s->w_graph, st));
                s->nfev += s->w_attempt_nfev; s->launches += s->w_attempt_kernels;
            } else if (enqueue(s)) {
                return -1;
            }
        }
        launched += burst;
        if (fetch_flags(s)) return -1;
        if (s->pin_flags[FL_DONE]) break;
    }
    TC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tc_ros2_advance(tc_handle s, int32_t max_attempts, int32_t poll_every) {
    return w_advance(s, 2, enqueue_ros2_attempt, max_attempts, poll_every);
}

// ---- third-order W-method "ROW3w" (4 stages, 3 right-hand sides; derivation: DESIGN.md section 3, CPU restatement:
// oracle/row3_oracle.py).  Order 3 for ANY Jacobian approximation, so the block structure of the ROS2 path carries
// over unchanged: FE / LP parts by Jacobi-PCG, unknown network nodes and MOR blocks by their small CG kernels, frozen
// operators for temperature dependent bodies, blocks left explicit while gamma*h*rho is inside the stability interval
// of the explicit limit (a third-order explicit Runge-Kutta method, real interval 2.51; bound 0.5 on gamma*h*2 max A_ii).
//   W u_i = g (F_i + sum_{j<i} c_ij u_j),  F_i = f(t + alpha_i h, y + h sum_j at_ij u_j),  F_4 = F_3,  W = I + g h A
//   y+ = y + h sum m_i u_i,  err = h sum e_i u_i
#define TC_ROW3_GAMMA 0.435866521508459
#define TC_ROW3_BOUND 0.5
__global__ void tc_stage_rhs_kernel(long long n, const double* __restrict__ F, TcLin L, double g, double* __restrict__ out,
                                    const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v = F[i];
        for (int j = 0; j < L.nk; ++j) v += L.c[j] * L.k[j][i];
        out[i] = g * v;
    }
}

extern "C" int tc_row3_begin(tc_handle s, double t0, double t_end, double rtol, double atol, double max_step,
                             double first_step, double pcg_rtol, int32_t pcg_maxit, int32_t pcg_mode) {
    TC_CHECK(s && s->finalized, "solver not finalized");
    TC_CHECK(!s->has_dist || s->dist_rows_global > 0, "W-methods on a row-partitioned body need tc_set_dist_global_rows");
    TC_CHECK(!s->has_dist || pcg_mode == TC_PCG_COOP, "a row-partitioned body needs pcg_mode TC_PCG_COOP");
    s->dist_full_rhs = s->has_dist;
    for (auto& p : s->mor)
        TC_CHECK(p.M.r <= 1024, "a MOR block with r > 1024 cannot be solved implicitly (its CG vectors do not fit shared "
                                "memory); use an explicit pair for this model");
    s->pcg_rtol = pcg_rtol; s->pcg_maxit = pcg_maxit; s->pcg_mode = pcg_mode;
    drop_graphs(s);
    s->w_bound = TC_ROW3_BOUND;
    if (const char* e = getenv("TC_W_EXPLICIT_BOUND")) s->w_bound = atof(e);
    if (begin_common(s, t0, t_end, rtol, atol, max_step, -1.0 / 3.0)) return -1;
    s->evals_per_attempt = 3;
    s->nfev_base = 1;
    if (enqueue_rhs(s, 0.0, s->y, s->K[0])) return -1;
    tc_copy_kernel<<<s->vec_grid, 256, 0, s->stream>>>(s->n_state, s->K[0], s->K[5], nullptr);   // "previous f(t+h, y+)"
    if (store_initial(s)) return -1;
    double h0 = first_step;
    if (!(h0 > 0)) h0 = 1e-3 * fabs(t_end - t0);
    if (set_ctrl(s, CT_HABS, h0)) return -1;
    if (!(t_end > t0)) {
        int one = 1;
        TC_CUDA(cudaMemcpyAsync(s->flags + FL_DONE, &one, sizeof(int), cudaMemcpyHostToDevice, s->stream));
        TC_CUDA(cudaStreamSynchronize(s->stream));
    }
    return 0;
}

static int enqueue_row3_attempt(tc_solver* s) {
    cudaStream_t st = s->stream;
    const int* done = s->flags + FL_DONE;
    const long long n = s->n_state;
    const int G = s->vec_grid;
    const double g = TC_ROW3_GAMMA;
    double *F1 = s->K[0], *u1 = s->K[1], *Fs = s->K[2], *rhs = s->K[3], *u2 = s->K[4], *fend = s->K[5], *u3 = s->K[6];
    double* u4 = s->K[2];                          // F_3 is consumed before u_4 is written
    tc_ros2_reuse_kernel<<<G, 256, 0, st>>>(n, fend, F1, s->flags);
    tc_prepare_kernel<<<1, 1, 0, st>>>(s->ctrl, s->flags);
    if (enqueue_jac_prepare(s, g)) return -1;
    TcLin L{};
    // stage 1
    L.nk = 0;
    tc_stage_rhs_kernel<<<G, 256, 0, st>>>(n, F1, L, g, rhs, done);
    if (enqueue_block_solve(s, g, rhs, u1, true)) return -1;
    if (enqueue_jac_correct(s, g, u1)) return -1;
    // stage 2
    L.nk = 1; L.k[0] = u1; L.c[0] = 1.1021597568860195;
    tc_lin_kernel<<<G, 256, 0, st>>>(n, s->y, L, s->ctrl, s->ystage, done);
    if (enqueue_rhs(s, 0.48039453938051824, s->ystage, Fs)) return -1;
    L.c[0] = -3.3064792706580586;
    tc_stage_rhs_kernel<<<G, 256, 0, st>>>(n, Fs, L, g, rhs, done);
    if (enqueue_block_solve(s, g, rhs, u2, true)) return -1;
    if (enqueue_jac_correct(s, g, u2)) return -1;
    // stage 3
    L.nk = 2; L.k[0] = u1; L.c[0] = 2.0458152611220233; L.k[1] = u2; L.c[1] = 1.1251523859013772;
    tc_lin_kernel<<<G, 256, 0, st>>>(n, s->y, L, s->ctrl, s->ystage, done);
    if (enqueue_rhs(s, 0.6753387630276377, s->ystage, Fs)) return -1;
    L.c[0] = -3.1394810593440803; L.c[1] = -2.5814150212946312;
    tc_stage_rhs_kernel<<<G, 256, 0, st>>>(n, Fs, L, g, rhs, done);
    if (enqueue_block_solve(s, g, rhs, u3, true)) return -1;
    if (enqueue_jac_correct(s, g, u3)) return -1;
    // stage 4 (same right-hand side F_3)
    L.nk = 3; L.k[0] = u1; L.c[0] = -4.6371650178370949; L.k[1] = u2; L.c[1] = -3.1668182208148394;
    L.k[2] = u3; L.c[2] = -2.0579924311095659;
    tc_stage_rhs_kernel<<<G, 256, 0, st>>>(n, Fs, L, g, rhs, done);
    if (enqueue_block_solve(s, g, rhs, u4, true)) return -1;
    if (enqueue_jac_correct(s, g, u4)) return -1;
    // new state, f(t + h, y+), error estimate, accept / reject, snapshot
    TcLin Ly{};
    Ly.nk = 4; Ly.k[0] = u1; Ly.k[1] = u2; Ly.k[2] = u3; Ly.k[3] = u4;
    Ly.c[0] = 3.6624707580663167; Ly.c[1] = 2.2151642855284215; Ly.c[2] = 1.5089213118874722; Ly.c[3] = 1.1081682922152405;
    tc_lin_kernel<<<G, 256, 0, st>>>(n, s->y, Ly, s->ctrl, s->ynew, done);
    if (enqueue_rhs(s, 1.0, s->ynew, fend)) return -1;
    TcLin Le = Ly;
    Le.c[0] = 0.25290167107671513; Le.c[1] = 0.04754813669718994; Le.c[2] = 0.23046773281334043; Le.c[3] = 0.4960290010539171;
    if (s->has_dist) {
        if (enqueue_dist_norm<0>(s, s->ynew, Le, nullptr, nullptr, nullptr, done)) return -1;
        tc_dist_control_kernel<<<1, 1, 0, st>>>(s->dist_ex, s->ctrl, s->flags);
    } else {
        tc_error_control_kernel<<<G, 256, 0, st>>>(n, s->y, s->ynew, Le, s->ctrl, s->flags, s->red_partial);
    }
    tc_accept_kernel<<<G, 256, 0, st>>>(n, s->net.u, s->y, s->ynew, nullptr, nullptr, s->flags, s->hist_io, s->probe_idx,
                                        s->n_probe);
    tc_snapshot_kernel<<<1, 256, 0, st>>>(s->net, s->ctrl, s->flags, s->hist_net, s->y);
    return 0;
}

extern "C" int tc_row3_advance(tc_handle s, int32_t max_attempts, int32_t poll_every) {
    return w_advance(s, 3, enqueue_row3_attempt, max_attempts, poll_every);
}

extern "C" int tc_timing(tc_handle s, int32_t enable, double* pcg_ms, int64_t* pcg_launches, int64_t* launches) {
    TC_CHECK(s, "null");
    TC_CUDA(cudaStreamSynchronize(s->stream));
    double ms = 0.0;
    for (size_t i = 0; i + 1 < s->ev.size(); i += 2) {
        float f = 0.f;
        TC_CUDA(cudaEventElapsedTime(&f, s->ev[i], s->ev[i + 1]));
        ms += f;
        cudaEventDestroy(s->ev[i]); cudaEventDestroy(s->ev[i + 1]);
    }
    if (pcg_ms) *pcg_ms = ms;
    if (pcg_launches) *pcg_launches = (int64_t)(s->ev.size() / 2);
    if (launches) *launches = s->launches;
    s->ev.clear();
    s->timing = enable != 0;
    s->launches = 0;
    return 0;
}

// phase profile of the cooperative PCG kernel (block 0, ns; recorded while timing is on; cleared on read):
// spmv, sync, pq sum, update, sync, rho sums, direction, sync
extern "C" int tc_pcg_profile(tc_handle s, uint64_t* out16) {
    TC_CUDA(cudaStreamSynchronize(s->stream));
    TC_CUDA(cudaMemcpy(out16, s->pcg_prof, sizeof(unsigned long long) * 16, cudaMemcpyDeviceToHost));
    TC_CUDA(cudaMemset(s->pcg_prof, 0, sizeof(unsigned long long) * 16));
    return 0;
}

extern "C" int tc_pcg_stats(tc_handle s, int64_t* stats4) {
    int isc[4];
    TC_CUDA(cudaMemcpyAsync(isc, s->pcg_isc, sizeof isc, cudaMemcpyDeviceToHost, s->stream));
    TC_CUDA(cudaStreamSynchronize(s->stream));
    stats4[0] = isc[2]; stats4[1] = isc[3]; stats4[2] = isc[1]; stats4[3] = s->coop_grid;
    return 0;
}

// status2: [0] status of the last failed PCG solve (0 none, 1 iteration limit reached above the tolerance,
// 2 breakdown: non-finite or non-positive curvature), [1] number of failed solves; cleared on read
extern "C" int tc_pcg_status(tc_handle s, int64_t* status2) {
    int isc[8];
    TC_CUDA(cudaMemcpyAsync(isc, s->pcg_isc, sizeof isc, cudaMemcpyDeviceToHost, s->stream));
    TC_CUDA(cudaStreamSynchronize(s->stream));
    status2[0] = isc[4]; status2[1] = isc[5];
    TC_CUDA(cudaMemsetAsync(s->pcg_isc + 4, 0, 2 * sizeof(int), s->stream));
    return 0;
}

// ---- stand-alone entry points --------------------------------------------------------------------
// One traced link program evaluated for m argument pairs (parity hook for the interpreter of vm.cuh)
__global__ void tc_vm_eval_kernel(const int* __restrict__ code, int n_instr, int result, const double* __restrict__ consts,
                                  const int* __restrict__ tab_dir, const double* __restrict__ tab_arena,
                                  const double* __restrict__ inputs, const double* __restrict__ a0,
                                  const double* __restrict__ a1, int m, double* __restrict__ out) {
    __shared__ int dir[4];
    if (threadIdx.x == 0) { dir[0] = 0; dir[1] = n_instr; dir[2] = 0; dir[3] = result; }
    __syncthreads();
    extern __shared__ double sreg[];            // [n_instr][blockDim]
    TcTables tb{tab_dir, tab_arena};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
        out[i] = tc_vm_run(code, consts, dir, 0, tb, inputs, a0[i], a1 ? a1[i] : 0.0, sreg + threadIdx.x, blockDim.x);
}

extern "C" int tc_vm_eval(void* cuda_stream, const int32_t* code, int32_t n_instr, int32_t result, const double* consts,
                          const int32_t* tab_dir, const double* tab_arena, const double* inputs,
                          const double* a0, const double* a1, int32_t m, double* out) {
    TC_CHECK(n_instr > 0 && n_instr <= TC_VM_MAX && result >= 0 && result < n_instr, "program length out of range");
    if (m <= 0) return 0;
    const size_t smem = sizeof(double) * 64 * (size_t)n_instr;
    static unsigned long long configured = 0ULL;        // per device (a process may drive several GPUs)
    int devid = 0;
    TC_CUDA(cudaGetDevice(&devid));
    if (devid >= 64 || !(configured >> devid & 1ULL)) {
        TC_CUDA(cudaFuncSetAttribute(tc_vm_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 8 * 400));
        if (devid < 64) configured |= 1ULL << devid;
    }
    TC_CHECK(n_instr <= 400, "program too long for the stand-alone evaluator");
    tc_vm_eval_kernel<<<(m + 63) / 64, 64, smem, (cudaStream_t)cuda_stream>>>(code, n_instr, result, consts, tab_dir, tab_arena,
                                                                           inputs, a0, a1, m, out);
    TC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tc_sell_spmv(void* cuda_stream, int32_t mode, int32_t n, int32_t nslices,
                            const int64_t* slice_ptr, const int32_t* col, const int16_t* col16, const double* val,
                            const double* x, double* y, const double* b, const double* mass, double shift,
                            int32_t grid_blocks) {
    cudaStream_t st = (cudaStream_t)cuda_stream;
    TcSell A{n, nslices, (const long long*)slice_ptr, col, val, (const short*)col16};
    int g = grid_blocks > 0 ? grid_blocks : grid_for(nslices, TC_SPMV_THREADS / 32, 148 * 8);
    static double* one_ctrl = nullptr;      // ctrl[1] = 1 so that shift == coef
    static double* scratch = nullptr;
    if (!one_ctrl) {
        TC_CUDA(cudaMalloc(&one_ctrl, sizeof(double) * 4));
        const double c[4] = {0.0, 1.0, 0.0, 0.0};
        TC_CUDA(cudaMemcpy(one_ctrl, c, sizeof c, cudaMemcpyHostToDevice));
        TC_CUDA(cudaMalloc(&scratch, sizeof(double) * 4 * TC_MAX_PARTIALS));
    }
    if (g > TC_MAX_PARTIALS) g = TC_MAX_PARTIALS;
    if (mode == SPMV_NEG)
        tc_sell_spmv_kernel<SPMV_NEG><<<g, TC_SPMV_THREADS, 0, st>>>(A, x, y, nullptr, nullptr, one_ctrl, 0.0, nullptr, nullptr);
    else if (mode == SPMV_B)
        tc_sell_spmv_kernel<SPMV_B><<<g, TC_SPMV_THREADS, 0, st>>>(A, x, y, b, nullptr, one_ctrl, 0.0, nullptr, nullptr);
    else
        tc_sell_spmv_kernel<SPMV_SHIFT><<<g, TC_SPMV_THREADS, 0, st>>>(A, x, y, nullptr, mass, one_ctrl, shift, scratch, nullptr);
    TC_CUDA(cudaGetLastError());
    return 0;
}

static int pcg_solve_impl(void* cuda_stream, int32_t mode, int32_t n, int32_t nslices,
                          const int64_t* slice_ptr, const int32_t* col, const double* val,
                          const double* mass, const double* adiag, const double* rhs, double* x,
                          double shift, double rtol, int32_t maxit, int32_t* iters, const float* binv);

extern "C" int tc_pcg_solve(void* cuda_stream, int32_t mode, int32_t n, int32_t nslices,
                            const int64_t* slice_ptr, const int32_t* col, const double* val,
                            const double* mass, const double* adiag, const double* rhs, double* x,
                            double shift, double rtol, int32_t maxit, int32_t* iters) {
    return pcg_solve_impl(cuda_stream, mode, n, nslices, slice_ptr, col, val, mass, adiag, rhs, x, shift, rtol, maxit,
                          iters, nullptr);
}
// same with the block-Jacobi preconditioner (binv: nslices x 32 x 32 floats, see tc_set_block_precond)
extern "C" int tc_pcg_solve_bj(void* cuda_stream, int32_t n, int32_t nslices, const int64_t* slice_ptr,
                               const int32_t* col, const double* val, const double* mass, const double* adiag,
                               const double* rhs, double* x, double shift, double rtol, int32_t maxit,
                               int32_t* iters, const float* binv) {
    TC_CHECK(binv, "tc_pcg_solve_bj needs the inverted blocks");
    return pcg_solve_impl(cuda_stream, TC_PCG_COOP_BJ, n, nslices, slice_ptr, col, val, mass, adiag, rhs, x, shift, rtol,
                          maxit, iters, binv);
}

static int pcg_solve_impl(void* cuda_stream, int32_t mode, int32_t n, int32_t nslices,
                          const int64_t* slice_ptr, const int32_t* col, const double* val,
                          const double* mass, const double* adiag, const double* rhs, double* x,
                          double shift, double rtol, int32_t maxit, int32_t* iters, const float* binv) {
    tc_solver tmp;
    tmp.stream = (cudaStream_t)cuda_stream;
    TcPcg P{};
    P.A = TcSell{n, nslices, (const long long*)slice_ptr, col, val, nullptr};
    // one scratch allocation: 4 vectors | partial banks + scalars + ctrl | barrier banks + totals | ints
    const size_t n_aux = 4 * TC_MAX_PARTIALS + 20, n_bar = TC_BAR_PART_DOUBLES + 2 * TC_BAR_MAXK;
    const size_t n_dbl = 4 * (size_t)n + n_aux + n_bar;
    char* blob = nullptr;
    TC_CUDA(cudaMalloc(&blob, sizeof(double) * n_dbl + sizeof(int) * (16 + TC_BAR_CTR_WORDS)));
    TC_CUDA(cudaMemsetAsync(blob + sizeof(double) * 4 * (size_t)n, 0,
                            sizeof(double) * (n_aux + n_bar) + sizeof(int) * (16 + TC_BAR_CTR_WORDS), tmp.stream));
    double* ws = (double*)blob;
    double* aux = ws + 4 * (size_t)n;
    P.bar.part = aux + n_aux; P.bar.res = P.bar.part + TC_BAR_PART_DOUBLES;
    int* isc = (int*)(blob + sizeof(double) * n_dbl);
    P.bar.ctr = (unsigned int*)(isc + 16);
    const double c[4] = {0.0, 1.0, 0.0, 0.0};
    double* ctrl = aux + 4 * TC_MAX_PARTIALS + 16;
    TC_CUDA(cudaMemcpyAsync(ctrl, c, sizeof c, cudaMemcpyHostToDevice, tmp.stream));
    TC_CUDA(cudaStreamSynchronize(tmp.stream));
    P.mass = mass; P.adiag = adiag; P.rhs = rhs; P.x = x;
    P.r = ws; P.p = ws + n; P.q = ws + 2 * (size_t)n; P.dinv = ws + 3 * (size_t)n;
    P.partial = aux; P.sc = aux + 4 * TC_MAX_PARTIALS; P.isc = isc;
    P.rtol = rtol; P.maxit = maxit; P.ctrl = ctrl; P.coef = shift; P.prof = nullptr;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int cgp = pcg_coop_grid(sms);
    int rc = launch_pcg(&tmp, P, mode, cgp, nullptr, binv);
    if (rc == 0) {
        cudaError_t e = cudaStreamSynchronize(tmp.stream);
        if (e != cudaSuccess) { tc_set_error(__FILE__, __LINE__, cudaGetErrorString(e)); rc = -1; }
    }
    int h_isc[4] = {0, 0, 0, 0};
    if (rc == 0) cudaMemcpy(h_isc, isc, sizeof h_isc, cudaMemcpyDeviceToHost);
    if (iters) *iters = h_isc[1];
    cudaFree(blob);
    tmp.pin_flags = nullptr;
    return rc;
}
