/* thermca_b200 — C ABI of the B200-native transient solve.
 *
 * Drop-in boundary for ONE path of steffenschroe/Thermca: the time integration inside
 * thermca/solver.py (`transient_solution`, solver.py:61-345; only caller
 * thermca/network.py:978-980).  The reference is pure Python and has no FFI of its own;
 * these are the entry points a ctypes binding on the reference side loads (see
 * INTEGRATION.md).  Every pointer marked `dev` is a DEVICE pointer owned by the caller
 * (torch tensors on the Python side); the library owns only scratch it allocates
 * itself.  All functions return 0 on success, <0 on error (`tc_last_error()`).
 * No function falls back to the CPU.
 */
#ifndef THERMCA_B200_H
#define THERMCA_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct tc_solver* tc_handle;

const char* tc_last_error(void);
int tc_version(void);

/* ---- network level (replaces Network._set_time_dependent_properties, network.py:747-868,
 *      solve_stationary_nodes / cond_to_differential_equation_representation and the
 *      network derivative, solver.py:30-58,208-255) ------------------------------------ */
typedef struct {
    const int32_t* iarena;      /* dev: all integer tables (index maps, programs, commands) */
    double* darena;             /* dev: all fp64 tables and the mutable network state */
    int32_t cmds_off;           /* iarena offset of the command stream (network.cuh) */
    int32_t prog_dir_off;       /* iarena offset: 4 ints per link program */
    int32_t tab_dir_off;        /* iarena offset: 2 ints per interpolation table */
    int32_t N, u, nnz;          /* nodes, unknown nodes, nnz of run_cond */
    int32_t o_T, o_heat, o_capy, o_vol, o_volcapy, o_condy, o_cond, o_clean, o_calc, o_inputs;
    int32_t o_indptr, o_indices, o_diag;
    /* surface-mean jobs (C_surf^T x, C_marg^T x, MOR C_surfs), solver.py:211-220 */
    int32_t n_mean_jobs, n_mean_blocks;
    const int32_t* job_of_block;    /* dev */
    const int32_t* first_block;     /* dev */
    const int32_t* nblk;            /* dev */
    const int64_t* ent_start;       /* dev */
    const int32_t* ent_idx;         /* dev: index into the stacked state */
    const double* ent_w;            /* dev */
    int32_t n_iarena, n_darena;     /* arena lengths in words (the network kernel warms its L1 with them) */
} tc_net_desc;

/* ---- full-order FE part (replaces FFEIO, thermca/fem/ffe_io.py:8-83 at run time) ------- */
typedef struct {
    int32_t n, nslices;
    int64_t y_off;                  /* offset of the part in the stacked state */
    const int64_t* slice_ptr;       /* dev: SELL-32 */
    const int32_t* col;             /* dev */
    const int16_t* col16;           /* dev, optional: col - row when the band fits 16 bits (else NULL) */
    double* val;                    /* dev: mutable (film coefficients live on the diagonal) */
    const double* mass;             /* dev: M_diag = 1 / M_inv */
    double* adiag;                  /* dev: dense copy of diag(A) */
    /* surface loads  b = B[:, nz] flux[nz]  (solver.py:279-281) */
    int32_t load_nrows;
    const int32_t* load_row; const int32_t* load_start; const int32_t* load_surf; const double* load_bval;
    /* films on the diagonal (ffe_io.py:76-82); fd_nrows == 0 when films are constant */
    int32_t fd_nrows;
    const int64_t* fd_pos; const int32_t* fd_row; const double* fd_abody;
    const int32_t* fd_start; const int32_t* fd_film; const double* fd_adata;
    int32_t films_doff;             /* darena offset of films[F] */
    int32_t flux_doff;              /* darena offset of surf_flux[S] */
    /* temperature dependent body (fe_system.py:185-193): when tdep != 0 the SELL arrays above hold
     * Lx_base (geometry only), `mass` holds Mx_diag and the product is formed edge by edge */
    int32_t tdep, condy_tab, volcapy_tab;
    int32_t ts_nrows;
    const int32_t* ts_row; const int32_t* ts_start; const int32_t* ts_surf; const double* ts_qval;
    const int32_t* ts_fstart; const int32_t* ts_film; const double* ts_adata;
    /* tdep only: operator frozen at the initial temperatures, A0 = M0^-1 (L0 + film diagonal) in SELL-32 with
     * its mass M0 and diagonal.  It is the Jacobian approximation of the W-method (tc_ros2_*); NULL = the part
     * is stepped explicitly inside the implicit scheme */
    int32_t j_nslices;
    const int64_t* j_slice_ptr; const int32_t* j_col; const double* j_val; const double* j_mass; const double* j_adiag;
    /* fd_rhs != 0: films vary and every right-hand side refreshes the diagonal from the fd_* lists (the reference's
     * per-call copy, ffe_io.py:76-82); 0 with fd_nrows > 0: the lists only serve tc_set_jac_films */
    int32_t fd_rhs;
} tc_ffe_desc;

/* ---- LP part (replaces LPIO at run time, thermca/lpm/lp_io.py:8-80); operator and load
 *      vector live inside the arenas because the network kernel updates them ------------- */
typedef struct {
    int32_t n, nslices;
    int64_t y_off;
    int32_t slice_ptr_ioff64;       /* iarena offset (in int32 words, 8-byte aligned) of int64 slice_ptr */
    int32_t col_ioff;
    int32_t val_doff;
    int32_t b_doff;                 /* dense load vector M_inv*(Qs flux + L_film_margs T_link) */
    int32_t mass_doff;
    int32_t adiag_doff;             /* dense diag(A), kept current by the network kernel */
    int32_t F, lfm_doff, minv_doff; /* L_film_margs[n][F] and M_inv[n] (lp_io.py:83-103), read by tc_set_jac_films */
} tc_lp_desc;

/* ---- MOR part (replaces MORIO at run time, thermca/fem/mor_io.py:78-92) ------------------ */
typedef struct {
    int32_t S, r, F;
    int64_t y_off;
    const double* A_body;           /* dev: S*r*r */
    const double* A_films;          /* dev: F*S*r*r */
    const double* B_T;              /* dev: S*r */
    int32_t films_doff, flux_doff;
} tc_mor_desc;

int tc_create(tc_handle* out, void* cuda_stream);
int tc_destroy(tc_handle h);
int tc_set_network(tc_handle h, const tc_net_desc* d);
int tc_add_ffe_part(tc_handle h, const tc_ffe_desc* d);
int tc_add_lp_part(tc_handle h, const tc_lp_desc* d);
int tc_add_mor_part(tc_handle h, const tc_mor_desc* d);
int tc_finalize(tc_handle h, int64_t n_state);
/* Films between a part surface and a STATIONARY node (thermca/solver.py:30-42; the reference's idiom for a
 * prescribed surface heat flow is StatNode + HeatSource + a stiff film, thermca/tests/test_lp_part.py:655-660).
 * The node follows the surface instantly, so inside the Jacobian of the implicit W-methods (tc_ros2_*, tc_row3_*)
 * such a film is the rank-one term u c^T (u: the film's load vector times theta = g_film / sum of the node's
 * conductances, kept by the network kernel at theta_doff; c: the surface-mean weights = mean job `job`), applied
 * through the Woodbury identity around the block solves.  The right-hand side is not affected.
 * rec7[n][7] = {kind (0 LP, 1 FE, 2 MOR), part, film, surface of the film, mean job, theta_doff, film_doff};
 * n <= 4; call after tc_finalize. */
int tc_set_jac_films(tc_handle h, int32_t n, const int32_t* rec7);

/* Right-hand side f(t, y) = update_time_derivative (solver.py:193-302). Stateful like the
 * reference: node temperatures, heat sources and conductances persist between calls. */
int tc_rhs(tc_handle h, double t, const double* y_dev, double* ydot_dev);

/* ---- explicit adaptive Runge-Kutta == scipy RK45 / RK23 as driven by simple_solve_ivp
 *      (solver.py:348-390), step control kept on the device -------------------------------- */
enum { TC_RK45 = 0, TC_RK23 = 1 };
int tc_set_state(tc_handle h, const double* y_dev);
int tc_get_state(tc_handle h, double* y_dev);
int tc_get_state_async(tc_handle h, double* y_dev);   /* enqueue only (tc_set_state never blocks either) */
/* history ring written on the device after accepted steps (store_results, solver.py:393-424):
 * hist_io[capacity][n_state - u], hist_net[capacity][1 + 3N + 2 nnz] = time | T | capy | heat | cond | clean */
int tc_set_history(tc_handle h, double* hist_io_dev, double* hist_net_dev, int32_t capacity, int32_t num_skip);
/* Probe mode of the history (results at scale): with n_probe > 0 a row of hist_io holds only the listed part
 * DOFs, hist_io[capacity][n_probe] (indices into the part block of the state, i.e. the io_temp layout of
 * solver.py:315-316), instead of the full part state the reference copies per stored step; n_probe = 0
 * restores full rows.  Call before tc_set_history. */
int tc_set_probes(tc_handle h, const int64_t* probe_idx_dev, int32_t n_probe);
int tc_rk_begin(tc_handle h, int32_t method, double t0, double t_end, double rtol, double atol,
                double max_step, double first_step);
/* Enqueue up to max_attempts step attempts (one CUDA graph replay each); stops early once the
 * device reports done. info[8] (host): done, accepted, rejected, status, stored, nfev, t (bits lo/hi) */
int tc_rk_advance(tc_handle h, int32_t max_attempts, int32_t poll_every);
int tc_get_info(tc_handle h, int64_t* info8, double* t_now, double* h_next);
int tc_reset_history(tc_handle h);
/* Diagnostic: out2[0] = 1 if the explicit attempt graph is instantiated; out2[1] = 2 / 3 if a ROS2 / ROW3w attempt
 * graph is instantiated, -1 if the runtime refused the capture (attempts are then launched kernel by kernel), 0 if none
 * has been captured since the last begin. */
int tc_graph_info(tc_handle h, int32_t* out2);

/* ---- implicit steps: solve (M + c h K) per FE part with Jacobi-PCG (new numerical method;
 *      the reference has no implicit scheme) ------------------------------ */
enum { TC_PCG_COOP = 0, TC_PCG_MULTI = 1, TC_PCG_COOP_BJ = 2 };   /* 2: block-Jacobi (tc_set_block_precond first) */
/* one fixed theta step of size h_step: (I + theta h A) d = h f(t + theta h, y); y += d */
int tc_theta_steps(tc_handle h, int32_t nsteps, double theta, double h_step, double pcg_rtol,
                   int32_t pcg_maxit, int32_t pcg_mode);
/* Adaptive linearly implicit 2-stage Rosenbrock-W method (ROS2, gamma = 1 + 1/sqrt 2) with embedded first-order
 * estimate and the RK step controller; serves the reference's `method='LSODA'` requests (thermca/solver.py:359-365).
 * Block Jacobian: FE / LP parts by Jacobi-PCG (temperature dependent FE bodies with their operator frozen at the
 * initial temperatures), unknown network nodes and MOR blocks by small CG solves; every block is left explicit for
 * an attempt while gamma*h*(Gershgorin bound of its operator) < 2 (env TC_W_EXPLICIT_BOUND overrides).  CPU
 * restatement: oracle/ros2_oracle.py. */
int tc_ros2_begin(tc_handle h, double t0, double t_end, double rtol, double atol, double max_step,
                  double first_step, double pcg_rtol, int32_t pcg_maxit, int32_t pcg_mode);
int tc_ros2_advance(tc_handle h, int32_t max_attempts, int32_t poll_every);
/* adaptive third-order W-method "ROW3w" (4 stages, 3 right-hand sides per attempt; same block structure, history and
 * controller as tc_ros2_*, error exponent -1/3); serves the reference's method='LSODA' requests (thermca/solver.py:359-362,
 * thermca/tests/test_fe_part.py:13-99,230-303) */
int tc_row3_begin(tc_handle h, double t0, double t_end, double rtol, double atol, double max_step,
                  double first_step, double pcg_rtol, int32_t pcg_maxit, int32_t pcg_mode);
int tc_row3_advance(tc_handle h, int32_t max_attempts, int32_t poll_every);
int tc_set_time(tc_handle h, double t);
/* Switch CUDA-event timing of the PCG launches on/off; returns (and clears) the time and launch
 * counts accumulated since the previous call. */
int tc_timing(tc_handle h, int32_t enable, double* pcg_ms, int64_t* pcg_launches, int64_t* launches);
int tc_pcg_profile(tc_handle h, uint64_t* out16);   /* per-phase ns of block 0 while timing is on */
/* PCG statistics of the implicit path: total iterations, solves, last iteration count */
int tc_pcg_stats(tc_handle h, int64_t* stats4);
/* block-Jacobi preconditioner of full-order FE part `part` (pcg_mode TC_PCG_COOP_BJ): inverses of the 32x32 diagonal
 * blocks of M + shift*K that belong to the SELL slices, FP32, [slice][j][i], symmetric; device pointer kept by reference */
int tc_set_block_precond(tc_handle h, int32_t part, const float* binv_dev, double shift);
/* status2[0]: status of the last FAILED PCG solve since the previous call (0 none, 1 iteration limit reached with
 * the residual above the tolerance, 2 breakdown: non-finite / non-positive curvature); status2[1]: number of
 * failed solves.  Cleared on read.  The host drivers raise on it (there is no silent degradation). */
int tc_pcg_status(tc_handle h, int64_t* status2);

/* ---- stand-alone kernels (parity hooks and micro-benchmarks) ----------------------------- */
/* y = -A x (mode 0), y = b - A x (mode 1), y = mass*(x + shift*A x) (mode 2) on a SELL-32 operator */
int tc_sell_spmv(void* cuda_stream, int32_t mode, int32_t n, int32_t nslices, const int64_t* slice_ptr,
                 const int32_t* col, const int16_t* col16, const double* val, const double* x, double* y,
                 const double* b, const double* mass, double shift, int32_t grid_blocks);
/* One traced link program (thermca_b200/trace.py: int32[4] per instruction) evaluated for m argument pairs:
 * what one call of a film / conductance closure `f(T0, T1, matl)` (thermca/network.py:781-785) returns.
 * tab_dir[2t] = offset of table t in tab_arena, tab_dir[2t+1] = its length (x values, then y values). */
int tc_vm_eval(void* cuda_stream, const int32_t* code, int32_t n_instr, int32_t result, const double* consts,
               const int32_t* tab_dir, const double* tab_arena, const double* inputs, const double* a0,
               const double* a1, int32_t m, double* out);
/* Jacobi-PCG on (M + shift K): solves mass*(I + shift*A) x = mass*rhs; returns iterations in *iters (host) */
int tc_pcg_solve(void* cuda_stream, int32_t mode, int32_t n, int32_t nslices, const int64_t* slice_ptr,
                 const int32_t* col, const double* val, const double* mass, const double* adiag,
                 const double* rhs, double* x, double shift, double rtol, int32_t maxit, int32_t* iters);

/* micro-benchmark: ns per grid barrier of an empty persistent kernel; mode 0 = reduce-barrier (csrc/gridbar.cuh)
 * carrying K in {0,1,2} values, 1 = cooperative_groups grid.sync(), 2 = last-arriver variant (K in {1,2}) */
int tc_bar_bench(void* cuda_stream, int32_t blocks, int32_t threads, int32_t iters, int32_t K, int32_t mode,
                 double* ns_per_barrier);

/* tc_pcg_solve with the block-Jacobi preconditioner (binv as in tc_set_block_precond) */
int tc_pcg_solve_bj(void* cuda_stream, int32_t n, int32_t nslices, const int64_t* slice_ptr, const int32_t* col,
                    const double* val, const double* mass, const double* adiag, const double* rhs, double* x,
                    double shift, double rtol, int32_t maxit, int32_t* iters, const float* binv);

/* ---- row-partitioned implicit step across GPUs (BASELINE config #5), csrc/dist.cu ------------
 * One process per GPU; the comm buffer of every rank is exported as a CUDA IPC handle and opened
 * by its peers; ghost values and the PCG scalar reductions then move by peer stores inside ONE
 * cooperative kernel per theta step (no NCCL call on the data path).  The reference has no
 * multi-process code; per rank this is thermca/solver.py:265-281 (right-hand side
 * of a full-order FE part) plus the implicit solve on the local rows.
 * Ghost values and reduction mailboxes are "LL" encoded in the exported buffer (every 8-byte granule carries a
 * sequence tag): no flags, no system fences.
 * Calls return 0, or a negative code with tc_last_error(): -3 = a wait on a peer timed out
 * (TC_DIST_TIMEOUT_S seconds, default 20), -4 = PCG iteration limit / breakdown. */
typedef struct tc_dist tc_dist;
int tc_dist_create(tc_dist** out, void* cuda_stream, int32_t rank, int32_t world, int64_t n_loc, int64_t n_ghost);
int tc_dist_ipc_handle(tc_dist* d, void* handle64);
/* handles: world x 64 bytes; peer_ghost[w] = ghost count of rank w */
int tc_dist_open_peers(tc_dist* d, const void* handles, const int64_t* peer_ghost);
/* test harness: the "ranks" are sibling handles of ONE process on one GPU (own streams); all[w] = handle of rank w */
int tc_dist_open_peers_local(tc_dist* d, tc_dist* const* all, const int64_t* peer_ghost);
/* halo plan (host arrays, copied): entry e sends local row send_row[e] to ghost slot send_dst[e] of rank
 * send_peer[e]; nbr = ranks this rank trades ghosts with */
int tc_dist_set_halo(tc_dist* d, int64_t n_send, const int32_t* send_row, const int32_t* send_peer,
                     const int64_t* send_dst, int32_t n_nbr, const int32_t* nbr);
/* SELL-32 operator of the local rows over LOCAL columns (device pointers, kept by reference; col16 optional =
 * column - row); bidx[s] (required when n_bnd > 0) = index of slice s among the n_bnd boundary slices or -1; couplings
 * to ghosts: side lists of the boundary slices, entry k of lane l of boundary slice b at gptr[b] + 32 k + l =
 * (ghost slot, operator value), padded with (0, 0.0) */
int tc_dist_set_operator(tc_dist* d, int32_t nslices, const int64_t* slice_ptr, const int32_t* col,
                         const int16_t* col16, const double* val, const double* mass, const double* adiag,
                         const double* b, int32_t n_bnd, const int32_t* bidx, const int64_t* gptr,
                         const int32_t* gslot, const double* gval);
int tc_dist_set_load(tc_dist* d, const double* b_local_dev);     /* swap the load vector (time-varying loads) */
int tc_dist_set_state_ptr(tc_dist* d, double* y_local_dev);      /* work on a caller-owned state block (NULL: own) */
/* surface means of a partitioned body: sum the per-rank partial sums of jobs [job0, job0 + njobs) over all ranks in
 * rank order (thermca/solver.py:211-220); partials / first_block / nblk: the job tables of tc_set_network */
int tc_dist_mean_exchange(tc_dist* d, double* partials_dev, const int32_t* first_block_dev, const int32_t* nblk_dev,
                          int32_t job0, int32_t njobs);
/* explicit stages of a partitioned body: out = out - A v on the local rows (out holds the load vector b on entry), the
 * ghost rows of v exchanged by peer stores inside the kernel; done_flag_dev (may be NULL): skip when *flag != 0 */
int tc_dist_residual(tc_dist* d, const double* v_local_dev, double* out_local_dev, const int32_t* done_flag_dev);
/* stage solve of the W-methods on the partitioned body: z = (I + coef * h A)^-1 rhs on the local rows (the fused
 * Jacobi-PCG kernel without product / commit), h read from *h_dev; *skip_flag_dev != 0 (may be NULL): leave z as it is */
int tc_dist_solve(tc_dist* d, double coef, const double* h_dev, const double* rhs_local_dev, double* z_local_dev,
                  double rtol, int32_t maxit, const int32_t* skip_flag_dev, const int32_t* done_flag_dev);
/* vals[j] <- sum over the ranks (rank order, bit-identical everywhere) for k <= 64 doubles on the device */
int tc_dist_sum_exchange(tc_dist* d, double* vals_dev, int32_t k, const int32_t* done_flag_dev);
int tc_dist_max_exchange(tc_dist* d, double* vals_dev, int32_t k, const int32_t* done_flag_dev);   /* same with max */
int32_t tc_dist_rank(tc_dist* d);
int64_t tc_dist_local_rows(tc_dist* d);
void* tc_dist_stream(tc_dist* d);
/* general networks on several GPUs: FE part `part` of solver h holds the LOCAL rows of a row-partitioned body; tc_rhs and
 * tc_theta_steps then add the peers' rows to its surface means and run its product / solve in the fused kernel */
int tc_attach_dist(tc_handle h, int32_t part, tc_dist* d, int32_t first_job, int32_t n_jobs);
/* rows of the partitioned body over ALL ranks: lets tc_rk_begin / tc_rk_advance run the adaptive explicit pairs on the
 * partitioned network (error and initial-step norms are means over the global state, thermca/solver.py:348-390) */
int tc_set_dist_global_rows(tc_handle h, int64_t n_rows_global);
int tc_dist_set_state(tc_dist* d, const double* y_local_dev);
int tc_dist_get_state(tc_dist* d, double* y_local_dev);          /* fails (-3/-4) if a step since the last reset failed */
int tc_dist_copy_state_async(tc_dist* d, double* y_local_dev, int32_t dir);   /* enqueue only: 0 set, 1 get */
int tc_dist_theta_steps(tc_dist* d, int32_t nsteps, double theta, double h, double rtol, int32_t maxit);
/* PCG formulation of tc_dist_theta_steps (the same on every rank): 0 = classic (two reductions and a barrier per
 * iteration), 1 = pipelined (Ghysels-Vanroose recurrences fused into the product sweep: ONE sweep and ONE reduction per
 * iteration; pays under strong scaling, few rows per GPU).  -2 when the body cannot be served by the pipelined kernel. */
int tc_dist_set_pcg(tc_dist* d, int32_t mode);
int32_t tc_dist_get_pcg(tc_dist* d);
int tc_dist_check(tc_dist* d);                                   /* synchronises; 0, -3 or -4 */
int tc_dist_reset_error(tc_dist* d);
/* iters_last, iters_total, comm error flag, solves + 1e6 * grid, failed solves, ghosts that were waited for
 * (counted only with TC_DIST_DIAG=1) */
int tc_dist_stats(tc_dist* d, int64_t* stats6);
int tc_dist_profile(tc_dist* d, uint64_t* out32);   /* per-phase wall time (ns) of block 0 / last block */
/* diagnostic (TC_DIST_DIAG=3 at set_operator): out[block * 10 + phase] accumulated ns of EVERY block (slot 8: SM id); returns the grid */
int tc_dist_block_profile(tc_dist* d, uint64_t* out, int32_t max_blocks);
/* teardown: tc_dist_close_peers on every rank, a host barrier, then tc_dist_destroy (CUDA IPC rule: the
 * exporter must not free while an importer has the buffer mapped) */
int tc_dist_close_peers(tc_dist* d);
int tc_dist_destroy(tc_dist* d);

/* ---- batched MOR ensemble (BASELINE config #3): B scenarios in lockstep, FP64 DMMA ------- */
/* One classical RK4 step of size h for X[S][r][B] with per-scenario films film[F][B](t) =
 * h0*(1 + 0.5 sin(2 pi t / tau)); operators shared by all scenarios. */
int tc_mor_batch_steps(void* cuda_stream, int32_t S, int32_t r, int32_t F, int32_t B,
                       const double* A_body, const double* A_films, const double* B_T,
                       const int32_t* film_surf, const double* film_h0, const double* film_tau,
                       const double* t_link, const double* q_surf, double* X, double* work,
                       double t0, double h, int32_t nsteps, int32_t use_dmma);
int64_t tc_mor_batch_work_doubles(int32_t S, int32_t r, int32_t F, int32_t B);
/* Fused persistent form of the same step (csrc/mor_fused.cu): one launch integrates nsteps RK4 steps; a CTA keeps a
 * (surface, 32 scenario columns) unit in shared memory over all stages and streams the operators, pre-packed once
 * into DMMA fragment order by tc_mor_fused_pack, through a bulk-copy ring.  Same reference arithmetic
 * (thermca/solver.py:282-299, thermca/fem/mor_io.py:90-92).  Supported: F <= 2, r <= 200 (tc_mor_fused_supported);
 * work = tc_mor_fused_work_bytes bytes of device memory (hand-over flags). */
int32_t tc_mor_fused_supported(int32_t S, int32_t r, int32_t F, int32_t B);
int64_t tc_mor_fused_packed_doubles(int32_t S, int32_t r, int32_t F);
int64_t tc_mor_fused_work_bytes(int32_t S, int32_t r, int32_t F, int32_t B);
int tc_mor_fused_pack(void* cuda_stream, int32_t S, int32_t r, int32_t F, const double* A_body,
                      const double* A_films, double* packed);
int tc_mor_fused_steps(void* cuda_stream, int32_t S, int32_t r, int32_t F, int32_t B, const double* packed,
                       const double* B_T, const int32_t* film_surf, const double* film_h0,
                       const double* film_tau, const double* t_link, const double* q_surf, double* X,
                       void* work, double t0, double h, int32_t nsteps);

/* ---- device-side generator of the synthetic Kuhn cuboid operator (BASELINE configs #2/#5) */
int tc_synth_cuboid_counts(int32_t nx, int32_t ny, int32_t nz, int64_t* n_rows, int64_t* n_slices,
                           int64_t* sell_entries);
int tc_synth_cuboid(void* cuda_stream, int32_t nx, int32_t ny, int32_t nz, double lx, double ly, double lz,
                    double jitter, double condy, double vol_capy, int64_t row_begin, int64_t row_end,
                    int64_t* slice_ptr, int32_t* col, int16_t* col16, double* val, double* mass, double* adiag,
                    double* q_bottom, double* q_upper);

/* ---- P1 operator build on the device (csrc/assembly.cu) --------------------------------------
 * Replaces the scikit-fem assembly behind FESystem.fem_matrices_from_skfem (thermca/fem/fe_system.py:322-400)
 * for linear tetrahedra: per element (key, value) contributions, an in-order segment sum after the caller has
 * sorted the keys, and the SELL-32 layout of A = scale * row_scale * L the solver kernels read. */
int tc_p1_tet_contributions(void* cuda_stream, int64_t n_points, int64_t n_tets, const double* points,
                            const int32_t* tets, int64_t* keys, double* vals, int32_t* mkeys, double* mvals);
int tc_p1_tri_loads(void* cuda_stream, int64_t n_tris, const double* points, const int32_t* tris, int32_t* keys,
                    double* vals);
int tc_segment_sum(void* cuda_stream, int64_t nseg, const int64_t* start, const double* vals, double* out);
int tc_csr_to_sell(void* cuda_stream, int32_t n, int32_t nslices, const int64_t* indptr, const int32_t* indices,
                   const double* data, const double* row_scale, double scale, const int64_t* slice_ptr,
                   int32_t* col, int16_t* col16, double* val, double* diag_out);

/* ---- Jacobi-PCG for K <= 8 right-hand sides sharing one matrix sweep per iteration (csrc/pcg_multi.cu) ----
 * A X = B, A symmetric positive definite in SELL-32 with diagonal `diag`; B, X: [n][K] with the K values of a
 * row contiguous; iters_dev[K]: device array of iteration counts (maxit = not converged).
 * work: tc_pcg_multi_work_doubles(n, K) doubles.  Replaces the per-surface solves of
 * arnoldi_iteration (thermca/fem/mor_io.py:206) when several surfaces advance together. */
int64_t tc_pcg_multi_work_doubles(int64_t n, int32_t K);
int tc_pcg_solve_multi(void* cuda_stream, int32_t n, int32_t nslices, const int64_t* slice_ptr, const int32_t* col,
                       const double* val, const double* diag, int32_t K, const double* B, double* X, double* work,
                       double rtol, int32_t maxit, int32_t* iters_dev);

/* ---- orthogonalisation kernels of the inverse-Krylov MOR basis build (csrc/krylov.cu) ------------
 * Replace the modified Gram-Schmidt loop of arnoldi_iteration (thermca/fem/mor_io.py:203-217): v is
 * orthogonalised against the first k columns of the column-major Q (column j at Q + j*ld) by `passes`
 * rounds of classical Gram-Schmidt; |v|^2 afterwards goes to nrm2_out[0] (device memory), the summed
 * projection coefficients to h_out[0..k) if not NULL.  work: tc_krylov_work_doubles(k_max) doubles. */
int64_t tc_krylov_work_doubles(int32_t k_max);
int tc_krylov_orthogonalize(void* cuda_stream, int64_t n, int32_t k, const double* Q, int64_t ld, double* v,
                            int32_t passes, double* work, double* h_out, double* nrm2_out);
/* out = v / sqrt(nrm2[0]) */
int tc_krylov_normalize(void* cuda_stream, int64_t n, const double* v, const double* nrm2, double* out);

#ifdef __cplusplus
}
#endif
#endif
