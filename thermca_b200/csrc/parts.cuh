// Part-level kernels: the FE / LP state-space operators and their surface coupling.
//
// Operator layout in HBM: SELL-32 ("sliced ELLPACK", slice height = one warp).  Slice s
// holds rows 32s..32s+31 column-major: entry k of local row l sits at
// slice_ptr[s] + 32*k + l, so a warp reads 256 B of values and 128 B of column
// indices per k, fully coalesced, and every thread owns one row: the row sum is a
// sequential, un-fused multiply-add chain in the stored entry order — the same order
// scipy's csr_matvec uses on the reference's CSR (thermca/solver.py:264,281 via
// thermca/_utils/sparse.py:10), which makes the product bit-identical to the CPU path.
// Short rows are padded with (value 0, column = own row).
#pragma once
#include "common.cuh"

struct TcSell {
    int n;                        // rows
    int nslices;
    const long long* slice_ptr;   // nslices + 1
    const int* col;
    const double* val;
    // optional compressed column indices: col = row + col16 (valid when the operator band is < 2^15, true
    // for slab-ordered meshes); halves the index traffic (4 -> 2 B per non-zero, ~10 % of the PCG bytes)
    const short* col16;
};

#define TC_SPMV_THREADS 256

enum { SPMV_NEG = 0, SPMV_B = 1, SPMV_SHIFT = 2 };

// MODE SPMV_NEG  : y = -A x                                  (FFE right-hand side, b added on surface rows after)
// MODE SPMV_B    : y = b - A x                               (LP right-hand side)
// MODE SPMV_SHIFT: y = mass * (x + shift * A x), partial[blockIdx] = sum x_i y_i
//                  shift = coef * ctrl[1] (= gamma * h, read on the device)
// Per-thread body (grid-stride over slices); returns this thread's share of sum x_i y_i.
// NC_X: gather x through the read-only (non-coherent) path.  Must be false when x is written
// by other blocks of the SAME launch (persistent PCG kernel): then x is read with ld.global.ca,
// which the gpu-scope fence of grid.sync() keeps coherent.
// NC_X = 1: read-only path; 0: ld.global.ca; 2: ld.global.cg (L2 only: data another GPU may have
// just written through NVLink, no L1 invalidation needed).
template <int NC_X>
__device__ __forceinline__ double tc_ld_x(const double* p) {
    if (NC_X == 1) return __ldg(p);
    double v;
    if (NC_X == 2) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    else asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// One group of N stored entries of a row: N value + N index loads in flight, then N gathers, then the un-fused
// multiply-add chain in stored order.
__device__ __forceinline__ int tc_ld_idx(const short* p, unsigned long long pol, bool ef) {
    return ef ? ldm_s16<true>(p, pol) : ld_stream_s16(p);
}
__device__ __forceinline__ int tc_ld_idx(const int* p, unsigned long long pol, bool ef) {
    return ef ? ldm_s32<true>(p, pol) : ld_stream_s32(p);
}
template <int N, int NC_X, bool EF, class CT>
__device__ __forceinline__ double tc_row_group(double sum, const double* v, const CT* c, const double* xb,
                                               unsigned long long pol) {
    double vv[N], xx[N];
    int cc[N];
#pragma unroll
    for (int j = 0; j < N; ++j) vv[j] = ldm_f64<EF>(v + 32 * j, pol);
#pragma unroll
    for (int j = 0; j < N; ++j) cc[j] = tc_ld_idx(c + 32 * j, pol, EF);
#pragma unroll
    for (int j = 0; j < N; ++j) xx[j] = tc_ld_x<NC_X>(xb + cc[j]);
#pragma unroll
    for (int j = 0; j < N; ++j) sum = __dadd_rn(sum, __dmul_rn(vv[j], xx[j]));
    return sum;
}
// Last full group (its indices `cc` are already here) and the R-entry tail of a row: the tail indices travel with the
// group's values, the tail values with the group's gathers.
template <int R, int NC_X, bool EF, class CT>
__device__ __forceinline__ double tc_row_finish(double sum, const double* v, const CT* c, const double* xb, const int* cc,
                                                unsigned long long pol) {
    double vv[4], xx[4];
    int tc[R > 0 ? R : 1];
#pragma unroll
    for (int j = 0; j < 4; ++j) vv[j] = ldm_f64<EF>(v + 32 * j, pol);
#pragma unroll
    for (int j = 0; j < R; ++j) tc[j] = tc_ld_idx(c + 32 * (4 + j), pol, EF);
#pragma unroll
    for (int j = 0; j < 4; ++j) xx[j] = tc_ld_x<NC_X>(xb + cc[j]);
#pragma unroll
    for (int j = 0; j < 4; ++j) sum = __dadd_rn(sum, __dmul_rn(vv[j], xx[j]));
    if (R > 0) {
        double tv[R > 0 ? R : 1], tx[R > 0 ? R : 1];
#pragma unroll
        for (int j = 0; j < R; ++j) tv[j] = ldm_f64<EF>(v + 32 * (4 + j), pol);
#pragma unroll
        for (int j = 0; j < R; ++j) tx[j] = tc_ld_x<NC_X>(xb + tc[j]);
#pragma unroll
        for (int j = 0; j < R; ++j) sum = __dadd_rn(sum, __dmul_rn(tv[j], tx[j]));
    }
    return sum;
}
// Row sum of one slice (one row per lane), entries in stored order (bit-identical to the sequential CSR product).
// The warps are LATENCY bound (measured: an SM that runs one block finishes its slices no sooner than an SM that runs
// two; 48 warps per SM with 40 registers are slower than 64 with 32), so what counts is the number of dependent
// memory round trips per slice.  A group of four entries needs values, indices -> gathers -> sum: two round trips; here
// the indices run ONE GROUP AHEAD (4 more registers), so a group's values, its gathers and the next group's indices
// are in flight together: one round trip per group.  All lanes of a slice share its width, so the tail (width mod 4)
// is handled behind warp-uniform branches, its indices prefetched with the last full group (a per-entry tail loop cost
// two round trips per entry).  15-wide rows of a tetrahedral mesh: 13 -> 6 round trips per slice.
template <int NC_X, bool EF, class CT>
__device__ __forceinline__ double tc_row_sum_cols(const double* v, const CT* c, const double* xb, int width,
                                                  unsigned long long pol) {
    double sum = 0.0;
    if (width >= 4) {
        int cc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) cc[j] = tc_ld_idx(c + 32 * j, pol, EF);
        int k = 0;
        for (; k + 8 <= width; k += 4) {                               // another full group follows
            double vv[4], xx[4];
            int nc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) vv[j] = ldm_f64<EF>(v + 32 * (k + j), pol);
#pragma unroll
            for (int j = 0; j < 4; ++j) nc[j] = tc_ld_idx(c + 32 * (k + 4 + j), pol, EF);
#pragma unroll
            for (int j = 0; j < 4; ++j) xx[j] = tc_ld_x<NC_X>(xb + cc[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) sum = __dadd_rn(sum, __dmul_rn(vv[j], xx[j]));
#pragma unroll
            for (int j = 0; j < 4; ++j) cc[j] = nc[j];
        }
        const int rem = width - k - 4;                                 // 0 .. 3, warp uniform
        if (rem == 3) sum = tc_row_finish<3, NC_X, EF>(sum, v + 32 * k, c + 32 * k, xb, cc, pol);
        else if (rem == 2) sum = tc_row_finish<2, NC_X, EF>(sum, v + 32 * k, c + 32 * k, xb, cc, pol);
        else if (rem == 1) sum = tc_row_finish<1, NC_X, EF>(sum, v + 32 * k, c + 32 * k, xb, cc, pol);
        else sum = tc_row_finish<0, NC_X, EF>(sum, v + 32 * k, c + 32 * k, xb, cc, pol);
    } else if (width == 3) sum = tc_row_group<3, NC_X, EF>(sum, v, c, xb, pol);
    else if (width == 2) sum = tc_row_group<2, NC_X, EF>(sum, v, c, xb, pol);
    else if (width == 1) sum = tc_row_group<1, NC_X, EF>(sum, v, c, xb, pol);
    return sum;
}
// xrow_off: index of local row 0 in the gathered vector (row partition: x is addressed by
// GLOBAL column, outputs by LOCAL row); 0 on a single GPU.
template <int NC_X, bool EF>
__device__ __forceinline__ double tc_slice_row_sum(const TcSell& A, const double* x, long long xrow_off, int s, int lane,
                                                   unsigned long long pol = 0ULL) {
    const long long base = A.slice_ptr[s];
    const int width = (int)((A.slice_ptr[s + 1] - base) >> 5);
    const double* v = A.val + base + lane;
    if (A.col16)                                                      // x at this thread's own row + relative column
        return tc_row_sum_cols<NC_X, EF>(v, A.col16 + base + lane, x + (xrow_off + s * 32 + lane), width, pol);
    return tc_row_sum_cols<NC_X, EF>(v, A.col + base + lane, x, width, pol);
}

template <int MODE, int NC_X = 1>
__device__ __forceinline__ double tc_sell_spmv_body(const TcSell& A, const double* x,
                                                    double* y, const double* __restrict__ b,
                                                    const double* __restrict__ mass, double shift,
                                                    long long xrow_off = 0, int s_begin = 0, int s_end = -1) {
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int warp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_block;
    double dot = 0.0;
    if (s_end < 0) s_end = A.nslices;
    for (int s = s_begin + warp; s < s_end; s += nwarps) {
        const double sum = tc_slice_row_sum<NC_X, false>(A, x, xrow_off, s, lane);
        const int row = s * 32 + lane;
        if (row < A.n) {
            if (MODE == SPMV_NEG) {
                y[row] = -sum;
            } else if (MODE == SPMV_B) {
                y[row] = __dsub_rn(b[row], sum);
            } else {
                const double xr = tc_ld_x<NC_X>(x + xrow_off + row);
                const double q = mass[row] * (xr + shift * sum);
                y[row] = q;
                dot += xr * q;
            }
        }
    }
    return dot;
}

template <int MODE>
__global__ void __launch_bounds__(TC_SPMV_THREADS)
tc_sell_spmv_kernel(TcSell A, const double* __restrict__ x, double* __restrict__ y,
                    const double* __restrict__ b, const double* __restrict__ mass,
                    const double* __restrict__ ctrl, double coef, double* __restrict__ partial,
                    const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    __shared__ double sh[32];
    double shift = 0.0;
    if (MODE == SPMV_SHIFT) shift = coef * ctrl[1];
    const double dot = tc_sell_spmv_body<MODE>(A, x, y, b, mass, shift);
    if (MODE == SPMV_SHIFT) {
        const double bs = block_sum(dot, sh);
        if (threadIdx.x == 0) partial[blockIdx.x] = bs;
    }
}

// ---- surface coupling of a full-order FE part ----------------------------------------
// Unique surface rows with their (surface, B value) lists: row r gets
//   ydot[r] += sum_{s in list, flux_s != 0} B[r,s] * flux_s        thermca/solver.py:279-281
struct TcSurfLoad {
    int nrows;
    const int* row;        // unique dof index
    const int* start;      // nrows + 1, into (surf, bval)
    const int* surf;
    const double* bval;
};
static __global__ void tc_ffe_surf_load_kernel(TcSurfLoad L, const double* __restrict__ flux,
                                        double* __restrict__ ydot, const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.nrows; i += gridDim.x * blockDim.x) {
        double bsum = 0.0;
        for (int k = L.start[i]; k < L.start[i + 1]; ++k) {
            const double f = flux[L.surf[k]];
            if (f != 0.0) bsum = __dadd_rn(bsum, __dmul_rn(L.bval[k], f));
        }
        const int r = L.row[i];
        ydot[r] = __dadd_rn(bsum, ydot[r]);
    }
}

// Film coefficients folded into the operator diagonal (thermca/fem/ffe_io.py:76-82):
//   A_ii = A_body_ii + sum_f film_f * a_f,i   in film order, written at the SELL position.
struct TcFilmDiag {
    int nrows;
    const long long* pos;  // position of A_ii in the SELL value array
    const int* row;        // dof index (for the dense diagonal copy)
    const double* abody;   // A_body_ii
    const int* start;      // nrows + 1
    const int* film;       // film index
    const double* adata;   // a_f,i
};
static __global__ void tc_ffe_film_diag_kernel(TcFilmDiag F, const double* __restrict__ films,
                                        double* __restrict__ val, double* __restrict__ adiag,
                                        const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < F.nrows; i += gridDim.x * blockDim.x) {
        double v = F.abody[i];
        for (int k = F.start[i]; k < F.start[i + 1]; ++k)
            v = __dadd_rn(v, __dmul_rn(films[F.film[k]], F.adata[k]));
        val[F.pos[i]] = v;
        if (adiag) adiag[F.row[i]] = v;
    }
}

// ---- temperature dependent FE bodies (FESystem.update_material_properties,
//      thermca/fem/fe_system.py:185-193 + FFEIO.update_material_properties ffe_io.py:67-74) ----
// The reference rebuilds the operator on every right-hand side: L_ij = k((T_i+T_j)/2) Lx_ij
// (diagonal = -row sum), M_i = c(T_i) Mx_i, A = M^-1 L: three full matrix passes.  Here the
// same numbers are formed edge by edge inside the product, nothing is rebuilt or stored:
//   ydot_i = -(1/M_i) * [ sum_{j!=i} k_ij Lx_ij T_j - (sum_{j!=i} k_ij Lx_ij) T_i ]
// Lx_base is stored SELL-32 with its stored diagonal entry (value 0) skipped.
__device__ __forceinline__ double tc_tab(const int* __restrict__ dir, const double* __restrict__ arena, int t, double x) {
    const int off = dir[2 * t], n = dir[2 * t + 1];
    const double* xp = arena + off;
    const double* fp = xp + n;
    if (isnan(x)) return x;
    if (n == 1) return fp[0];
    if (x <= xp[0]) return fp[0];
    if (x >= xp[n - 1]) return fp[n - 1];
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= xp[mid]) lo = mid; else hi = mid;
    }
    const double slope = __ddiv_rn(__dsub_rn(fp[lo + 1], fp[lo]), __dsub_rn(xp[lo + 1], xp[lo]));
    return __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xp[lo])), fp[lo]);
}

struct TcTdep {
    TcSell Lx;                 // Lx_base (geometry only)
    const double* Mx;          // lumped geometric mass
    const int* tab_dir;        // table directory (iarena)
    const double* arena;       // darena
    int condy_tab, volcapy_tab;
};
static __global__ void __launch_bounds__(TC_SPMV_THREADS)
tc_ffe_tdep_spmv_kernel(TcTdep P, const double* __restrict__ x, double* __restrict__ y,
                        const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int warp = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warps_per_block;
    for (int s = warp; s < P.Lx.nslices; s += nwarps) {
        const long long base = P.Lx.slice_ptr[s];
        const int width = (int)((P.Lx.slice_ptr[s + 1] - base) >> 5);
        const int row = s * 32 + lane;
        const bool live = row < P.Lx.n;
        const double xi = live ? x[row] : 0.0;
        const double minv = live ? __ddiv_rn(1.0, __dmul_rn(tc_tab(P.tab_dir, P.arena, P.volcapy_tab, xi), P.Mx[row])) : 0.0;
        double acc = 0.0, offsum = 0.0;
        for (int k = 0; k < width; ++k) {
            const double lx = ld_stream_f64(P.Lx.val + base + 32 * k + lane);
            const int c = ld_stream_s32(P.Lx.col + base + 32 * k + lane);
            if (c == row || !live) continue;
            const double xc = __ldg(x + c);
            const double kij = __dmul_rn(tc_tab(P.tab_dir, P.arena, P.condy_tab, __ddiv_rn(__dadd_rn(xi, xc), 2.0)), lx);
            offsum = __dadd_rn(offsum, kij);
            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(minv, kij), xc));
        }
        if (live) {
            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(minv, -offsum), xi));
            y[row] = -acc;
        }
    }
}
// surface rows of a temperature dependent part: load with the current 1/M_i and the film
// diagonal (A_films keep the construction-time M_inv, as in the reference)
struct TcTdepSurf {
    int nrows;
    const int* row; const int* start; const int* surf; const double* qval;      // Qs entries
    const int* fstart; const int* film; const double* adata;                     // film diagonal lists
};
static __global__ void tc_ffe_tdep_surf_kernel(TcTdepSurf L, TcTdep P, const double* __restrict__ flux,
                                               const double* __restrict__ films, const double* __restrict__ x,
                                               double* __restrict__ ydot, const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.nrows; i += gridDim.x * blockDim.x) {
        const int r = L.row[i];
        const double xi = x[r];
        const double minv = __ddiv_rn(1.0, __dmul_rn(tc_tab(P.tab_dir, P.arena, P.volcapy_tab, xi), P.Mx[r]));
        double bsum = 0.0;
        for (int k = L.start[i]; k < L.start[i + 1]; ++k) {
            const double f = flux[L.surf[k]];
            if (f != 0.0) bsum = __dadd_rn(bsum, __dmul_rn(__dmul_rn(minv, L.qval[k]), f));
        }
        double fd = 0.0;
        for (int k = L.fstart[i]; k < L.fstart[i + 1]; ++k) fd = __dadd_rn(fd, __dmul_rn(films[L.film[k]], L.adata[k]));
        ydot[r] = __dadd_rn(bsum, __dsub_rn(ydot[r], __dmul_rn(fd, xi)));
    }
}

// ---- weighted gathers: surface mean temperatures -------------------------------------
// Job j = sum_k w[k] * y[idx[k]] over its entry range; split over nblk[j] blocks whose
// partial sums are combined in block order by the network kernel (C_SURF_MEAN).
struct TcMeanJobs {
    int njobs;
    const int* job_of_block;     // grid size
    const int* first_block;      // njobs
    const int* nblk;             // njobs
    const long long* ent_start;  // njobs + 1
    const int* idx;              // index into y (absolute, state offset included)
    const double* w;
};
#define TC_MEAN_THREADS 256
static __global__ void __launch_bounds__(TC_MEAN_THREADS)
tc_surf_mean_kernel(TcMeanJobs J, const double* __restrict__ y, double* __restrict__ partials,
                    const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    __shared__ double sh[32];
    const int j = J.job_of_block[blockIdx.x];
    const int bj = blockIdx.x - J.first_block[j];
    const long long lo = J.ent_start[j], hi = J.ent_start[j + 1];
    const long long per = (hi - lo + J.nblk[j] - 1) / J.nblk[j];
    const long long b0 = lo + per * bj;
    const long long b1 = b0 + per < hi ? b0 + per : hi;
    double s = 0.0;
    for (long long k = b0 + threadIdx.x; k < b1; k += blockDim.x) s += J.w[k] * y[J.idx[k]];
    const double bs = block_sum(s, sh);
    if (threadIdx.x == 0) partials[blockIdx.x] = bs;
}

// ---- MOR parts: S independent dense r x r blocks ------------------------------------------
//   xdot_s = -(A_body_s + sum_f film_f A_film_{f,s}) x_s + B_T[s] * flux_s   thermca/solver.py:282-299
struct TcMor {
    int S, r, F;
    const double* A_body;   // S*r*r
    const double* A_films;  // F*S*r*r
    const double* B_T;      // S*r
};
static __global__ void tc_mor_rhs_kernel(TcMor M, const double* __restrict__ films,
                                  const double* __restrict__ flux, const double* __restrict__ x,
                                  double* __restrict__ xdot, const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    const int rows = M.S * M.r;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const size_t blk = (size_t)M.S * M.r * M.r;
    for (int row = warp; row < rows; row += nwarps) {
        const int s = row / M.r, i = row - s * M.r;
        const double* ab = M.A_body + ((size_t)s * M.r + i) * M.r;
        const double* xs = x + (size_t)s * M.r;
        double acc = 0.0;
        for (int j = lane; j < M.r; j += 32) {
            double fsum = 0.0;       // sum(f * A_film) starts from 0 like Python's sum()
            for (int f = 0; f < M.F; ++f)
                fsum = __dadd_rn(fsum, __dmul_rn(films[f], M.A_films[f * blk + ((size_t)s * M.r + i) * M.r + j]));
            const double aij = M.F > 0 ? __dadd_rn(ab[j], fsum) : ab[j];
            acc += aij * xs[j];
        }
        acc = warp_sum(acc);
        if (lane == 0) xdot[row] = __dadd_rn(-acc, __dmul_rn(M.B_T[row], flux[s]));
    }
}
