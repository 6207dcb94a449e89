// Fused network kernel: one CTA replays, per right-hand-side call, everything the
// reference does at network level (all of it is O(#nodes + #links), ~10..10^3 items):
//   * surface / margin mean temperatures into the node vector  thermca/solver.py:208-220
//   * Input step tables                                         thermca/input.py:135-146
//   * material tables, flow / film / conductance closures, heat, flux and
//     bound-temperature closures, with the integer-indexed scatter into the CSR
//     conductance data                                          thermca/network.py:757-796
//   * film coefficients of parts from the cleaned conductance   thermca/network.py:803-817
//   * LP film (+) margin series conductance, LP<->LP exchange, margin leak
//                                       thermca/lpm/lp_io.py:68-103, network.py:828-862
//   * stationary nodes (Gauss-Seidel, list order)               thermca/solver.py:30-42
//   * differential-equation form + network derivative           thermca/solver.py:45-58,239-255
//   * per-part surface flux vectors                             thermca/solver.py:260-290
// The work list is a command stream built by thermca_b200/device.py in the reference's
// statement order; commands are separated by __syncthreads() so later groups see
// earlier writes exactly like the sequential Python does.
#pragma once
#include "vm.cuh"

enum TcCmd {
    C_END = 0, C_SET_TU, C_SURF_MEAN, C_INPUTS, C_MATL_GROUP, C_CAPY, C_FLOW_CONST, C_FLOW_FUNC,
    C_FILM_FUNC, C_HEAT_FUNC, C_FLUX_FUNC, C_BOUND_FUNC, C_PART_FILMS, C_LP_UPDATE, C_LP2LP,
    C_LP_LEAK, C_STAT, C_DIAG, C_NET_DERIV, C_PART_FLUX, C_LP_LOAD, C_COPY_D, C_LP_TDEP, C_FUNCS, C_JTHETA
};

struct TcNetParams {
    const int* iarena;
    double* darena;
    const int* cmds;          // command stream (in iarena or separate)
    const int* prog_dir;
    const int* tab_dir;
    int N, u, nnz;
    int o_T, o_heat, o_capy, o_vol, o_volcapy, o_condy, o_cond, o_clean, o_calc, o_inputs;  // darena
    int o_indptr, o_indices, o_diag;                                                         // iarena
    int n_iarena, n_darena;   // arena lengths in words
};

#define TC_NET_THREADS 256

// time = ctrl[0] + c * ctrl[1]   (ctrl lives on the device: t, h of the running step)
__global__ void __launch_bounds__(TC_NET_THREADS, 1)
tc_network_kernel(TcNetParams P, const double* __restrict__ ctrl, double c_stage,
                  const double* __restrict__ y, double* __restrict__ ydot,
                  const double* __restrict__ partials, const int* __restrict__ done_flag) {
    if (done_flag && *done_flag) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int* I = P.iarena;
    double* D = P.darena;
    double* T = D + P.o_T;
    double* heat = D + P.o_heat;
    double* capy = D + P.o_capy;
    double* volcapy = D + P.o_volcapy;
    double* condy = D + P.o_condy;
    double* cond = D + P.o_cond;
    double* clean = D + P.o_clean;
    double* calc = D + P.o_calc;
    double* inputs = D + P.o_inputs;
    const int* indptr = I + P.o_indptr;
    const int* indices = I + P.o_indices;
    TcTables tb{P.tab_dir, D};
    double reg[TC_VM_MAX];
    // The command interpreter and the link programs are chains of dependent loads; a fresh launch starts with a
    // cold L1, so each first touch of a line would be an L2 round trip in that chain.  Pull the (small) arenas
    // into this SM's L1 with one parallel sweep instead (bounded: large LP operators at the arena tail are left out).
    {
        const int li = min(P.n_iarena, 16384) >> 5, ld = min(P.n_darena, 12288) >> 4;     // 128-byte lines
        for (int l = tid; l < li; l += nt) asm volatile("prefetch.global.L1 [%0];" ::"l"(I + 32 * l));
        for (int l = tid; l < ld; l += nt) asm volatile("prefetch.global.L1 [%0];" ::"l"(D + 16 * l));
    }
    const double time = __dadd_rn(ctrl[0], __dmul_rn(c_stage, ctrl[1]));

    const int* pc = P.cmds;
    for (;;) {
        const int op = pc[0];
        const int len = pc[1];          // number of argument words
        const int* a = pc + 2;
        if (op == C_END) break;
        switch (op) {
        case C_SET_TU:
            for (int i = tid; i < P.u; i += nt) T[i] = y[i];
            break;
        case C_SURF_MEAN: {             // a: njobs, then {partial_off, nblk, target, ref_doff}
            const int nj = a[0];
            for (int j = tid; j < nj; j += nt) {
                const int* q = a + 1 + 4 * j;
                double s = 0.0;
                for (int b = 0; b < q[1]; ++b) s += partials[q[0] + b];
                if (q[3] >= 0) s += D[q[3]];
                T[q[2]] = s;
            }
        } break;
        case C_INPUTS: {                // a: n, then {t_off, v_off, len}
            const int n = a[0];
            for (int k = tid; k < n; k += nt) {
                const int* q = a + 1 + 3 * k;
                const double* tt = D + q[0];
                int cnt = 0;            // bisect_right(t, time)
                for (int i = 0; i < q[2]; ++i) cnt += (tt[i] <= time) ? 1 : 0;
                int idx = cnt - 1;
                if (idx < 0) idx += q[2];   // Python negative index wraps
                inputs[k] = D[q[1] + idx];
            }
        } break;
        case C_MATL_GROUP: {            // n, idx_off, tab_vc, tab_k
            const int* idx = I + a[1];
            for (int i = tid; i < a[0]; i += nt) {
                const double mt = T[idx[i]];
                volcapy[idx[i]] = tc_interp(tb, a[2], mt);
                condy[idx[i]] = tc_interp(tb, a[3], mt);
            }
        } break;
        case C_CAPY: {                  // n, idx_off
            const int* idx = I + a[1];
            const double* vol = D + P.o_vol;
            for (int i = tid; i < a[0]; i += nt) capy[idx[i]] = __dmul_rn(vol[idx[i]], volcapy[idx[i]]);
        } break;
        case C_FLOW_CONST: {            // n, didx, i0, i1, vf_doff
            const int *dx = I + a[1], *i0 = I + a[2], *i1 = I + a[3];
            const double* vf = D + a[4];
            for (int i = tid; i < a[0]; i += nt)
                cond[dx[i]] = __ddiv_rn(__dmul_rn(vf[i], __dadd_rn(volcapy[i0[i]], volcapy[i1[i]])), 2.0);
        } break;
        // C_FLOW_FUNC .. C_BOUND_FUNC: superseded by C_FUNCS (all closure groups in one command); the enum values stay
        // reserved so that the command numbering of thermca_b200/device.py does not shift
        case C_FUNCS: {
            // All closure groups of one right-hand side at once.  The closures read only node temperatures,
            // Inputs and material tables (network.py:773-796 calls them before any bound temperature is
            // written back), so phase 1 evaluates every group concurrently - one warp per group, because
            // different programs diverge, lanes = the links of the group - and phase 2 applies the results
            // group by group in the reference's statement order (later groups overwrite earlier ones).
            // a: ngroups, res_doff, then per group 10 words {kind, prog, n, res_start, p0..p5}:
            //   kind 0 flow: didx, i0, i1 | 1 film: fwd, bwd, i0, i1, areas_doff | 2 heat: m, idxs, counts_doff
            //   | 3 flux: idx, areas_doff | 4 bound: idx
            const int ng = a[0];
            double* res = D + a[1];
            const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
            for (int g = warp; g < ng; g += nw) {
                const int* q = a + 2 + 10 * g;
                const int kind = q[0], prog = q[1], n = q[2];
                double* out = res + q[3];
                for (int i = lane; i < n; i += 32) {
                    double x0 = 0.0, x1 = 0.0;
                    if (kind <= 1) {
                        const int *i0 = I + (kind == 0 ? q[5] : q[6]), *i1 = I + (kind == 0 ? q[6] : q[7]);
                        x0 = T[i0[i]]; x1 = T[i1[i]];
                    } else if (kind == 2) {
                        const int m = q[4];
                        const int* idx = I + q[5];
                        double sum = 0.0;
                        for (int j = 0; j < m; ++j) sum += T[idx[i * m + j]];
                        x0 = __ddiv_rn(sum, (double)m);
                    } else if (kind == 3) {
                        x0 = T[(I + q[4])[i]];
                    }
                    out[i] = tc_vm_run(I, D, P.prog_dir, prog, tb, inputs, x0, x1, reg);
                }
            }
            __syncthreads();
            for (int g = 0; g < ng; ++g) {
                const int* q = a + 2 + 10 * g;
                const int kind = q[0], n = q[2];
                const double* f = res + q[3];
                if (kind == 0) {
                    const int *dx = I + q[4], *i0 = I + q[5], *i1 = I + q[6];
                    for (int i = tid; i < n; i += nt)
                        cond[dx[i]] = __ddiv_rn(__dmul_rn(f[i], __dadd_rn(volcapy[i0[i]], volcapy[i1[i]])), 2.0);
                } else if (kind == 1) {
                    const int *fw = I + q[4], *bw = I + q[5];
                    const double* ar = D + q[8];
                    for (int i = tid; i < n; i += nt) {
                        const double cv = __dmul_rn(f[i], ar[i]);
                        clean[fw[i]] = cv; clean[bw[i]] = cv;
                        cond[fw[i]] = cv; cond[bw[i]] = cv;
                    }
                } else if (kind == 2) {
                    const int m = q[4];
                    const int* idx = I + q[5];
                    const double* cnt = D + q[6];
                    for (int k = tid; k < n; k += nt) {
                        const double hq = __ddiv_rn(f[k], cnt[k]);
                        for (int j = 0; j < m; ++j) heat[idx[k * m + j]] = hq;
                    }
                } else if (kind == 3) {
                    const int* idx = I + q[4];
                    const double* ar = D + q[5];
                    for (int i = tid; i < n; i += nt) heat[idx[i]] = __dmul_rn(f[i], ar[i]);
                } else {
                    const int* idx = I + q[4];
                    for (int i = tid; i < n; i += nt) T[idx[i]] = f[i];
                }
                __syncthreads();
            }
        } break;
        case C_PART_FILMS: {            // F, films_doff, rc_idx, dock_doff
            const int* rc = I + a[2];
            for (int f = tid; f < a[0]; f += nt) D[a[1] + f] = __ddiv_rn(clean[rc[f]], D[a[3] + f]);
        } break;
        case C_LP_UPDATE: {
            // dof, F, films_doff, lfm_doff (dof*F row-major), lsum_doff (F), Ldata_doff,
            // face_dir ioff (F x {areas_doff, smx_ioff, marg_ioff, count}), minv_doff,
            // abody_diag_doff (dof), sellval_doff, diagpos_ioff (position of A_ii in SELL values),
            // adiag_doff (dense copy of the updated diagonal, used by the implicit solves)
            const int dof = a[0], F = a[1];
            double* lfm = D + a[3];
            for (int i = tid; i < dof * F; i += nt) lfm[i] = 0.0;
            __syncthreads();
            // series conductance per face, accumulated in face order (numpy add.at)
            for (int f = tid; f < F; f += nt) {
                const int* q = I + a[6] + 4 * f;
                const double* fa = D + q[0];
                const int* smx = I + q[1];
                const int* mg = I + q[2];
                const double film = D[a[2] + f];
                for (int j = 0; j < q[3]; ++j) {
                    const double fc = __dmul_rn(fa[j], film);
                    const double mc = -D[a[5] + smx[j]];
                    const double sv = __ddiv_rn(__dmul_rn(fc, mc), __dadd_rn(fc, mc));
                    lfm[mg[j] * F + f] = __dadd_rn(lfm[mg[j] * F + f], sv);
                }
            }
            __syncthreads();
            for (int f = tid; f < F; f += nt) {     // column sums (axis 0, row order)
                double s = 0.0;
                for (int i = 0; i < dof; ++i) s = __dadd_rn(s, lfm[i * F + f]);
                D[a[4] + f] = s;
            }
            // diagonal of the state matrix: A_body_ii + M_inv_i * sum_f Lfm[i, f], written
            // straight into the part's SELL value array (thermca/lpm/lp_io.py:76-79)
            const int* dpos = I + a[10];
            for (int i = tid; i < dof; i += nt) {
                double rs = 0.0;
                for (int f = 0; f < F; ++f) rs = __dadd_rn(rs, lfm[i * F + f]);
                const double dv = __dadd_rn(D[a[8] + i], __dmul_rn(D[a[7] + i], rs));
                D[a[9] + dpos[i]] = dv;
                D[a[11] + i] = dv;
            }
        } break;
        case C_LP2LP: {                 // films0_doff+f0, lsum1_doff+f1, area0_doff ; films1.., lsum0.., area1..
            if (tid == 0) {
                D[a[0]] = __ddiv_rn(D[a[1]], D[a[2]]);
                D[a[3]] = __ddiv_rn(D[a[4]], D[a[5]]);
            }
        } break;
        case C_LP_LEAK: {               // F, fwd, bwd, lsum_doff
            const int *fw = I + a[1], *bw = I + a[2];
            for (int f = tid; f < a[0]; f += nt) { cond[fw[f]] = D[a[3] + f]; cond[bw[f]] = D[a[3] + f]; }
        } break;
        case C_STAT: {                  // n, idx  (sequential Gauss-Seidel)
            if (tid == 0) {
                const int* idx = I + a[1];
                for (int s = 0; s < a[0]; ++s) {
                    const int i = idx[s];
                    double sum_cond = 0.0, ap = heat[i];
                    for (int j = indptr[i]; j < indptr[i + 1]; ++j) {
                        sum_cond = __dadd_rn(sum_cond, cond[j]);
                        ap = __dadd_rn(ap, __dmul_rn(cond[j], T[indices[j]]));
                    }
                    if (sum_cond == 0.0) continue;
                    T[i] = __dsub_rn(T[i], __dsub_rn(T[i], __ddiv_rn(ap, sum_cond)));
                }
            }
        } break;
        case C_DIAG: {
            for (int i = tid; i < P.nnz; i += nt) calc[i] = -cond[i];
            __syncthreads();
            const int* dg = I + P.o_diag;
            for (int r = tid; r < P.u; r += nt) {
                double s = 0.0;
                for (int j = indptr[r]; j < indptr[r + 1]; ++j) s = __dadd_rn(s, calc[j]);
                calc[dg[r]] = -s;
            }
        } break;
        case C_NET_DERIV: {             // inv_capy * (q_u - L_uk T_k - L_uu T_u)
            for (int r = tid; r < P.u; r += nt) {
                double suu = 0.0, suk = 0.0;
                for (int j = indptr[r]; j < indptr[r + 1]; ++j) {
                    const int cidx = indices[j];
                    const double pr = __dmul_rn(calc[j], T[cidx]);
                    if (cidx < P.u) suu = __dadd_rn(suu, pr); else suk = __dadd_rn(suk, pr);
                }
                const double ic = __ddiv_rn(1.0, capy[r]);
                ydot[r] = __dmul_rn(ic, __dsub_rn(__dsub_rn(heat[r], suk), suu));
            }
        } break;
        case C_PART_FLUX: {
            // S, F, flux_doff, surf_net_ioff, areas_doff, film_surf_ioff, films_doff, link_ioff, ref_doff(-1)
            const int S = a[0], F = a[1];
            double* flux = D + a[2];
            const int* sn = I + a[3];
            for (int s = tid; s < S; s += nt) flux[s] = __ddiv_rn(heat[sn[s]], D[a[4] + s]);
            __syncthreads();
            if (tid == 0) {             // numpy add.at: accumulate in film order
                const int *fs = I + a[5], *lk = I + a[7];
                const double ref = a[8] >= 0 ? D[a[8]] : 0.0;
                for (int f = 0; f < F; ++f) {
                    const double tl = a[8] >= 0 ? __dsub_rn(T[lk[f]], ref) : T[lk[f]];
                    flux[fs[f]] = __dadd_rn(flux[fs[f]], __dmul_rn(D[a[6] + f], tl));
                }
            }
        } break;
        case C_LP_LOAD: {
            // dof, S, F, b_doff, qs_doff (dof*S), surf_net_ioff, areas_doff, lfm_doff (dof*F),
            // link_ioff, minv_doff
            const int dof = a[0], S = a[1], F = a[2];
            const int *sn = I + a[5], *lk = I + a[8];
            for (int i = tid; i < dof; i += nt) {
                double flow = 0.0;
                for (int s = 0; s < S; ++s)
                    flow = __dadd_rn(flow, __dmul_rn(D[a[4] + i * S + s], __ddiv_rn(heat[sn[s]], D[a[6] + s])));
                double fl2 = 0.0;
                for (int f = 0; f < F; ++f) fl2 = __dadd_rn(fl2, __dmul_rn(D[a[7] + i * F + f], T[lk[f]]));
                flow = __dadd_rn(flow, fl2);
                D[a[3] + i] = __dmul_rn(D[a[9] + i], flow);
            }
        } break;
        case C_LP_TDEP: {
            // Temperature dependent LP body (LPSystem.update_material_properties, lpm/lp_system.py:106-118,
            // LPIO.update_material_properties lp_io.py:59-66).
            // a: dof, n_rows, nnzL, y_off, Lx_doff, row_ioff, col_ioff, Ldata_doff, Lindptr_ioff, diag_ioff,
            //    Mx_doff, minv_doff, mass_doff, condy_tab, vc_tab, tall_doff, nsurf, surfdir_ioff
            //    (nsurf x {marg_ioff, surfidx_ioff, count}), n_ab, fromL_ioff, abrow_ioff, sellpos_ioff,
            //    val_doff, abody_diag_doff
            const int dof = a[0], n_rows = a[1], nnzL = a[2];
            const double* x = y + a[3];
            double* Tall = D + a[15];
            for (int i = tid; i < dof; i += nt) Tall[i] = x[i];
            for (int s = 0; s < a[16]; ++s) {
                const int* q = I + a[17] + 3 * s;
                const int *mg = I + q[0], *sf = I + q[1];
                for (int j = tid; j < q[2]; j += nt) Tall[sf[j]] = x[mg[j]];
            }
            __syncthreads();
            const int *row = I + a[5], *col = I + a[6];
            double* Ld = D + a[7];
            for (int k = tid; k < nnzL; k += nt) {
                const double ct = __ddiv_rn(__dadd_rn(Tall[row[k]], Tall[col[k]]), 2.0);
                Ld[k] = __dmul_rn(tc_interp(tb, a[13], ct), D[a[4] + k]);
            }
            __syncthreads();
            const int *lip = I + a[8], *dg = I + a[9];
            for (int i = tid; i < dof; i += nt) {
                double sum = 0.0;
                for (int k = lip[i]; k < lip[i + 1]; ++k) sum = __dadd_rn(sum, Ld[k]);
                Ld[dg[i]] = -sum;
                const double md = __dmul_rn(tc_interp(tb, a[14], x[i]), D[a[10] + i]);
                D[a[12] + i] = md;
                D[a[11] + i] = __ddiv_rn(1.0, md);
            }
            __syncthreads();
            const int *fl = I + a[19], *abr = I + a[20], *sp = I + a[21];
            for (int k = tid; k < a[18]; k += nt) D[a[22] + sp[k]] = __dmul_rn(D[a[11] + abr[k]], Ld[fl[k]]);
            for (int i = tid; i < dof; i += nt) D[a[23] + i] = __dmul_rn(D[a[11] + i], Ld[dg[i]]);
            (void)n_rows;
        } break;
        case C_JTHETA: {
            // For the Jacobian of the implicit W-methods only (tc_set_jac_films; the right-hand side is untouched):
            // film f of a part leads to the stationary node n = link[f].  T_n = (q_n + sum_k g_k T_k) / G
            // (C_STAT above), so dT_n / dT_surface = theta = g_f / G with g_f the entry of the part's surface
            // node in row n and G the row sum.  A node that only carries a heat source (the reference's idiom
            // for a prescribed surface heat flow) has theta = 1: the flux does not depend on the surface.
            // a: F, theta_doff, link_ioff, surfnode_ioff, statflag_ioff
            const int *lk = I + a[2], *sn = I + a[3], *sf = I + a[4];
            for (int f = tid; f < a[0]; f += nt) {
                double theta = 0.0;
                if (sf[f]) {
                    double G = 0.0, g = 0.0;
                    for (int j = indptr[lk[f]]; j < indptr[lk[f] + 1]; ++j) {
                        G = __dadd_rn(G, cond[j]);
                        if (indices[j] == sn[f]) g = __dadd_rn(g, cond[j]);
                    }
                    if (G > 0.0) theta = __ddiv_rn(g, G);
                }
                D[a[1] + f] = theta;
            }
        } break;
        case C_COPY_D: {                // n, src_doff, dst_doff
            for (int i = tid; i < a[0]; i += nt) D[a[2] + i] = D[a[1] + i];
        } break;
        default: break;
        }
        __syncthreads();
        pc += 2 + len;
    }
}
