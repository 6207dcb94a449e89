// Batched MOR ensemble step (BASELINE config #3): B load cases of one MOR-reduced FE body
// advanced in lockstep.  The reference evaluates one scenario at a time in Python
// (thermca/solver.py:282-299, thermca/fem/mor_io.py:90-92):
//     A_s = A_body_s + sum_f film_f A_film_{f,s};   xdot_s = -A_s x_s + B_T[s] * flux_s
// Here the scenario axis is the GEMM N dimension: for every surface s
//     Z_o (r x B) = W_{o,s} (r x r) . X_s (r x B),  o = 0 (body), 1..F (films)
// is a true dense contraction done with FP64 tensor-core DMMA (mma.sync m8n8k4 f64;
// tcgen05 has no f64 kind), and the per-scenario film weighting, the input term and the
// Runge-Kutta stage update are fused into the epilogue, so A_s is never materialised per
// scenario (that would be S*r^2*8 B = ~1 MB per scenario and HBM-bound).
// Layout: X[s][i][b], scenario index fastest.  Classical RK4, fixed step, four launches
// per step.
#include "common.cuh"
#include "../../include/thermca_b200.h"

#define MB_TM 64      // rows per CTA tile
// scenarios per CTA tile: 16*NC (NC = 8-column DMMA tiles per warp; 4 for <=4 operators, 2 above)
// k chunk: 16 (8 when more than 4 operators share the 48 KB of static shared memory)
#define MB_THREADS 256

struct MorBatch {
    int S, r, F, B;
    const double* A_body;   // [S][r][r]
    const double* A_films;  // [F][S][r][r]
    const double* B_T;      // [S][r]
    const int* film_surf;   // [F]
    const double* film_h0;  // [F][B]
    const double* film_tau; // [F][B]
    const double* t_link;   // [F]   link temperature minus reference temperature
    const double* q_surf;   // [S]   constant flux density per surface
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ double film_at(const MorBatch& M, int f, int b, double t) {
    const double TWO_PI = 6.283185307179586476925;
    return M.film_h0[(size_t)f * M.B + b] * (1.0 + 0.5 * sin(TWO_PI * t / M.film_tau[(size_t)f * M.B + b]));
}

// One RK stage:  k = f(t, Xin);  Xout = Xbase + c_out*h*k (if Xout);  Xacc = (first ? Xbase : Xacc) + c_acc*h*k
template <int NOPS, int NC, bool DMMA>
__global__ void __launch_bounds__(MB_THREADS)
tc_mor_stage_kernel(MorBatch M, const double* __restrict__ Xin, const double* __restrict__ Xbase,
                    double* __restrict__ Xout, double* __restrict__ Xacc, double t, double h,
                    double c_out, double c_acc, int first) {
    constexpr int MB_TK = NOPS > 4 ? 8 : 16;
    __shared__ double sA[NOPS][MB_TM][MB_TK + 1];
    constexpr int MB_TN = 16 * NC;
    __shared__ double sX[MB_TK][MB_TN + 8];
    const int s = blockIdx.z;
    const int i0 = blockIdx.y * MB_TM;
    const int b0 = blockIdx.x * MB_TN;
    const int r = M.r, B = M.B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp & 3) * 16;     // warp tile origin (rows)   : 4 x 16
    const int wn = (warp >> 2) * 8 * NC; // warp tile origin (cols)   : 2 x 8*NC
    const double* Xs = Xin + (size_t)s * r * B;
    double acc[NOPS][2][NC][2];
#pragma unroll
    for (int o = 0; o < NOPS; ++o)
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[o][a][c][0] = acc[o][a][c][1] = 0.0;

    for (int k0 = 0; k0 < r; k0 += MB_TK) {
        // operator chunks: NOPS x (64 x 16), row-major in global (k contiguous)
        for (int e = tid; e < NOPS * MB_TM * MB_TK; e += MB_THREADS) {
            const int o = e / (MB_TM * MB_TK);
            const int rem = e - o * (MB_TM * MB_TK);
            const int ri = rem / MB_TK, kk = rem - ri * MB_TK;
            const int gi = i0 + ri, gk = k0 + kk;
            double v = 0.0;
            if (gi < r && gk < r) {
                const double* W = (o == 0) ? M.A_body + (size_t)s * r * r
                                           : M.A_films + ((size_t)(o - 1) * M.S + s) * r * r;
                v = W[(size_t)gi * r + gk];
            }
            sA[o][ri][kk] = v;
        }
        // state chunk: 16 x 64 (scenario contiguous)
        for (int e = tid; e < MB_TK * MB_TN; e += MB_THREADS) {
            const int kk = e / MB_TN, bb = e - kk * MB_TN;
            const int gk = k0 + kk, gb = b0 + bb;
            sX[kk][bb] = (gk < r && gb < B) ? Xs[(size_t)gk * B + gb] : 0.0;
        }
        __syncthreads();
        if (DMMA) {
#pragma unroll
            for (int k4 = 0; k4 < MB_TK; k4 += 4) {
                double bf[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) bf[c] = sX[k4 + (lane & 3)][wn + 8 * c + (lane >> 2)];
#pragma unroll
                for (int o = 0; o < NOPS; ++o)
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        const double af = sA[o][wm + 8 * a + (lane >> 2)][k4 + (lane & 3)];
#pragma unroll
                        for (int c = 0; c < NC; ++c) dmma884(acc[o][a][c][0], acc[o][a][c][1], af, bf[c]);
                    }
            }
        } else {
            // same thread->element ownership as the DMMA C fragment, plain FMA
#pragma unroll 4
            for (int kk = 0; kk < MB_TK; ++kk)
#pragma unroll
                for (int o = 0; o < NOPS; ++o)
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        const double af = sA[o][wm + 8 * a + (lane >> 2)][kk];
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            acc[o][a][c][0] += af * sX[kk][wn + 8 * c + 2 * (lane & 3)];
                            acc[o][a][c][1] += af * sX[kk][wn + 8 * c + 2 * (lane & 3) + 1];
                        }
                    }
        }
        __syncthreads();
    }
    // fused epilogue: film weighting, input term, RK stage update
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gi = i0 + wm + 8 * a + (lane >> 2);
                const int gb = b0 + wn + 8 * c + 2 * (lane & 3) + e;
                if (gi >= r || gb >= B) continue;
                double z = acc[0][a][c][e];
                double flux = M.q_surf[s];
#pragma unroll
                for (int f = 0; f < NOPS - 1; ++f) {      // NOPS == F + 1
                    const double fv = film_at(M, f, gb, t);
                    z += fv * acc[f + 1][a][c][e];
                    if (M.film_surf[f] == s) flux += fv * M.t_link[f];
                }
                const double kval = -z + M.B_T[(size_t)s * r + gi] * flux;
                const size_t idx = ((size_t)s * r + gi) * B + gb;
                const double xb = Xbase[idx];
                if (Xout) Xout[idx] = xb + c_out * h * kval;
                Xacc[idx] = (first ? xb : Xacc[idx]) + c_acc * h * kval;
            }
}


// ---- pipelined variant: cp.async (LDGSTS) double buffering, conflict-free padded tiles, 2 CTAs/SM ----
// Same math and epilogue as tc_mor_stage_kernel<.., true>; the k-chunks of the operator tiles and of
// the state tile are prefetched one chunk ahead so the DMMA pipe is not stalled on L2 latency.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

struct TcTrue { static constexpr bool value = true; };
struct TcFalse { static constexpr bool value = false; };

#define MP_TK 16
#define MP_LDA (MP_TK + 4)     // 20 doubles: rows 16-byte aligned, A-fragment reads conflict free
#define MP_TN 64
#define MP_LDX (MP_TN + 4)     // 68 doubles

// film coefficients of every scenario at the three stage times t, t+h/2, t+h of one RK4 step, evaluated once
// per step (3*F*B values) instead of once per output element inside the GEMM epilogue
__global__ void tc_mor_films_kernel(MorBatch M, double t, double h, double* __restrict__ fv) {
    const int fb = M.F * M.B;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * fb; i += gridDim.x * blockDim.x) {
        const int k = i / fb, j = i - k * fb;
        const int f = j / M.B;
        fv[i] = film_at(M, f, j - f * M.B, t + 0.5 * h * k);
    }
}

// dynamic shared memory of the pipelined kernel: two stages of (NOPS x 72 x 20 operator + 16 x 68 state)
// doubles, reused as the 72 x 72 staging tile of the epilogue
__host__ __device__ constexpr int mor_pipe_smem_doubles(int nops) {
    return 2 * (nops * (MB_TM + 8) * MP_LDA + MP_TK * MP_LDX) > (MB_TM + 8) * (MP_TN + 8)
               ? 2 * (nops * (MB_TM + 8) * MP_LDA + MP_TK * MP_LDX) : (MB_TM + 8) * (MP_TN + 8);
}
// + the epilogue's small operands, fetched with the first chunk: film values of the tile's scenario columns
// [F][64] and the B_T rows of the tile [72]
__host__ __device__ constexpr int mor_pipe_smem_total(int nops) {
    return mor_pipe_smem_doubles(nops) + (nops > 1 ? nops - 1 : 1) * MP_TN + MB_TM + 8;
}

// a remainder of 1..8 rows past the last full 64-row tile is folded into that tile
__host__ __device__ __forceinline__ bool mor_fold_rows(int r) {
    const int rem = r % MB_TM;
    return r > MB_TM && rem > 0 && rem <= 8;
}

template <int NOPS>
__global__ void __launch_bounds__(MB_THREADS, (NOPS <= 3 ? 2 : 1))
tc_mor_stage_pipe_kernel(MorBatch M, const double* __restrict__ fv, const double* __restrict__ Xin,
                         const double* __restrict__ Xbase,
                         double* __restrict__ Xout, double* __restrict__ Xacc, double t, double h,
                         double c_out, double c_acc, int first) {
    extern __shared__ __align__(16) double smem[];
    asm volatile("griddepcontrol.launch_dependents;");      // the next stage's CTAs may fill the tail of this grid
    constexpr int TMX = MB_TM + 8;              // rows per operator buffer: tile + one folded 8-row strip
    constexpr int A_STAGE = NOPS * TMX * MP_LDA;
    constexpr int X_STAGE = MP_TK * MP_LDX;
    double* sA0 = smem;                         // [2][NOPS][72][20]
    double* sX0 = smem + 2 * A_STAGE;           // [2][16][68]
    double* sF = smem + mor_pipe_smem_doubles(NOPS);                         // [NOPS-1][64] film values of the columns
    double* sBT = sF + (NOPS > 1 ? NOPS - 1 : 1) * MP_TN;                    // [72] B_T rows of the tile
    // CTA order: scenario tile fastest, then surface, row tile slowest, so the cheap ragged last row tiles
    // (r % 64 rows) are scheduled last and fill the tail of the final wave
    const int lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    const int rest = lin / gridDim.x;
    const int s = rest % gridDim.z;
    const int ty = rest / gridDim.z;
    const int i0 = ty * MB_TM;
    const int b0 = (lin - rest * gridDim.x) * MP_TN;
    const int r = M.r, B = M.B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp & 3) * 16;
    const int wn = (warp >> 2) * 32;
    // A remainder of at most 8 rows (r = 200: 3 x 64 + 8) is folded into the last full tile: every warp takes
    // one extra 8x8 block of that strip (columns wn + 8*(warp&3)), so no latency-bound sliver tile exists.
    const bool ext = mor_fold_rows(r) && ty == (int)gridDim.y - 1;      // CTA-uniform
    const double* Xs = Xin + (size_t)s * r * B;
    double acc[NOPS][2][4][2];
    double accx[NOPS][2];
#pragma unroll
    for (int o = 0; o < NOPS; ++o)
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[o][a][c][0] = acc[o][a][c][1] = 0.0;

    // every thread owns the same copy slots in every chunk (MB_THREADS = 256 = 32 rows x 8 pairs):
    // operator slot q -> operator q/2, row (tid>>3) + 32*(q&1), pair tid&7; state slot q -> k-row (tid>>5) + 8q;
    // folded strip: thread -> operator tid>>6, row 64 + ((tid>>3)&7), pair tid&7
    static_assert(MB_THREADS == 256 && MB_TM == 64 && MP_TK == 16 && MP_TN == 64, "copy slot mapping");
    const int a_ri = tid >> 3, a_kp = tid & 7;
    const bool a_in[2] = {i0 + a_ri < r, i0 + a_ri + 32 < r};
    const size_t a_off = (size_t)(i0 + a_ri) * r + 2 * a_kp;
    const int a_dst = a_ri * MP_LDA + 2 * a_kp;
    const double* Wb = M.A_body + (size_t)s * r * r;
    const double* Wf = M.A_films + (size_t)s * r * r;
    const size_t w_stride = (size_t)M.S * r * r;
    const int x_kk = tid >> 5, x_bp = tid & 31;
    const bool x_in = b0 + 2 * x_bp < B;
    const size_t x_off = (size_t)x_kk * B + b0 + 2 * x_bp;
    const int x_dst = x_kk * MP_LDX + 2 * x_bp;
    auto load_ops = [&](int stage, int k0) {
        double* sA = sA0 + stage * A_STAGE;
        const bool k_in = k0 + 2 * a_kp < r;            // r is even: a pair is either inside or outside
#pragma unroll
        for (int q = 0; q < 2 * NOPS; ++q) {
            const int o = q >> 1, half = q & 1;
            const double* W = o == 0 ? Wb : Wf + (size_t)(o - 1) * w_stride;
            const bool ok = k_in && a_in[half];
            cp_async16(sA + a_dst + (o * TMX + 32 * half) * MP_LDA,
                       ok ? (const void*)(W + a_off + (size_t)(32 * half) * r + k0) : (const void*)Wb, ok ? 16 : 0);
        }
        if (ext && tid < NOPS * 64) {
            const int e_o = tid >> 6, e_ri = (tid >> 3) & 7;
            const double* W = e_o == 0 ? Wb : Wf + (size_t)(e_o - 1) * w_stride;
            const bool ok = k_in && i0 + MB_TM + e_ri < r;
            cp_async16(sA + (e_o * TMX + MB_TM + e_ri) * MP_LDA + 2 * a_kp,
                       ok ? (const void*)(W + (size_t)(i0 + MB_TM + e_ri) * r + 2 * a_kp + k0) : (const void*)Wb, ok ? 16 : 0);
        }
    };
    auto load_state = [&](int stage, int k0) {
        double* sX = sX0 + stage * X_STAGE;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const bool ok = x_in && k0 + x_kk + 8 * q < r;
            cp_async16(sX + x_dst + 8 * q * MP_LDX,
                       ok ? (const void*)(Xs + x_off + (size_t)(k0 + 8 * q) * B) : (const void*)Xs, ok ? 16 : 0);
        }
        cp_async_commit();
    };
    auto load_chunk = [&](int stage, int k0) { load_ops(stage, k0); load_state(stage, k0); };

    const bool live[2] = {i0 + wm < r, i0 + wm + 8 < r};      // warp-uniform
    const bool full_tile = i0 + MB_TM <= r;                     // CTA-uniform
    const int nchunks = (r + MP_TK - 1) / MP_TK;
    const int cx = warp & 3;                                    // column block of the folded strip
    // The k loop.  RAGGED (a last row tile that could not be folded) skips 8-row strips that lie past r; EXTRA
    // adds the folded strip; plain full tiles run without predicates.  One barrier per chunk: once every warp
    // has passed it, chunk ch is visible AND everybody is done with chunk ch-1, whose buffer the prefetch of
    // chunk ch+1 may therefore overwrite.
    auto k_loop = [&](auto ragged, auto extra) {
        // Programmatic dependent launch: this grid may have been scheduled while the previous stage kernel was
        // still draining.  The operators do not depend on it, so their first chunk is already in flight when the
        // grid dependency (previous stage complete and flushed) is awaited; everything after reads the state.
        load_ops(0, 0);
        asm volatile("griddepcontrol.wait;" ::: "memory");
        // the epilogue's small operands ride along with the first chunk (their L2 latency is off the tail)
        if (tid < (NOPS - 1) * 32) {
            const int f = tid >> 5, cp2 = 2 * (tid & 31);
            const bool ok = b0 + cp2 < B;
            cp_async16(sF + f * MP_TN + cp2, ok ? (const void*)(fv + (size_t)f * B + b0 + cp2) : (const void*)Xs, ok ? 16 : 0);
        } else if (tid >= 128 && tid < 128 + TMX / 2) {
            const int rp = 2 * (tid - 128);
            const bool ok = i0 + rp < r;
            cp_async16(sBT + rp, ok ? (const void*)(M.B_T + (size_t)s * r + i0 + rp) : (const void*)Xs, ok ? 16 : 0);
        }
        load_state(0, 0);
        for (int ch = 0; ch < nchunks; ++ch) {
            const int st = ch & 1;
            cp_async_wait<0>();
            __syncthreads();
            if (ch + 1 < nchunks) load_chunk(st ^ 1, (ch + 1) * MP_TK);
            const double* sA = sA0 + st * A_STAGE;
            const double* sX = sX0 + st * X_STAGE;
            if (decltype(ragged)::value && !live[0]) continue;
#pragma unroll
            for (int k4 = 0; k4 < MP_TK; k4 += 4) {
                double bf[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) bf[c] = sX[(k4 + (lane & 3)) * MP_LDX + wn + 8 * c + (lane >> 2)];
#pragma unroll
                for (int o = 0; o < NOPS; ++o)
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        if (decltype(ragged)::value && !live[a]) continue;
                        const double af = sA[(o * TMX + wm + 8 * a + (lane >> 2)) * MP_LDA + k4 + (lane & 3)];
#pragma unroll
                        for (int c = 0; c < 4; ++c) dmma884(acc[o][a][c][0], acc[o][a][c][1], af, bf[c]);
                    }
                if (decltype(extra)::value) {
                    const double bx = sX[(k4 + (lane & 3)) * MP_LDX + wn + 8 * cx + (lane >> 2)];
#pragma unroll
                    for (int o = 0; o < NOPS; ++o) {
                        const double af = sA[(o * TMX + MB_TM + (lane >> 2)) * MP_LDA + k4 + (lane & 3)];
                        dmma884(accx[o][0], accx[o][1], af, bx);
                    }
                }
            }
        }
        __syncthreads();
    };
    if (ext) {
#pragma unroll
        for (int o = 0; o < NOPS; ++o) accx[o][0] = accx[o][1] = 0.0;
        k_loop(TcFalse{}, TcTrue{});
    } else if (full_tile) {
        k_loop(TcFalse{}, TcFalse{});
    } else {
        k_loop(TcTrue{}, TcFalse{});
    }

    // fused epilogue, pass 1: k = -(Z_0 + sum_f film_f Z_f) + B_T * flux per accumulator element, staged
    // through the (now free) shared memory so that pass 2 can move the state in whole 512-byte rows with
    // all its loads in flight at once (the accumulators no longer occupy the registers then).
    constexpr int LDK = MP_TN + 8;                  // 72: double2 stores of a quarter warp hit distinct banks
    static_assert(TMX * LDK <= mor_pipe_smem_doubles(NOPS), "stage tile fits the pipeline buffers");
    double* sK = smem;
    {
        const double q_s = M.q_surf[s];
        double bt[2];
#pragma unroll
        for (int a = 0; a < 2; ++a) bt[a] = sBT[wm + 8 * a + (lane >> 2)];          // zero-filled past r
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int cl = wn + 8 * c + 2 * (lane & 3);
            double fvv[NOPS > 1 ? NOPS - 1 : 1][2];
            double flux[2] = {q_s, q_s};
#pragma unroll
            for (int f = 0; f < NOPS - 1; ++f) {
                const double2 v = *reinterpret_cast<const double2*>(sF + f * MP_TN + cl);  // zero-filled past B
                fvv[f][0] = v.x; fvv[f][1] = v.y;
                if (M.film_surf[f] == s) { const double tl = M.t_link[f]; flux[0] += v.x * tl; flux[1] += v.y * tl; }
            }
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                double kv[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double z = acc[0][a][c][e];
#pragma unroll
                    for (int f = 0; f < NOPS - 1; ++f) z += fvv[f][e] * acc[f + 1][a][c][e];
                    kv[e] = -z + bt[a] * flux[e];
                }
                *reinterpret_cast<double2*>(sK + (wm + 8 * a + (lane >> 2)) * LDK + cl) = make_double2(kv[0], kv[1]);
            }
            if (ext && c == cx) {                    // folded strip: row 64 + lane/4 of the tile
                const double btx = sBT[MB_TM + (lane >> 2)];
                double kv[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double z = accx[0][e];
#pragma unroll
                    for (int f = 0; f < NOPS - 1; ++f) z += fvv[f][e] * accx[f + 1][e];
                    kv[e] = -z + btx * flux[e];
                }
                *reinterpret_cast<double2*>(sK + (MB_TM + (lane >> 2)) * LDK + cl) = make_double2(kv[0], kv[1]);
            }
        }
    }
    __syncthreads();
    // pass 2: RK stage update; thread -> rows (tid>>5) + 8j (j = 8: the folded strip), column pair tid&31
    {
        const double ho = c_out * h, ha = c_acc * h;
        const int cl = 2 * (tid & 31);
        const int gb = b0 + cl;
        const int r0 = tid >> 5;
        const int row_end = min(r, i0 + (ext ? TMX : MB_TM));
        if (gb < B) {
            double2 xb[9], xa[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const int gi = i0 + r0 + 8 * j;
                if (gi < row_end) {
                    const size_t idx = ((size_t)s * r + gi) * B + gb;
                    xb[j] = *reinterpret_cast<const double2*>(Xbase + idx);
                    if (!first) xa[j] = *reinterpret_cast<const double2*>(Xacc + idx);
                }
            }
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const int gi = i0 + r0 + 8 * j;
                if (gi < row_end) {
                    const size_t idx = ((size_t)s * r + gi) * B + gb;
                    const double2 kv = *reinterpret_cast<const double2*>(sK + (r0 + 8 * j) * LDK + cl);
                    if (Xout) *reinterpret_cast<double2*>(Xout + idx) = make_double2(xb[j].x + ho * kv.x, xb[j].y + ho * kv.y);
                    const double2 base = first ? xb[j] : xa[j];
                    *reinterpret_cast<double2*>(Xacc + idx) = make_double2(base.x + ha * kv.x, base.y + ha * kv.y);
                }
            }
        }
    }
}

template <int NOPS>
static int launch_stage_pipe(int S, int r, int B, cudaStream_t st, const MorBatch& M, double* fv, const double* Xin,
                             const double* Xbase, double* Xout, double* Xacc, double t, double h, double c_out,
                             double c_acc, int first) {
    const size_t smem = sizeof(double) * mor_pipe_smem_total(NOPS);
    // per device (function attributes are per device; a process may drive several GPUs)
    static unsigned long long configured = 0ULL;
    int devid = 0;
    TC_CUDA(cudaGetDevice(&devid));
    if (devid >= 64 || !(configured >> devid & 1ULL)) {
        TC_CUDA(cudaFuncSetAttribute(tc_mor_stage_pipe_kernel<NOPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (devid < 64) configured |= 1ULL << devid;
    }
    dim3 grid((B + MP_TN - 1) / MP_TN, mor_fold_rows(r) ? r / MB_TM : (r + MB_TM - 1) / MB_TM, S);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(MB_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const double* fvc = fv;
    TC_CUDA(cudaLaunchKernelEx(&cfg, tc_mor_stage_pipe_kernel<NOPS>, M, fvc, Xin, Xbase, Xout, Xacc, t, h, c_out, c_acc,
                               first));
    return 0;
}

extern "C" int64_t tc_mor_batch_work_doubles(int32_t S, int32_t r, int32_t F, int32_t B) {
    return 3LL * S * r * B + 3LL * (F > 0 ? F : 1) * B;
}

template <int NOPS, int NC>
static void launch_stage(bool dmma, int S, int r, int B, cudaStream_t st, const MorBatch& M, const double* Xin,
                         const double* Xbase, double* Xout, double* Xacc, double t, double h, double c_out,
                         double c_acc, int first) {
    dim3 grid((B + 16 * NC - 1) / (16 * NC), (r + MB_TM - 1) / MB_TM, S);
    if (dmma)
        tc_mor_stage_kernel<NOPS, NC, true><<<grid, MB_THREADS, 0, st>>>(M, Xin, Xbase, Xout, Xacc, t, h, c_out, c_acc, first);
    else
        tc_mor_stage_kernel<NOPS, NC, false><<<grid, MB_THREADS, 0, st>>>(M, Xin, Xbase, Xout, Xacc, t, h, c_out, c_acc, first);
}

extern "C" int tc_mor_batch_steps(void* cuda_stream, int32_t S, int32_t r, int32_t F, int32_t B,
                                  const double* A_body, const double* A_films, const double* B_T,
                                  const int32_t* film_surf, const double* film_h0, const double* film_tau,
                                  const double* t_link, const double* q_surf, double* X, double* work,
                                  double t0, double h, int32_t nsteps, int32_t use_dmma) {
    TC_CHECK(F >= 0 && F <= 5, "batched MOR supports up to 5 film links per body");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    MorBatch M{S, r, F, B, A_body, A_films, B_T, film_surf, film_h0, film_tau, t_link, q_surf};
    const size_t n = (size_t)S * r * B;
    double* o1 = work; double* o2 = work + n; double* acc = work + 2 * n;
    double* fv3 = work + 3 * n;                // [3][F][B] film values at t, t+h/2, t+h of the current step
    const bool dm = use_dmma != 0;
    const bool pipe = use_dmma == 2 && F <= 3 && (r % 2) == 0 && (B % 2) == 0;
    // tk: index of the stage time among (t, t+h/2, t+h)
    auto stage = [&](const double* Xin, const double* Xb, double* Xout, double* Xacc, double t, int tk, double c_out,
                     double c_acc, int first) {
        if (pipe) {
            double* fvtab = fv3 + (size_t)tk * F * B;
            switch (F) {
                case 0: launch_stage_pipe<1>(S, r, B, st, M, fvtab, Xin, Xb, Xout, Xacc, t, h, c_out, c_acc, first); break;
                case 1: launch_stage_pipe<2>(S, r, B, st, M, fvtab, Xin, Xb, Xout, Xacc, t, h, c_out, c_acc, first); break;
                case 2: launch_stage_pipe<3>(S, r, B, st, M, fvtab, Xin, Xb, Xout, Xacc, t, h, c_out, c_acc, first); break;
                default: launch_stage_pipe<4>(S, r, B, st, M, fvtab, Xin, Xb, Xout, Xacc, t, h, c_out, c_acc, first); break;
            }
            return;
        }
        switch (F) {
            case 0: launch_stage<1, 4>(dm, S, r, B, st, M, Xin, Xb, Xout, Xacc, t, h, c_out, c_acc, first); break;
            case 1: launch_stage<2, 4>(dm, S, r, B, st, M, Xin, Xb, Xout, Xacc, t, h, c_out, c_acc, first); break;
            case 2: launch_stage<3, 4>(dm, S, r, B, st, M, Xin, Xb, Xout, Xacc, t, h, c_out, c_acc, first); break;
            case 3: launch_stage<4, 4>(dm, S, r, B, st, M, Xin, Xb, Xout, Xacc, t, h, c_out, c_acc, first); break;
            case 4: launch_stage<5, 2>(dm, S, r, B, st, M, Xin, Xb, Xout, Xacc, t, h, c_out, c_acc, first); break;
            default: launch_stage<6, 2>(dm, S, r, B, st, M, Xin, Xb, Xout, Xacc, t, h, c_out, c_acc, first); break;
        }
    };
    double* cur = X;
    double* nxt = acc;
    for (int k = 0; k < nsteps; ++k) {
        const double t = t0 + k * h;
        if (pipe && F > 0) tc_mor_films_kernel<<<(3 * F * B + 255) / 256, 256, 0, st>>>(M, t, h, fv3);
        stage(cur, cur, o1, nxt, t, 0, 0.5, 1.0 / 6.0, 1);
        stage(o1, cur, o2, nxt, t + 0.5 * h, 1, 0.5, 1.0 / 3.0, 0);
        stage(o2, cur, o1, nxt, t + 0.5 * h, 1, 1.0, 1.0 / 3.0, 0);
        stage(o1, cur, nullptr, nxt, t + h, 2, 0.0, 1.0 / 6.0, 0);
        double* tmp = cur; cur = nxt; nxt = tmp;
    }
    if (cur != X) TC_CUDA(cudaMemcpyAsync(X, cur, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    TC_CUDA(cudaGetLastError());
    return 0;
}
