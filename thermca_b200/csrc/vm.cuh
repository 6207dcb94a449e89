// Interpreter for traced link / source programs (thermca_b200/trace.py).
//
// One thread evaluates one link: the straight-line program replays, in IEEE fp64 and in
// the same operation order, what the reference's Python closure computes with numpy
// (thermca/network.py:773-796, closures in thermca/lib/*).  `interp` follows
// numpy.interp (used by thermca/materials.py:120-130,359): clamped at both ends,
// slope*(x - xp[j]) + fp[j] without fused multiply-add.
#pragma once
#include "common.cuh"

#define TC_VM_MAX 512

enum TcOp {
    OP_ARG = 0, OP_CONST, OP_INPUT,
    OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_MIN, OP_MAX,
    OP_LT, OP_LE, OP_GT, OP_GE, OP_EQ, OP_NE, OP_AND, OP_OR,
    OP_NEG, OP_ABS, OP_SQRT, OP_EXP, OP_LOG, OP_LOG10, OP_SIN, OP_COS, OP_TAN, OP_TANH,
    OP_ARCTAN, OP_NOT, OP_SQUARE, OP_RECIP, OP_SIGN, OP_CBRT, OP_LOG2, OP_EXP2,
    OP_FLOOR, OP_CEIL, OP_ISNAN,
    OP_WHERE, OP_INTERP, OP_ARCTAN2, OP_FMOD
};

// Table directory: tab_dir[2*t] = offset of x in darena, tab_dir[2*t+1] = length;
// y follows x directly (offset + length).
struct TcTables {
    const int* dir;
    const double* arena;
};

__device__ __forceinline__ double tc_interp(const TcTables& tb, int t, double x) {
    const int off = tb.dir[2 * t], n = tb.dir[2 * t + 1];
    const double* xp = tb.arena + off;
    const double* fp = xp + n;
    if (isnan(x)) return x;
    if (n == 1) return fp[0];
    if (x <= xp[0]) return fp[0];
    if (x >= xp[n - 1]) return fp[n - 1];
    int lo = 0, hi = n - 1;          // xp[lo] <= x < xp[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= xp[mid]) lo = mid; else hi = mid;
    }
    const double slope = __ddiv_rn(__dsub_rn(fp[lo + 1], fp[lo]), __dsub_rn(xp[lo + 1], xp[lo]));
    return __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xp[lo])), fp[lo]);
}

// prog_dir[4*p .. 4*p+3] = {code offset in iarena (int4 per instruction), n_instr,
//                           consts offset in darena, result index}
__device__ double tc_vm_run(const int* __restrict__ iarena, const double* __restrict__ darena,
                            const int* __restrict__ prog_dir, int prog, const TcTables& tb,
                            const double* __restrict__ inputs, double a0, double a1,
                            double* __restrict__ regs, int stride = 1) {
    // value k lives at regs[k * stride]: stride 1 for a private array, blockDim for a shared-memory file
    struct RegFile {
        double* p; int s;
        __device__ __forceinline__ double& operator[](int k) const { return p[k * s]; }
    } reg{regs, stride};
    const int4* code = reinterpret_cast<const int4*>(iarena + prog_dir[4 * prog]);
    const int n = prog_dir[4 * prog + 1];
    const double* consts = darena + prog_dir[4 * prog + 2];
    // The interpreter is a chain of dependent latencies (fetch -> dispatch -> operands -> op), so the next
    // instruction word is fetched before the current one is dispatched, and the frequent cheap operations
    // (leaves, + - *) are decided with two compares instead of the jump table of the general switch.
    int4 nxt = code[0];
    for (int k = 0; k < n; ++k) {
        const int4 ins = nxt;
        if (k + 1 < n) nxt = code[k + 1];
        double v;
        if (ins.x <= OP_MUL && ins.x != OP_INPUT) {
            if (ins.x >= OP_ADD) {
                const double a = reg[ins.y], b = reg[ins.z];
                v = ins.x == OP_ADD ? __dadd_rn(a, b) : (ins.x == OP_SUB ? __dsub_rn(a, b) : __dmul_rn(a, b));
            } else {
                v = ins.x == OP_CONST ? consts[ins.y] : (ins.y == 0 ? a0 : a1);
            }
            reg[k] = v;
            continue;
        }
        switch (ins.x) {
            case OP_INPUT: v = inputs[ins.y]; break;
            case OP_DIV: v = __ddiv_rn(reg[ins.y], reg[ins.z]); break;
            case OP_POW: v = pow(reg[ins.y], reg[ins.z]); break;
            case OP_MIN: { double a = reg[ins.y], b = reg[ins.z]; v = (isnan(a) || isnan(b)) ? (a + b) : fmin(a, b); } break;
            case OP_MAX: { double a = reg[ins.y], b = reg[ins.z]; v = (isnan(a) || isnan(b)) ? (a + b) : fmax(a, b); } break;
            case OP_LT: v = reg[ins.y] < reg[ins.z] ? 1.0 : 0.0; break;
            case OP_LE: v = reg[ins.y] <= reg[ins.z] ? 1.0 : 0.0; break;
            case OP_GT: v = reg[ins.y] > reg[ins.z] ? 1.0 : 0.0; break;
            case OP_GE: v = reg[ins.y] >= reg[ins.z] ? 1.0 : 0.0; break;
            case OP_EQ: v = reg[ins.y] == reg[ins.z] ? 1.0 : 0.0; break;
            case OP_NE: v = reg[ins.y] != reg[ins.z] ? 1.0 : 0.0; break;
            case OP_AND: v = (reg[ins.y] != 0.0 && reg[ins.z] != 0.0) ? 1.0 : 0.0; break;
            case OP_OR: v = (reg[ins.y] != 0.0 || reg[ins.z] != 0.0) ? 1.0 : 0.0; break;
            case OP_NEG: v = -reg[ins.y]; break;
            case OP_ABS: v = fabs(reg[ins.y]); break;
            case OP_SQRT: v = sqrt(reg[ins.y]); break;
            case OP_EXP: v = exp(reg[ins.y]); break;
            case OP_LOG: v = log(reg[ins.y]); break;
            case OP_LOG10: v = log10(reg[ins.y]); break;
            case OP_SIN: v = sin(reg[ins.y]); break;
            case OP_COS: v = cos(reg[ins.y]); break;
            case OP_TAN: v = tan(reg[ins.y]); break;
            case OP_TANH: v = tanh(reg[ins.y]); break;
            case OP_ARCTAN: v = atan(reg[ins.y]); break;
            case OP_NOT: v = reg[ins.y] == 0.0 ? 1.0 : 0.0; break;
            case OP_SQUARE: v = __dmul_rn(reg[ins.y], reg[ins.y]); break;
            case OP_RECIP: v = __ddiv_rn(1.0, reg[ins.y]); break;
            case OP_SIGN: { double a = reg[ins.y]; v = isnan(a) ? a : (a > 0.0 ? 1.0 : (a < 0.0 ? -1.0 : 0.0)); } break;
            case OP_CBRT: v = cbrt(reg[ins.y]); break;
            case OP_LOG2: v = log2(reg[ins.y]); break;
            case OP_EXP2: v = exp2(reg[ins.y]); break;
            case OP_FLOOR: v = floor(reg[ins.y]); break;
            case OP_CEIL: v = ceil(reg[ins.y]); break;
            case OP_ISNAN: v = isnan(reg[ins.y]) ? 1.0 : 0.0; break;
            case OP_WHERE: v = reg[ins.y] != 0.0 ? reg[ins.z] : reg[ins.w]; break;
            case OP_INTERP: v = tc_interp(tb, ins.y, reg[ins.z]); break;
            case OP_ARCTAN2: v = atan2(reg[ins.y], reg[ins.z]); break;
            case OP_FMOD: v = fmod(reg[ins.y], reg[ins.z]); break;
            default: v = __longlong_as_double(0x7ff8000000000000LL); break;
        }
        reg[k] = v;
    }
    return reg[prog_dir[4 * prog + 3]];
}
