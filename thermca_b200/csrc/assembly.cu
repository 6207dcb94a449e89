// P1 (linear tetrahedra) operator build on the device - SURVEY.md §8f row 2.
//
// The reference assembles through scikit-fem (thermca/fem/fe_system.py:322-400):
//   Lx = asm(laplace),  Mx_diag = rowsum(asm(mass)),  Qs[:, s] = asm(unit_load on surface s)
// which for P1 elements is, per tetrahedron with volume V and barycentric gradients g_i,
//   Lx_ij += V g_i.g_j,   Mx_i += V/4   (row sum of V/20 (1 + delta_ij)),
// and per surface triangle with area a:  Qs_i += a/3.
// Here one thread per element writes its contributions as (key, value) pairs; the host side
// (thermca_b200/assembly.py) orders them with a stable radix sort and the segment kernel below
// sums every matrix entry in element order, so the result is run-to-run identical (no atomics).
// The CSR result is then laid out as the SELL-32 operator A = M^-1 L the solver kernels consume,
// without a round trip through host memory.
#include "common.cuh"
#include "../../include/thermca_b200.h"

__global__ void tc_p1_tet_kernel(long long n_points, long long n_tets, const double* __restrict__ pts,
                                 const int* __restrict__ tets, long long* __restrict__ keys,
                                 double* __restrict__ vals, int* __restrict__ mkeys, double* __restrict__ mvals) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_tets;
         e += (long long)gridDim.x * blockDim.x) {
        int v[4];
        double p[4][3];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[i] = tets[4 * e + i];
#pragma unroll
            for (int k = 0; k < 3; ++k) p[i][k] = pts[3 * (long long)v[i] + k];
        }
        double a[3], b[3], c[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { a[k] = p[1][k] - p[0][k]; b[k] = p[2][k] - p[0][k]; c[k] = p[3][k] - p[0][k]; }
        // rows of the inverse edge matrix = gradients of lambda_1..3 (times det)
        double g[4][3];
        g[1][0] = b[1] * c[2] - b[2] * c[1]; g[1][1] = b[2] * c[0] - b[0] * c[2]; g[1][2] = b[0] * c[1] - b[1] * c[0];
        g[2][0] = c[1] * a[2] - c[2] * a[1]; g[2][1] = c[2] * a[0] - c[0] * a[2]; g[2][2] = c[0] * a[1] - c[1] * a[0];
        g[3][0] = a[1] * b[2] - a[2] * b[1]; g[3][1] = a[2] * b[0] - a[0] * b[2]; g[3][2] = a[0] * b[1] - a[1] * b[0];
        const double det = a[0] * g[1][0] + a[1] * g[1][1] + a[2] * g[1][2];
        const double inv = 1.0 / det;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            g[1][k] *= inv; g[2][k] *= inv; g[3][k] *= inv;
            g[0][k] = -(g[1][k] + g[2][k] + g[3][k]);
        }
        const double vol = fabs(det) / 6.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                keys[16 * e + 4 * i + j] = (long long)v[i] * n_points + v[j];
                vals[16 * e + 4 * i + j] = vol * (g[i][0] * g[j][0] + g[i][1] * g[j][1] + g[i][2] * g[j][2]);
            }
            mkeys[4 * e + i] = v[i];
            mvals[4 * e + i] = vol / 4.0;
        }
    }
}

__global__ void tc_p1_tri_kernel(long long n_tris, const double* __restrict__ pts, const int* __restrict__ tris,
                                 int* __restrict__ keys, double* __restrict__ vals) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_tris;
         e += (long long)gridDim.x * blockDim.x) {
        double p[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) p[i][k] = pts[3 * (long long)tris[3 * e + i] + k];
        double a[3], b[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { a[k] = p[1][k] - p[0][k]; b[k] = p[2][k] - p[0][k]; }
        const double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
        const double area = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
#pragma unroll
        for (int i = 0; i < 3; ++i) { keys[3 * e + i] = tris[3 * e + i]; vals[3 * e + i] = area / 3.0; }
    }
}

// out[s] = vals[start[s]] + ... + vals[start[s+1]-1], summed left to right
__global__ void tc_segment_sum_kernel(long long nseg, const long long* __restrict__ start,
                                      const double* __restrict__ vals, double* __restrict__ out) {
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < nseg;
         s += (long long)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (long long i = start[s]; i < start[s + 1]; ++i) acc += vals[i];
        out[s] = acc;
    }
}

// CSR (sorted columns) -> SELL-32, value = scale * row_scale[row] * data; padding (0, own row)
__global__ void tc_csr_to_sell_kernel(int n, int nslices, const long long* __restrict__ indptr,
                                      const int* __restrict__ indices, const double* __restrict__ data,
                                      const double* __restrict__ row_scale, double scale,
                                      const long long* __restrict__ slice_ptr, int* __restrict__ col,
                                      short* __restrict__ col16, double* __restrict__ val,
                                      double* __restrict__ diag_out) {
    const long long total = (long long)nslices * 32;
    for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < total;
         row += (long long)gridDim.x * blockDim.x) {
        const int sl = (int)(row >> 5), lane = (int)(row & 31);
        const long long base = slice_ptr[sl];
        const int width = (int)((slice_ptr[sl + 1] - base) >> 5);
        const bool in = row < n;
        const long long lo = in ? indptr[row] : 0, hi = in ? indptr[row + 1] : 0;
        const double rs = in ? scale * (row_scale ? row_scale[row] : 1.0) : 0.0;
        const int own = in ? (int)row : n - 1;
        double dg = 0.0;
        for (int k = 0; k < width; ++k) {
            int c = own;
            double v = 0.0;
            if (lo + k < hi) {
                c = indices[lo + k];
                v = rs * data[lo + k];
                if (c == row) dg = v;
            }
            const long long p = base + 32LL * k + lane;
            col[p] = c;
            val[p] = v;
            if (col16) col16[p] = (short)(c - (int)row);
        }
        if (in && diag_out) diag_out[row] = dg;
    }
}

static int grid_1d(long long n, int threads) {
    long long g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > 148LL * 32) g = 148LL * 32;
    return (int)g;
}

extern "C" int tc_p1_tet_contributions(void* cuda_stream, int64_t n_points, int64_t n_tets, const double* points,
                                       const int32_t* tets, int64_t* keys, double* vals, int32_t* mkeys,
                                       double* mvals) {
    TC_CHECK(n_points > 0 && n_points < (1LL << 31), "vertex count must fit 31 bits");
    if (n_tets == 0) return 0;
    tc_p1_tet_kernel<<<grid_1d(n_tets, 256), 256, 0, (cudaStream_t)cuda_stream>>>(
        n_points, n_tets, points, tets, (long long*)keys, vals, mkeys, mvals);
    TC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tc_p1_tri_loads(void* cuda_stream, int64_t n_tris, const double* points, const int32_t* tris,
                               int32_t* keys, double* vals) {
    if (n_tris == 0) return 0;
    tc_p1_tri_kernel<<<grid_1d(n_tris, 256), 256, 0, (cudaStream_t)cuda_stream>>>(n_tris, points, tris, keys, vals);
    TC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tc_segment_sum(void* cuda_stream, int64_t nseg, const int64_t* start, const double* vals, double* out) {
    if (nseg == 0) return 0;
    tc_segment_sum_kernel<<<grid_1d(nseg, 256), 256, 0, (cudaStream_t)cuda_stream>>>(nseg, (const long long*)start,
                                                                                       vals, out);
    TC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int tc_csr_to_sell(void* cuda_stream, int32_t n, int32_t nslices, const int64_t* indptr,
                              const int32_t* indices, const double* data, const double* row_scale, double scale,
                              const int64_t* slice_ptr, int32_t* col, int16_t* col16, double* val,
                              double* diag_out) {
    TC_CHECK(n > 0 && nslices == (n + 31) / 32, "slice count does not match the row count");
    tc_csr_to_sell_kernel<<<grid_1d((long long)nslices * 32, 256), 256, 0, (cudaStream_t)cuda_stream>>>(
        n, nslices, (const long long*)indptr, indices, data, row_scale, scale, (const long long*)slice_ptr, col, col16,
        val, diag_out);
    TC_CUDA(cudaGetLastError());
    return 0;
}
