// Device-side generator of the synthetic Kuhn-cuboid FE operator (BASELINE configs #2/#5).
//
// Builds, row by row and without atomics, what FESystem + FFEIO derive from a P1
// tetrahedral mesh (thermca/fem/fe_system.py:322-400, thermca/fem/ffe_io.py:85-117):
//   A_body = M^-1 L,  L_ab = condy * sum_tets vol * grad(lam_a).grad(lam_b),
//   M_a = vol_capy * sum_tets vol / 4  (row-sum lumped mass),
//   Qs[:, surf] = sum_facets area / 3  for the surfaces "bottom" (z = 0) and
//   "upper_faces" (the other five faces).
// Each thread owns one vertex: it walks the <= 24 incident tetrahedra of the 8
// surrounding cells (6 Kuhn tetrahedra per cell, path v000 -> +e_p0 -> +e_p1 -> +e_p2)
// and accumulates its own operator row (15 possible neighbours).  Output is SELL-32
// with a fixed width of 15; rows [row_begin, row_end) only (row partition for multi-GPU),
// column indices are global.  thermca_b200/synth.py is the host restatement this is
// checked against at small sizes.
#include "common.cuh"
#include "../../include/thermca_b200.h"

struct SynthGeo {
    int nx, ny, nz;
    double lx, ly, lz, jitter;
    double condy, vol_capy;
    long long row_begin, row_end;
};

__constant__ int c_perm[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};

__device__ __forceinline__ void synth_point(const SynthGeo& g, int i, int j, int k, double* p) {
    const double PI = 3.14159265358979323846;
    const double hx = g.lx / g.nx, hy = g.ly / g.ny, hz = g.lz / g.nz;
    const double x = hx * i, y = hy * j, z = hz * k;
    const double xi = x / g.lx, et = y / g.ly, ze = z / g.lz;
    p[0] = x + g.jitter * hx * sin(2 * PI * xi) * cos(3 * PI * et) * cos(2 * PI * ze);
    p[1] = y + g.jitter * hy * cos(2 * PI * xi + 0.3) * sin(2 * PI * et) * cos(PI * ze);
    p[2] = z + g.jitter * hz * cos(PI * xi) * cos(2 * PI * et + 0.7) * sin(2 * PI * ze);
    // keep the six faces exactly planar
    if (i == 0) p[0] = 0.0; if (i == g.nx) p[0] = g.lx;
    if (j == 0) p[1] = 0.0; if (j == g.ny) p[1] = g.ly;
    if (k == 0) p[2] = 0.0; if (k == g.nz) p[2] = g.lz;
}

__device__ __forceinline__ long long synth_vid(const SynthGeo& g, int i, int j, int k) {
    return ((long long)j * (g.nx + 1) + i) * (g.nz + 1) + k;
}

// slot of neighbour offset (dx,dy,dz): 15 admissible offsets, ordered by linear index offset
__device__ __forceinline__ int synth_slot(int dx, int dy, int dz, const int* order27) {
    return order27[(dy + 1) * 9 + (dx + 1) * 3 + (dz + 1)];
}

__device__ double tri_area(const double* a, const double* b, const double* c) {
    const double u0 = b[0] - a[0], u1 = b[1] - a[1], u2 = b[2] - a[2];
    const double v0 = c[0] - a[0], v1 = c[1] - a[1], v2 = c[2] - a[2];
    const double cx = u1 * v2 - u2 * v1, cy = u2 * v0 - u0 * v2, cz = u0 * v1 - u1 * v0;
    return 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
}

// load share of vertex (i,j,k) on the face {axis = fixed} : sum over adjacent triangles area/3
__device__ double face_load(const SynthGeo& g, int axis, int fixed, int i, int j, int k) {
    int v[3] = {i, j, k};
    int n[3] = {g.nx, g.ny, g.nz};
    const int fa = axis == 0 ? 1 : 0;            // free axes in increasing order
    const int fb = axis == 2 ? 1 : 2;
    double acc = 0.0;
    for (int da = -1; da <= 0; ++da)
        for (int db = -1; db <= 0; ++db) {
            const int a0 = v[fa] + da, b0 = v[fb] + db;
            if (a0 < 0 || b0 < 0 || a0 >= n[fa] || b0 >= n[fb]) continue;
            // square (a0,b0): triangles (v00,v10,v11) and (v00,v01,v11)
            int c00[3], c10[3], c01[3], c11[3];
            c00[axis] = c10[axis] = c01[axis] = c11[axis] = fixed;
            c00[fa] = a0; c00[fb] = b0; c10[fa] = a0 + 1; c10[fb] = b0;
            c01[fa] = a0; c01[fb] = b0 + 1; c11[fa] = a0 + 1; c11[fb] = b0 + 1;
            double p00[3], p10[3], p01[3], p11[3];
            synth_point(g, c00[0], c00[1], c00[2], p00);
            synth_point(g, c10[0], c10[1], c10[2], p10);
            synth_point(g, c01[0], c01[1], c01[2], p01);
            synth_point(g, c11[0], c11[1], c11[2], p11);
            const int la = -da, lb = -db;        // local position of the vertex in the square
            const bool is00 = la == 0 && lb == 0, is11 = la == 1 && lb == 1;
            const bool is10 = la == 1 && lb == 0, is01 = la == 0 && lb == 1;
            if (is00 || is10 || is11) acc += tri_area(p00, p10, p11) / 3.0;
            if (is00 || is01 || is11) acc += tri_area(p00, p01, p11) / 3.0;
        }
    return acc;
}

__global__ void tc_synth_cuboid_kernel(SynthGeo g, const int* __restrict__ order27,
                                       const long long* __restrict__ off15,
                                       int* __restrict__ col, short* __restrict__ col16, double* __restrict__ val,
                                       double* __restrict__ mass, double* __restrict__ adiag,
                                       double* __restrict__ q_bottom, double* __restrict__ q_upper) {
    const long long nloc = g.row_end - g.row_begin;
    const long long npad = ((nloc + 31) / 32) * 32;
    for (long long rl = blockIdx.x * (long long)blockDim.x + threadIdx.x; rl < npad;
         rl += (long long)gridDim.x * blockDim.x) {
        const long long row = g.row_begin + rl;
        if (rl >= nloc) {        // padding rows of the last slice: zeros pointing at an owned row
            const long long pb = (rl >> 5) * (32LL * 15) + (rl & 31);
            for (int s = 0; s < 15; ++s) {          // point at the last owned row (keeps col - row small)
                col[pb + 32LL * s] = (int)(g.row_end - 1); val[pb + 32LL * s] = 0.0;
                if (col16) col16[pb + 32LL * s] = (short)(g.row_end - 1 - row);
            }
            continue;
        }
        const int k = (int)(row % (g.nz + 1));
        const long long t = row / (g.nz + 1);
        const int i = (int)(t % (g.nx + 1));
        const int j = (int)(t / (g.nx + 1));
        double acc[15];
#pragma unroll
        for (int s = 0; s < 15; ++s) acc[s] = 0.0;
        double vol_sum = 0.0;
        for (int cz = 0; cz < 2; ++cz)
            for (int cy = 0; cy < 2; ++cy)
                for (int cx = 0; cx < 2; ++cx) {
                    const int oi = i - cx, oj = j - cy, ok = k - cz;        // cell origin
                    if (oi < 0 || oj < 0 || ok < 0 || oi >= g.nx || oj >= g.ny || ok >= g.nz) continue;
                    const int loc[3] = {cx, cy, cz};
                    for (int pm = 0; pm < 6; ++pm) {
                        // path points of this tetrahedron
                        int d[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
                        for (int s = 0; s < 3; ++s) {
                            d[s + 1][0] = d[s][0]; d[s + 1][1] = d[s][1]; d[s + 1][2] = d[s][2];
                            d[s + 1][c_perm[pm][s]] += 1;
                        }
                        int me = -1;
                        for (int s = 0; s < 4; ++s)
                            if (d[s][0] == loc[0] && d[s][1] == loc[1] && d[s][2] == loc[2]) me = s;
                        if (me < 0) continue;
                        double P[4][3];
                        for (int s = 0; s < 4; ++s) synth_point(g, oi + d[s][0], oj + d[s][1], ok + d[s][2], P[s]);
                        // gradients of barycentric coordinates: rows of inverse of edge matrix
                        double J[3][3];
                        for (int a = 0; a < 3; ++a)
                            for (int b = 0; b < 3; ++b) J[a][b] = P[a + 1][b] - P[0][b];
                        const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1])
                                         - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
                                         + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
                        const double vol = fabs(det) / 6.0;
                        double G[4][3];   // grad lam_s
                        const double id = 1.0 / det;
                        // inverse(J) columns are the gradients of lam_1..3
                        G[1][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * id;
                        G[1][1] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * id;
                        G[1][2] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * id;
                        G[2][0] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
                        G[2][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
                        G[2][2] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
                        G[3][0] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
                        G[3][1] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
                        G[3][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
                        for (int b = 0; b < 3; ++b) G[0][b] = -(G[1][b] + G[2][b] + G[3][b]);
                        vol_sum += vol;
                        for (int s = 0; s < 4; ++s) {
                            const double kij = vol * (G[me][0] * G[s][0] + G[me][1] * G[s][1] + G[me][2] * G[s][2]);
                            const int slot = synth_slot(d[s][0] - loc[0], d[s][1] - loc[1], d[s][2] - loc[2], order27);
                            acc[slot] += kij;
                        }
                    }
                }
        const double m = g.vol_capy * vol_sum / 4.0;
        const double minv = 1.0 / m;
        const long long base = (rl >> 5) * (32LL * 15) + (rl & 31);
        const int n[3] = {g.nx, g.ny, g.nz};
        (void)n;
#pragma unroll
        for (int s = 0; s < 15; ++s) {
            const long long c = row + off15[s];
            // neighbour exists iff its (i,j,k) is inside the grid
            const int dz = (int)off15[15 + 3 * s + 2], dx = (int)off15[15 + 3 * s + 0], dy = (int)off15[15 + 3 * s + 1];
            const bool inside = i + dx >= 0 && i + dx <= g.nx && j + dy >= 0 && j + dy <= g.ny &&
                                k + dz >= 0 && k + dz <= g.nz;
            const double a = inside ? minv * (g.condy * acc[s]) : 0.0;
            col[base + 32LL * s] = inside ? (int)c : (int)row;
            if (col16) col16[base + 32LL * s] = inside ? (short)off15[s] : (short)0;
            val[base + 32LL * s] = a;
            if (dx == 0 && dy == 0 && dz == 0) adiag[rl] = a;
        }
        mass[rl] = m;
        double qb = 0.0, qu = 0.0;
        if (k == 0) qb += face_load(g, 2, 0, i, j, k);
        if (k == g.nz) qu += face_load(g, 2, g.nz, i, j, k);
        if (i == 0) qu += face_load(g, 0, 0, i, j, k);
        if (i == g.nx) qu += face_load(g, 0, g.nx, i, j, k);
        if (j == 0) qu += face_load(g, 1, 0, i, j, k);
        if (j == g.ny) qu += face_load(g, 1, g.ny, i, j, k);
        q_bottom[rl] = qb;
        q_upper[rl] = qu;
    }
}

__global__ void tc_fill_slice_ptr_kernel(long long nslices, long long* slice_ptr) {
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s <= nslices;
         s += (long long)gridDim.x * blockDim.x) slice_ptr[s] = s * 32LL * 15;
}

extern "C" int tc_synth_cuboid_counts(int32_t nx, int32_t ny, int32_t nz, int64_t* n_rows, int64_t* n_slices,
                                      int64_t* sell_entries) {
    const int64_t n = (int64_t)(nx + 1) * (ny + 1) * (nz + 1);
    *n_rows = n;
    *n_slices = (n + 31) / 32;
    *sell_entries = *n_slices * 32 * 15;
    return 0;
}

extern "C" int tc_synth_cuboid(void* cuda_stream, int32_t nx, int32_t ny, int32_t nz, double lx, double ly,
                               double lz, double jitter, double condy, double vol_capy, int64_t row_begin,
                               int64_t row_end, int64_t* slice_ptr, int32_t* col, int16_t* col16, double* val,
                               double* mass, double* adiag, double* q_bottom, double* q_upper) {
    cudaStream_t st = (cudaStream_t)cuda_stream;
    // the 15 admissible neighbour offsets, ordered by linear index offset (ascending column)
    struct Off { long long lin; int dx, dy, dz; };
    Off offs[15];
    int cnt = 0;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx)
            for (int dz = -1; dz <= 1; ++dz) {
                const bool pos = dx >= 0 && dy >= 0 && dz >= 0, neg = dx <= 0 && dy <= 0 && dz <= 0;
                if (!(pos || neg)) continue;
                offs[cnt++] = Off{((long long)dy * (nx + 1) + dx) * (nz + 1) + dz, dx, dy, dz};
            }
    TC_CHECK(cnt == 15, "offset table");
    for (int a = 0; a < 15; ++a)
        for (int b = a + 1; b < 15; ++b)
            if (offs[b].lin < offs[a].lin) { Off t = offs[a]; offs[a] = offs[b]; offs[b] = t; }
    int order27[27];
    long long off15[15 + 45];
    for (int q = 0; q < 27; ++q) order27[q] = -1;
    for (int s = 0; s < 15; ++s) {
        order27[(offs[s].dy + 1) * 9 + (offs[s].dx + 1) * 3 + (offs[s].dz + 1)] = s;
        off15[s] = offs[s].lin;
        off15[15 + 3 * s + 0] = offs[s].dx; off15[15 + 3 * s + 1] = offs[s].dy; off15[15 + 3 * s + 2] = offs[s].dz;
    }
    int* d_order = nullptr; long long* d_off = nullptr;
    TC_CUDA(cudaMalloc(&d_order, sizeof order27));
    TC_CUDA(cudaMalloc(&d_off, sizeof off15));
    TC_CUDA(cudaMemcpyAsync(d_order, order27, sizeof order27, cudaMemcpyHostToDevice, st));
    TC_CUDA(cudaMemcpyAsync(d_off, off15, sizeof off15, cudaMemcpyHostToDevice, st));
    SynthGeo g{nx, ny, nz, lx, ly, lz, jitter, condy, vol_capy, row_begin, row_end};
    const long long nloc = row_end - row_begin;
    const long long nsl = (nloc + 31) / 32;
    tc_fill_slice_ptr_kernel<<<256, 256, 0, st>>>(nsl, (long long*)slice_ptr);
    long long gb = (nsl * 32 + 127) / 128;
    if (gb > 148 * 16) gb = 148 * 16;
    if (gb < 1) gb = 1;
    // the band of the Kuhn cuboid is one y-layer: relative columns fit 16 bits iff that is < 2^15
    const long long band = (long long)(nx + 1) * (nz + 1) + (nz + 1) + 1;
    if (band >= 32768) col16 = nullptr;
    tc_synth_cuboid_kernel<<<(int)gb, 128, 0, st>>>(g, d_order, d_off, col, (short*)col16, val, mass, adiag, q_bottom, q_upper);
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_order); cudaFree(d_off);
    return 0;
}
