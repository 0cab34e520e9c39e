// kernels_geometry.cu -- geometry_preprocessing_new on the device.
//
// Replaces geometry_preprocessing_new (MP/Geometry_preprocessing.F90:9-512): node classification, the boundary-node
// lists of the colour-gradient chain and the wall normals of the wetting model.  The reference runs it serially on
// rank 0 over the WHOLE lattice with two FP64 copies and ten ghost layers, and flags itself as too slow for large
// domains (:4-8).  Everything in it is a local stencil (18-neighbour classification, 4 x 27-point smoothing, radius-2
// ISO8 gradient), so one GPU processes just the z window around its slab:
//
//   k_geo_extend    wall array of the window -> extended int8 array (replicate / periodic ghost layers, :56-108)
//   k_geo_classify  1 -> 2 (solid boundary), 0 -> -1 (fluid boundary)                         (:121-143)
//   k_geo_smooth    one pass of the 27-point smoothing, ping-pong between two FP64 arrays      (:145-169)
//   k_geo_rows      per lattice row: number of listed solid / fluid boundary nodes             (:171-185)
//   k_geo_emit      ordered (k outer, i inner) emission of both lists, neighbour lists and la_weight (:195-225)
//   k_geo_normals   ISO8 gradient of the smoothed field at the fluid boundary nodes            (:227-383)
//
// Integer results are bit-exact by construction.  The FP64 sums use __dmul_rn / __dadd_rn in the reference's term
// order, so ptxas cannot contract them: the normals are bit-identical to an un-contracted CPU build in BOTH library
// builds (tests/test_geometry_gpu.py).
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "mflbm_internal.cuh"

namespace mflbm {

struct Geo {
    int nxG, nyG, nzG;
    int gl;            // ghost layers of the extended arrays (6 + overlap_phi = 10, :41-42)
    int ek0, ek1;      // global z range of the extended arrays
    int wk0, wk1;      // global planes held in the caller's window
    int iper, jper, kper;
    long long ex, ey, ez;
    __host__ __device__ __forceinline__ size_t E(int i, int j, int k) const {
        return (size_t)(i + gl - 1) + (size_t)ex * ((size_t)(j + gl - 1) + (size_t)ey * (size_t)(k - ek0));
    }
};

// source coordinate of a ghost cell: replicate the boundary plane, or wrap when periodic
__device__ __forceinline__ int geo_map(int i, int n, int per) {
    if (i >= 1 && i <= n) return i;
    if (!per) return i < 1 ? 1 : n;
    return i < 1 ? i + n : i - n;
}

__global__ void __launch_bounds__(128) k_geo_extend(const Geo g, const int8_t *__restrict__ win, int8_t *__restrict__ wt,
                                                    double *__restrict__ ws1, double *__restrict__ ws2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1 - g.gl;
    const int j = (int)blockIdx.y + 1 - g.gl;
    const int k = (int)blockIdx.z + g.ek0;
    if (i > g.nxG + g.gl) return;
    const int si = geo_map(i, g.nxG, g.iper), sj = geo_map(j, g.nyG, g.jper);
    // z: planes of the window map to themselves; beyond it only the lattice ends are extended (the caller's window
    // already carries the neighbour slabs' / the wrapped planes everywhere else)
    int sk = k;
    if (k < g.wk0) sk = g.kper ? k + g.nzG : 1;
    else if (k > g.wk1) sk = g.kper ? k - g.nzG : g.nzG;
    const int8_t v = win[(size_t)(si - 1) + (size_t)g.nxG * ((size_t)(sj - 1) + (size_t)g.nyG * (size_t)(sk - g.wk0))];
    const size_t c = g.E(i, j, k);
    wt[c] = v;
    ws1[c] = (double)v;
    ws2[c] = (double)v;
}

__global__ void __launch_bounds__(128) k_geo_classify(const Geo g, const int8_t *__restrict__ wt, int8_t *__restrict__ cls) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 2 - g.gl;
    const int j = (int)blockIdx.y + 2 - g.gl;
    const int k = (int)blockIdx.z + g.ek0 + 1;
    if (i > g.nxG + g.gl - 1) return;
    const size_t c = g.E(i, j, k);
    const int8_t w = wt[c];
    int8_t out = w;
    if (w == 1) {
#pragma unroll
        for (int n = 1; n <= 18; n++)
            if (wt[g.E(i + EX(n), j + EY(n), k + EZ(n))] <= 0) out = 2;
    } else if (w == 0) {
#pragma unroll
        for (int n = 1; n <= 18; n++)
            if (wt[g.E(i + EX(n), j + EY(n), k + EZ(n))] >= 1) out = -1;
    }
    cls[c] = out;
}

// 27-point stencil in the reference's summation order (:150-163): centre, 6 faces, 12 edges, 8 corners
__host__ __device__ constexpr int GEO_SX(int n) {
    constexpr int t[27] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1};
    return t[n];
}
__host__ __device__ constexpr int GEO_SY(int n) {
    constexpr int t[27] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, -1, 1, -1, 1};
    return t[n];
}
__host__ __device__ constexpr int GEO_SZ(int n) {
    constexpr int t[27] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1, -1, 1, 1, -1, -1, 1, 1, -1};
    return t[n];
}
__host__ __device__ constexpr double GEO_SW(int n) {  // weight by squared distance: 8/27, 2/27, 1/54, 1/216 (:24)
    constexpr double we[4] = {8.0 / 27.0, 2.0 / 27.0, 1.0 / 54.0, 1.0 / 216.0};
    return we[GEO_SX(n) * GEO_SX(n) + GEO_SY(n) * GEO_SY(n) + GEO_SZ(n) * GEO_SZ(n)];
}

__global__ void __launch_bounds__(128) k_geo_smooth(const Geo g, const double *__restrict__ src, double *__restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 2 - g.gl;
    const int j = (int)blockIdx.y + 2 - g.gl;
    const int k = (int)blockIdx.z + g.ek0 + 1;
    if (i > g.nxG + g.gl - 1) return;
    const size_t c = g.E(i, j, k);
    const long long sy = g.ex, sz = g.ex * g.ey;
    double acc = 0.0;
#pragma unroll
    for (int n = 0; n < 27; n++)
        acc = __dadd_rn(acc, __dmul_rn(src[c + GEO_SX(n) + sy * GEO_SY(n) + sz * GEO_SZ(n)], GEO_SW(n)));
    dst[c] = acc;
}

// ---- ordered list construction: one warp per lattice row (j, k) of the scan range 1-4 .. n+4 -------------------
struct GeoScan {
    int ks0, ks1;      // global z range scanned for this slab
    int ophi;          // 4
    int idz, nz, nx, ny;
    int rows_per_plane;  // nyG + 2*ophi
};

__device__ __forceinline__ bool keep_solid(const GeoScan &s, int i, int j, int kl) {
    return i >= -2 && i <= s.nx + 3 && j >= -2 && j <= s.ny + 3 && kl >= -2 && kl <= s.nz + 3;
}
__device__ __forceinline__ bool keep_fluid(const GeoScan &s, int i, int j, int kl) {
    return i >= -1 && i <= s.nx + 2 && j >= -1 && j <= s.ny + 2 && kl >= -1 && kl <= s.nz + 2;
}

// counts[4*row + {0,1,2,3}] = kept solid, kept fluid, all solid, all fluid boundary nodes of the row
__global__ void __launch_bounds__(128) k_geo_rows(const Geo g, const GeoScan s, const int8_t *__restrict__ cls, int *__restrict__ counts, int nrows) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int lane = threadIdx.x & 31;
    const int k = s.ks0 + row / s.rows_per_plane, j = 1 - s.ophi + row % s.rows_per_plane;
    const int kl = k - s.idz * s.nz;
    int c[4] = {0, 0, 0, 0};
    for (int i = 1 - s.ophi + lane; i <= g.nxG + s.ophi; i += 32) {
        const int8_t t = cls[g.E(i, j, k)];
        if (t == 2) { c[2]++; c[0] += keep_solid(s, i, j, kl); }
        if (t == -1) { c[3]++; c[1] += keep_fluid(s, i, j, kl); }
    }
#pragma unroll
    for (int m = 0; m < 4; m++) {
        int v = c[m];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) counts[4 * row + m] = v;
    }
}

// offs[2*row + {0,1}] = index of the row's first kept solid / fluid node in the output lists
__global__ void __launch_bounds__(128) k_geo_emit(const Geo g, const GeoScan s, const int8_t *__restrict__ cls, const int *__restrict__ offs,
                                                  int nrows, mflbm_solid_node *__restrict__ solid, mflbm_fluid_node *__restrict__ fluid, double theta) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int lane = threadIdx.x & 31;
    const int k = s.ks0 + row / s.rows_per_plane, j = 1 - s.ophi + row % s.rows_per_plane;
    const int kl = k - s.idz * s.nz;
    int os = offs[2 * row], of = offs[2 * row + 1];
    const unsigned below = (1u << lane) - 1u;
    for (int i0 = 1 - s.ophi; i0 <= g.nxG + s.ophi; i0 += 32) {
        const int i = i0 + lane;
        int8_t t = 0;
        if (i <= g.nxG + s.ophi) t = cls[g.E(i, j, k)];
        const bool is_s = t == 2 && keep_solid(s, i, j, kl), is_f = t == -1 && keep_fluid(s, i, j, kl);
        const unsigned ms = __ballot_sync(0xffffffffu, is_s), mf = __ballot_sync(0xffffffffu, is_f);
        if (is_s) {
            mflbm_solid_node sn;
            sn.ix = i; sn.iy = j; sn.iz = kl;
            int cnt = 0;
            double law = 0.0;
#pragma unroll
            for (int n = 0; n < 18; n++) sn.neighbor_list[n] = 0;
#pragma unroll
            for (int n = 1; n <= 18; n++)
                if (cls[g.E(i + EX(n), j + EY(n), k + EZ(n))] <= 0) {
                    law = __dadd_rn(law, n <= 6 ? 1.0 / 18.0 : 1.0 / 36.0);
                    sn.neighbor_list[cnt++] = n;
                }
            sn.i_fluid_num = cnt;
            sn.la_weight = law;
            solid[os + __popc(ms & below)] = sn;
        }
        if (is_f) {
            mflbm_fluid_node fn;
            fn.ix = i; fn.iy = j; fn.iz = kl; fn.pad_ = 0;
            fn.nwx = fn.nwy = fn.nwz = 0.0;
            fn.theta = theta;
            fluid[of + __popc(mf & below)] = fn;
        }
        os += __popc(ms);
        of += __popc(mf);
    }
}

// ISO8 stencil of the wall normal grouped by weight (:234-377): each term is ws(x+o) - ws(x-o), accumulated left to right
struct Off3 {
    signed char a, b, c;
};
__device__ __constant__ int ISO8_CNT[7] = {1, 4, 4, 1, 8, 12, 4};
__device__ __constant__ double ISO8_W[7] = {4.0 / 45.0, 1.0 / 21.0, 2.0 / 105.0, 5.0 / 504.0, 1.0 / 315.0, 1.0 / 630.0, 1.0 / 5040.0};
__device__ __constant__ Off3 ISO8_T[3][34] = {
    {{1,0,0},{1,1,0},{1,-1,0},{1,0,1},{1,0,-1},{1,1,1},{1,1,-1},{1,-1,1},{1,-1,-1},{2,0,0},
     {2,1,0},{2,-1,0},{2,0,1},{2,0,-1},{1,2,0},{1,-2,0},{1,0,2},{1,0,-2},
     {2,1,1},{2,1,-1},{2,-1,1},{2,-1,-1},{1,2,1},{1,2,-1},{1,-2,1},{1,-2,-1},{1,1,2},{1,1,-2},{1,-1,2},{1,-1,-2},
     {2,2,0},{2,-2,0},{2,0,2},{2,0,-2}},
    {{0,1,0},{1,1,0},{-1,1,0},{0,1,1},{0,1,-1},{1,1,1},{1,1,-1},{-1,1,-1},{-1,1,1},{0,2,0},
     {2,1,0},{-2,1,0},{0,2,1},{0,2,-1},{1,2,0},{-1,2,0},{0,1,2},{0,1,-2},
     {2,1,1},{2,1,-1},{-2,1,1},{-2,1,-1},{1,2,1},{1,2,-1},{-1,2,1},{-1,2,-1},{1,1,2},{1,1,-2},{-1,1,2},{-1,1,-2},
     {2,2,0},{-2,2,0},{0,2,2},{0,2,-2}},
    {{0,0,1},{0,1,1},{0,-1,1},{1,0,1},{-1,0,1},{1,1,1},{1,-1,1},{-1,1,1},{-1,-1,1},{0,0,2},
     {0,1,2},{0,-1,2},{2,0,1},{-2,0,1},{0,2,1},{0,-2,1},{1,0,2},{-1,0,2},
     {2,1,1},{2,-1,1},{-2,1,1},{-2,-1,1},{1,2,1},{1,-2,1},{-1,2,1},{-1,-2,1},{1,1,2},{1,-1,2},{-1,1,2},{-1,-1,2},
     {0,2,2},{0,-2,2},{2,0,2},{-2,0,2}}};

__device__ __forceinline__ double iso8(const double *__restrict__ ws, size_t c, long long sy, long long sz, int axis) {
    double res = 0.0;
    int t = 0;
    for (int grp = 0; grp < 7; grp++) {
        double acc = 0.0;
        for (int m = 0; m < ISO8_CNT[grp]; m++, t++) {
            const Off3 o3 = ISO8_T[axis][t];
            const long long o = o3.a + sy * o3.b + sz * o3.c;
            if (m == 0) acc = __dsub_rn(ws[c + o], ws[c - o]);
            else {
                acc = __dadd_rn(acc, ws[c + o]);
                acc = __dsub_rn(acc, ws[c - o]);
            }
        }
        res = grp == 0 ? __dmul_rn(ISO8_W[0], acc) : __dadd_rn(res, __dmul_rn(ISO8_W[grp], acc));
    }
    return res;
}

__global__ void __launch_bounds__(128) k_geo_normals(const Geo g, const double *__restrict__ ws, mflbm_fluid_node *__restrict__ fluid, int n, int idz_nz) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    mflbm_fluid_node f = fluid[m];
    const size_t c = g.E(f.ix, f.iy, f.iz + idz_nz);
    const long long sy = g.ex, sz = g.ex * g.ey;
    const double nwx = iso8(ws, c, sy, sz, 0), nwy = iso8(ws, c, sy, sz, 1), nwz = iso8(ws, c, sy, sz, 2);
    const double s2 = __dadd_rn(__dadd_rn(__dmul_rn(nwx, nwx), __dmul_rn(nwy, nwy)), __dmul_rn(nwz, nwz));
    const double tmp = __ddiv_rn(1.0, __dadd_rn(__dsqrt_rn(s2), 1.110223025e-16));  // eps, MP/Module.F90:8
    f.nwx = __dmul_rn(nwx, tmp);
    f.nwy = __dmul_rn(nwy, tmp);
    f.nwz = __dmul_rn(nwz, tmp);
    fluid[m] = f;
}

}  // namespace mflbm

using namespace mflbm;

static thread_local std::string g_geo_err;
extern "C" const char *mflbm_geometry_last_error(void) { return g_geo_err.c_str(); }

#define GEO_CU(x)                                                                            \
    do {                                                                                     \
        cudaError_t e_ = (x);                                                                \
        if (e_ != cudaSuccess) {                                                             \
            g_geo_err = std::string(#x) + ": " + cudaGetErrorString(e_);                     \
            rc = MFLBM_ERR_CUDA;                                                             \
            goto done;                                                                       \
        }                                                                                    \
    } while (0)

extern "C" int mflbm_geometry_preprocess(const mflbm_geometry_config *cfg, const int8_t *walls_window, mflbm_solid_node **solid_out,
                                         int32_t *num_solid, mflbm_fluid_node **fluid_out, int32_t *num_fluid,
                                         int64_t *num_solid_scanned, int64_t *num_fluid_scanned) {
    if (!cfg || !walls_window || !solid_out || !fluid_out || !num_solid || !num_fluid) {
        g_geo_err = "null argument";
        return MFLBM_ERR_ARG;
    }
    if (cfg->struct_size != (int)sizeof(mflbm_geometry_config)) {
        g_geo_err = "mflbm_geometry_config.struct_size mismatch";
        return MFLBM_ERR_ARG;
    }
    const int nxG = cfg->nxGlobal, nyG = cfg->nyGlobal, nzG = cfg->nzGlobal;
    if (nxG < 1 || nyG < 1 || nzG < 1 || cfg->npz < 1 || nzG % cfg->npz || cfg->idz < 0 || cfg->idz >= cfg->npz || cfg->wk1 < cfg->wk0) {
        g_geo_err = "bad lattice / slab / window description";
        return MFLBM_ERR_ARG;
    }
    const int nz = nzG / cfg->npz, gl = 10, ophi = 4;
    const bool whole = cfg->wk0 == 1 && cfg->wk1 == nzG;
    Geo g;
    g.nxG = nxG; g.nyG = nyG; g.nzG = nzG; g.gl = gl;
    g.wk0 = cfg->wk0; g.wk1 = cfg->wk1;
    g.iper = cfg->iper; g.jper = cfg->jper; g.kper = cfg->kper;
    g.ek0 = cfg->wk0; g.ek1 = cfg->wk1;
    if (whole || cfg->kper == 0) {  // only the lattice ends are extended; elsewhere the window carries real planes
        if (cfg->wk0 == 1) g.ek0 = 1 - gl;
        if (cfg->wk1 == nzG) g.ek1 = nzG + gl;
    }
    g.ex = nxG + 2 * gl; g.ey = nyG + 2 * gl; g.ez = g.ek1 - g.ek0 + 1;
    GeoScan s;
    s.ophi = ophi; s.idz = cfg->idz; s.nz = nz; s.nx = nxG; s.ny = nyG;
    s.ks0 = whole ? 1 - ophi : std::max(cfg->idz * nz + 1 - 3, cfg->kper ? -(1 << 30) : 1 - ophi);
    s.ks1 = whole ? nzG + ophi : std::min(cfg->idz * nz + nz + 3, cfg->kper ? (1 << 30) : nzG + ophi);
    s.rows_per_plane = nyG + 2 * ophi;
    // the scan range needs the classification (radius 1) and the normals (smoothing radius 4 + ISO8 radius 2) around it
    if (s.ks0 - 1 < g.ek0 + 1 || s.ks1 + 1 > g.ek1 - 1 || s.ks0 - 2 - 4 < g.ek0 || s.ks1 + 2 + 4 > g.ek1) {
        g_geo_err = "wall window too small: it must reach the lattice end or extend >= 10 planes beyond the slab";
        return MFLBM_ERR_ARG;
    }
    const size_t ntot = (size_t)g.ex * g.ey * g.ez;
    const size_t nwin = (size_t)nxG * nyG * (size_t)(cfg->wk1 - cfg->wk0 + 1);
    const int nrows = (s.ks1 - s.ks0 + 1) * s.rows_per_plane;
    int rc = MFLBM_OK;
    int8_t *d_win = nullptr, *d_wt = nullptr, *d_cls = nullptr;
    double *d_a = nullptr, *d_b = nullptr;
    int *d_counts = nullptr, *d_offs = nullptr;
    mflbm_solid_node *d_solid = nullptr;
    mflbm_fluid_node *d_fluid = nullptr;
    std::vector<int> counts, offs;
    long long ns = 0, nf = 0, ns_all = 0, nf_all = 0;
    *solid_out = nullptr; *fluid_out = nullptr; *num_solid = 0; *num_fluid = 0;
    {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            g_geo_err = "no CUDA device: this library has no CPU fallback";
            return MFLBM_ERR_CUDA;
        }
    }
    if (cfg->device >= 0) GEO_CU(cudaSetDevice(cfg->device));
    GEO_CU(cudaMalloc((void **)&d_win, nwin));
    GEO_CU(cudaMalloc((void **)&d_wt, ntot));
    GEO_CU(cudaMalloc((void **)&d_cls, ntot));
    GEO_CU(cudaMalloc((void **)&d_a, ntot * sizeof(double)));
    GEO_CU(cudaMalloc((void **)&d_b, ntot * sizeof(double)));
    GEO_CU(cudaMalloc((void **)&d_counts, (size_t)nrows * 4 * sizeof(int)));
    GEO_CU(cudaMalloc((void **)&d_offs, (size_t)nrows * 2 * sizeof(int)));
    GEO_CU(cudaMemcpy(d_win, walls_window, nwin, cudaMemcpyHostToDevice));
    {
        const dim3 gext((unsigned)((g.ex + 127) / 128), (unsigned)g.ey, (unsigned)g.ez);
        k_geo_extend<<<gext, 128>>>(g, d_win, d_wt, d_a, d_b);
        GEO_CU(cudaMemcpy(d_cls, d_wt, ntot, cudaMemcpyDeviceToDevice));  // cells outside the classified range keep the wall value
        const dim3 gin((unsigned)((g.ex - 2 + 127) / 128), (unsigned)(g.ey - 2), (unsigned)(g.ez - 2));
        k_geo_classify<<<gin, 128>>>(g, d_wt, d_cls);
        for (int it = 0; it < 4; it++) {  // the reference copies ws2 back into ws1 after every pass; swapping is the same
            k_geo_smooth<<<gin, 128>>>(g, d_a, d_b);
            std::swap(d_a, d_b);
        }
        // d_a now holds the four-times smoothed field (= the reference's ws2)
        k_geo_rows<<<(nrows + 3) / 4, 128>>>(g, s, d_cls, d_counts, nrows);
        GEO_CU(cudaGetLastError());
    }
    counts.resize((size_t)nrows * 4);
    offs.resize((size_t)nrows * 2);
    GEO_CU(cudaMemcpy(counts.data(), d_counts, counts.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (int r = 0; r < nrows; r++) {
        offs[2 * r] = (int)ns;
        offs[2 * r + 1] = (int)nf;
        ns += counts[4 * r];
        nf += counts[4 * r + 1];
        ns_all += counts[4 * r + 2];
        nf_all += counts[4 * r + 3];
    }
    if (ns >= (1LL << 31) || nf >= (1LL << 31)) {
        g_geo_err = "boundary-node list exceeds int32";
        rc = MFLBM_ERR_ARG;
        goto done;
    }
    GEO_CU(cudaMemcpy(d_offs, offs.data(), offs.size() * sizeof(int), cudaMemcpyHostToDevice));
    GEO_CU(cudaMalloc((void **)&d_solid, (size_t)std::max<long long>(ns, 1) * sizeof(mflbm_solid_node)));
    GEO_CU(cudaMalloc((void **)&d_fluid, (size_t)std::max<long long>(nf, 1) * sizeof(mflbm_fluid_node)));
    k_geo_emit<<<(nrows + 3) / 4, 128>>>(g, s, d_cls, d_offs, nrows, d_solid, d_fluid, cfg->theta);
    if (nf > 0) k_geo_normals<<<(unsigned)((nf + 127) / 128), 128>>>(g, d_a, d_fluid, (int)nf, cfg->idz * nz);
    GEO_CU(cudaGetLastError());
    *solid_out = (mflbm_solid_node *)malloc((size_t)std::max<long long>(ns, 1) * sizeof(mflbm_solid_node));
    *fluid_out = (mflbm_fluid_node *)malloc((size_t)std::max<long long>(nf, 1) * sizeof(mflbm_fluid_node));
    if (!*solid_out || !*fluid_out) {
        g_geo_err = "out of host memory";
        rc = MFLBM_ERR_STATE;
        goto done;
    }
    if (ns > 0) GEO_CU(cudaMemcpy(*solid_out, d_solid, (size_t)ns * sizeof(mflbm_solid_node), cudaMemcpyDeviceToHost));
    if (nf > 0) GEO_CU(cudaMemcpy(*fluid_out, d_fluid, (size_t)nf * sizeof(mflbm_fluid_node), cudaMemcpyDeviceToHost));
    *num_solid = (int32_t)ns;
    *num_fluid = (int32_t)nf;
    if (num_solid_scanned) *num_solid_scanned = ns_all;
    if (num_fluid_scanned) *num_fluid_scanned = nf_all;
done:
    cudaFree(d_win); cudaFree(d_wt); cudaFree(d_cls); cudaFree(d_a); cudaFree(d_b);
    cudaFree(d_counts); cudaFree(d_offs); cudaFree(d_solid); cudaFree(d_fluid);
    if (rc != MFLBM_OK) {
        free(*solid_out); free(*fluid_out);
        *solid_out = nullptr; *fluid_out = nullptr;
    }
    return rc;
}

extern "C" void mflbm_geometry_free(void *list) { free(list); }
