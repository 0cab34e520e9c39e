// kernels_monitor.cu -- macroscopic variables and monitor reductions (warp-shuffle + one shared-memory
// stage per block; one block per z-slice so the per-slice profiles of the reference fall out directly).
//
// Replaces compute_macro_vars (MP/Misc.F90:372-430, SP/Misc.F90:368-423), the device parts of monitor
// (MP/Monitor.F90:27-85, SP/Monitor.F90:18-59), cal_saturation (MP/Monitor.F90:527-538),
// monitor_breakthrough (:483-495), monitor_multiphase_steady_phasefield (:303-334) and
// monitor_multiphase_steady_capillarypressure (:383-423).
#include "gradient.cuh"

namespace mflbm {

template <bool MP, bool SPARSE>
__global__ void __launch_bounds__(128) k_macro(const Dev P) {
    int c, cl;
    if (SPARSE) {  // fluid nodes only; u,v,w,rho are zero at solid nodes (they are zero-initialised and never written)
        const int n = blockIdx.x * blockDim.x + threadIdx.x;
        if (n >= P.nA) return;
        c = P.cellA[n];
        cl = n;
    } else {
        const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
        if (i > P.g.nx) return;
        c = P.g.cell(i, j, k);
        cl = c;
    }
    const int wi = SPARSE ? 0 : P.walls[c];
    double ft[19];
#pragma unroll
    for (int q = 0; q < 19; q++) ft[q] = MP ? P.f[q][cl] + P.gg[q][cl] : P.f[q][cl];
    P.rho[c] = (ft[0] + ft[1] + ft[2] + ft[3] + ft[4] + ft[5] + ft[6] + ft[7] + ft[8] + ft[9] + ft[10] + ft[11] + ft[12] + ft[13] +
                ft[14] + ft[15] + ft[16] + ft[17] + ft[18]) * (1 - wi);
    double fx = 0.0, fy = 0.0, fz = P.force_Z;
    if (MP) {
        const double tmp = 0.5 * P.gamma * (SPARSE ? curvature_at(P, c) : P.curv[c]) * P.c_norm[c];
        fx = tmp * P.cn_x[c];
        fy = tmp * P.cn_y[c];
        fz = tmp * P.cn_z[c] + P.force_Z;
    }
    P.u[c] = (ft[1] - ft[2] + ft[7] - ft[8] + ft[9] - ft[10] + ft[11] - ft[12] + ft[13] - ft[14] - 0.5 * fx) * (1 - wi);
    P.v[c] = (ft[3] - ft[4] + ft[7] + ft[8] - ft[9] - ft[10] + ft[15] - ft[16] + ft[17] - ft[18] - 0.5 * fy) * (1 - wi);
    P.w[c] = (ft[5] - ft[6] + ft[11] + ft[12] - ft[13] - ft[14] + ft[15] + ft[16] - ft[17] - ft[18] - 0.5 * fz) * (1 - wi);
    if (MP && !SPARSE) P.phi[c] = 0.0 * wi + P.phi[c] * (1 - wi);
}

// phi <- 0 at solid nodes of 1..n (last statement of compute_macro_vars, MP/Misc.F90:424) for the sparse layout
__global__ void __launch_bounds__(128) k_phi_zero_walls(const Dev P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i > P.g.nx) return;
    const int c = P.g.cell(i, j, k);
    const int wi = P.walls[c];
    if (wi) P.phi[c] = 0.0 * wi + P.phi[c] * (1 - wi);
}

void launch_macro(mflbm_ctx *c, cudaStream_t st) {
    const Dev &P = c->d;
    dim3 grid((P.g.nx + 127) / 128, P.g.ny, P.g.nz);
    if (P.sparse) {
        const int nb = (P.nA + 127) / 128;
        if (nb > 0) {
            if (P.multiphase) k_macro<true, true><<<nb, 128, 0, st>>>(P);
            else k_macro<false, true><<<nb, 128, 0, st>>>(P);
            c->launches++;
        }
        if (P.multiphase) {
            k_phi_zero_walls<<<grid, 128, 0, st>>>(P);
            c->launches++;
        }
        return;
    }
    if (P.multiphase) k_macro<true, false><<<grid, 128, 0, st>>>(P);
    else k_macro<false, false><<<grid, 128, 0, st>>>(P);
    c->launches++;
}

// ---- block reduction of NV values (sum or max per slot) ----
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], const bool (&is_max)[NV], double *out, int stride, int slot0) {
    __shared__ double sm[NV][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int n = 0; n < NV; n++) {
        double x = v[n];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_down_sync(0xffffffffu, x, o);
            x = is_max[n] ? fmax(x, y) : x + y;
        }
        if (lane == 0) sm[n][wid] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        const int n = threadIdx.x;
        double x = sm[n][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) x = is_max[n] ? fmax(x, sm[n][w]) : x + sm[n][w];
        out[(slot0 + n) * stride + blockIdx.x] = x;
    }
}

// out layout: 10 rows of nz: fl1 fl2 vol1 vol2 mass1 mass2 pre umax usq1 usq2 (singlephase: fl, -, -, -, -, -, pre, umax)
template <bool MP>
__global__ void __launch_bounds__(256) k_monitor(const Dev P, double *out) {
    const int k = blockIdx.x + 1;
    const int nx = P.g.nx, ny = P.g.ny, nz = P.g.nz;
    double v[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int n = threadIdx.x; n < nx * ny; n += blockDim.x) {
        const int j = n / nx + 1, i = n - (j - 1) * nx + 1;
        const int c = P.g.cell(i, j, k);
        const double uu = P.u[c], vv = P.v[c], ww = P.w[c], rho = P.rho[c];
        if (MP) {
            const int wi = P.walls[c];
            const double ph = P.phi[c];
            const double temp = (uu * uu + vv * vv + ww * ww) * (1 - wi);
            v[7] = fmax(v[7], temp);
            if (ph > 0.999) v[8] += temp;
            else if (ph < -0.999) v[9] += temp;
            v[2] += 0.5 * (1.0 + ph) * (1 - wi);
            v[3] += 0.5 * (1.0 - ph) * (1 - wi);
            v[4] += rho * 0.5 * (1.0 + ph) * (1 - wi);
            v[5] += rho * 0.5 * (1.0 - ph) * (1 - wi);
            v[0] += ww * 0.5 * (1.0 + ph) * (1 - wi);
            v[1] += ww * 0.5 * (1.0 - ph) * (1 - wi);
            v[6] += rho * (1 - wi);
        } else {  // SP/Monitor.F90:28-59: no wall factor (compute_macro_vars already zeroed walls)
            v[7] = fmax(v[7], uu * uu + vv * vv + ww * ww);
            v[0] += ww;
            v[6] += rho;
        }
    }
    const bool is_max[10] = {false, false, false, false, false, false, false, true, false, false};
    block_reduce<10>(v, is_max, out, nz, 0);
}

void launch_monitor(mflbm_ctx *c, cudaStream_t st, double *out) {
    const Dev &P = c->d;
    if (P.multiphase) k_monitor<true><<<P.g.nz, 256, 0, st>>>(P, out);
    else k_monitor<false><<<P.g.nz, 256, 0, st>>>(P, out);
    c->launches++;
}

// cal_saturation: out rows v1, v2, each nz * MFLBM_SAT_SEG partial sums (slice k, segment s at [row][k - 1 + nz * s]); the
// host adds them up.  One warp per row of the slice (coalesced, no index division), MFLBM_SAT_SEG blocks per slice so
// that a 512-slice lattice fills the 148 SMs more than once.
__global__ void __launch_bounds__(256) k_saturation(const Dev P, double *out) {
    const int k = blockIdx.x + 1;
    const int nx = P.g.nx, ny = P.g.ny;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double v[2] = {0, 0};
    for (int j = 1 + blockIdx.y * 8 + wid; j <= ny; j += 8 * gridDim.y) {
        const int c0 = P.g.cell(1, j, k);
        for (int i = lane; i < nx; i += 32) {
            const int wi = P.walls[c0 + i];
            const double ph = P.phi[c0 + i];
            v[0] += 0.5 * (1.0 + ph) * (1 - wi);
            v[1] += 0.5 * (1.0 - ph) * (1 - wi);
        }
    }
    __shared__ double sm[2][8];
#pragma unroll
    for (int n = 0; n < 2; n++) {
        double x = v[n];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) sm[n][wid] = x;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        const int n = threadIdx.x;
        double x = sm[n][0];
        for (int w = 1; w < 8; w++) x += sm[n][w];
        out[(size_t)n * P.g.nz * gridDim.y + blockIdx.x + (size_t)P.g.nz * blockIdx.y] = x;
    }
}

// sparse layout: the fluid nodes are exactly the A list, so the sum runs over it (12 B per fluid node instead of 9 B per
// lattice cell); out rows v1, v2 of gridDim.x block partials
__global__ void __launch_bounds__(256) k_saturation_list(const Dev P, double *out) {
    double v[2] = {0, 0};
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < P.nA; n += gridDim.x * blockDim.x) {
        const double ph = P.phi[P.cellA[n]];
        v[0] += 0.5 * (1.0 + ph);
        v[1] += 0.5 * (1.0 - ph);
    }
    const bool is_max[2] = {false, false};
    block_reduce<2>(v, is_max, out, gridDim.x, 0);
}

// returns the number of partial sums per row
int launch_saturation(mflbm_ctx *c, cudaStream_t st, double *out) {
    c->launches++;
    if (c->d.sparse) {
        int nb = 148 * 8;
        if (2 * nb > c->red_len) nb = c->red_len / 2;
        if (nb > (c->d.nA + 255) / 256) nb = (c->d.nA + 255) / 256;
        if (nb < 1) nb = 1;
        k_saturation_list<<<nb, 256, 0, st>>>(c->d, out);
        return nb;
    }
    k_saturation<<<dim3(c->d.g.nz, MFLBM_SAT_SEG), 256, 0, st>>>(c->d, out);
    return c->d.g.nz * MFLBM_SAT_SEG;
}

// breakthrough: count of fluid nodes with phi>0 on plane nz-1 (integer, exact); out[0..ny) per-row counts
__global__ void __launch_bounds__(256) k_breakthrough(const Dev P, double *out) {
    const int j = blockIdx.x + 1;
    const int k = P.g.nz - 1;
    double v[1] = {0};
    for (int i = threadIdx.x + 1; i <= P.g.nx; i += blockDim.x) {
        const int c = P.g.cell(i, j, k);
        if (P.walls[c] == 0 && P.phi[c] > 0.0) v[0] += 1.0;
    }
    const bool is_max[1] = {false};
    block_reduce<1>(v, is_max, out, P.g.ny, 0);
}

void launch_breakthrough(mflbm_ctx *c, cudaStream_t st, double *out) {
    k_breakthrough<<<c->d.g.ny, 256, 0, st>>>(c->d, out);
    c->launches++;
}

// steady state, phase field: rows umax, dphimax; also phi_old <- phi on 1..n
__global__ void __launch_bounds__(256) k_steady_phasefield(const Dev P, double *out) {
    const int k = blockIdx.x + 1;
    const int nx = P.g.nx, ny = P.g.ny;
    double v[2] = {0, 0};
    for (int n = threadIdx.x; n < nx * ny; n += blockDim.x) {
        const int j = n / nx + 1, i = n - (j - 1) * nx + 1;
        const int c = P.g.cell(i, j, k);
        const int wi = P.walls[c];
        const double uu = P.u[c], vv = P.v[c], ww = P.w[c];
        v[0] = fmax(v[0], (uu * uu + vv * vv + ww * ww) * (1 - wi));
        const double ph = P.phi[c];
        v[1] = fmax(v[1], fabs(ph - P.phi_old[c]) * (1 - wi));
        P.phi_old[c] = ph;
    }
    const bool is_max[2] = {true, true};
    block_reduce<2>(v, is_max, out, P.g.nz, 0);
}

void launch_steady_phasefield(mflbm_ctx *c, cudaStream_t st, double *out) {
    k_steady_phasefield<<<c->d.g.nz, 256, 0, st>>>(c->d, out);
    c->launches++;
}

// steady state, capillary pressure: rows umax, pre_w, pre_nw, i_w, i_nw (counts are exact in double up to 2^53)
__global__ void __launch_bounds__(256) k_steady_cappres(const Dev P, double *out) {
    const int k = blockIdx.x + 1;
    const int nx = P.g.nx, ny = P.g.ny;
    double v[5] = {0, 0, 0, 0, 0};
    for (int n = threadIdx.x; n < nx * ny; n += blockDim.x) {
        const int j = n / nx + 1, i = n - (j - 1) * nx + 1;
        const int c = P.g.cell(i, j, k);
        const int wi = P.walls[c];
        const double uu = P.u[c], vv = P.v[c], ww = P.w[c];
        v[0] = fmax(v[0], (uu * uu + vv * vv + ww * ww) * (1 - wi));
        if (wi == 0) {
            const double ph = P.phi[c], rho = P.rho[c];
            if (ph < -0.99) { v[1] += rho; v[3] += 1.0; }
            if (ph > 0.99) { v[2] += rho; v[4] += 1.0; }
        }
    }
    const bool is_max[5] = {true, false, false, false, false};
    block_reduce<5>(v, is_max, out, P.g.nz, 0);
}

void launch_steady_cappres(mflbm_ctx *c, cudaStream_t st, double *out) {
    k_steady_cappres<<<c->d.g.nz, 256, 0, st>>>(c->d, out);
    c->launches++;
}

// ---- layout conversion between the caller's Fortran arrays (ghost width o) and the padded grid ----
template <typename T>
__global__ void k_repack(const Grid g, T *grid, T *packed, int o, int kbase, int to_grid) {
    // packed is (1-o:nx+o, 1-o:ny+o, gridDim.z planes), i fastest; plane kk maps to grid plane k = kk + kbase
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1 - o;
    const int j = (int)blockIdx.y + 1 - o;
    const int kk = blockIdx.z;  // 0..nzp-1
    if (i > g.nx + o) return;
    const size_t p = (size_t)(i + o - 1) + (size_t)(g.nx + 2 * o) * ((size_t)(j + o - 1) + (size_t)(g.ny + 2 * o) * kk);
    const int c = g.cell(i, j, kk + kbase);
    if (to_grid) grid[c] = packed[p];
    else packed[p] = grid[c];
}

void launch_repack(mflbm_ctx *c, cudaStream_t st, double *grid, double *packed, int ghost, int nplanes_z, int kbase, bool to_grid) {
    const Grid &g = c->d.g;
    dim3 gr((g.nx + 2 * ghost + 127) / 128, g.ny + 2 * ghost, nplanes_z);
    k_repack<double><<<gr, 128, 0, st>>>(g, grid, packed, ghost, kbase, to_grid ? 1 : 0);
    c->launches++;
}

void launch_repack_i8(mflbm_ctx *c, cudaStream_t st, int8_t *grid, int8_t *packed, int ghost, bool to_grid) {
    const Grid &g = c->d.g;
    dim3 gr((g.nx + 2 * ghost + 127) / 128, g.ny + 2 * ghost, g.nz + 2 * ghost);
    k_repack<int8_t><<<gr, 128, 0, st>>>(g, grid, packed, ghost, 1 - ghost, to_grid ? 1 : 0);
    c->launches++;
}

}  // namespace mflbm
