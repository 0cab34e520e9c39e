// kernels_gradient.cu -- colour gradient / interface normal / curvature, reference-order dataflow.
//
// Replaces color_gradient (MP/Phase_gradient.F90:5-204) and alter_color_gradient_solid_surface
// (MP/Phase_gradient.F90:210-265): five launches K3..K7 exactly like the reference's five loop nests.
#include "mflbm_internal.cuh"

namespace mflbm {

#define ISO4_1 (1.0 / 6.0)
#define ISO4_2 (1.0 / 12.0)

// the three ISO4 central-difference shapes, terms in the reference's source order
template <typename F>
__device__ __forceinline__ double ddx(F v) {
    return ISO4_1 * (v(1, 0, 0) - v(-1, 0, 0)) +
           ISO4_2 * (v(1, 1, 0) - v(-1, -1, 0) + v(1, -1, 0) - v(-1, 1, 0) + v(1, 0, 1) - v(-1, 0, -1) + v(1, 0, -1) - v(-1, 0, 1));
}
template <typename F>
__device__ __forceinline__ double ddy(F v) {
    return ISO4_1 * (v(0, 1, 0) - v(0, -1, 0)) +
           ISO4_2 * (v(1, 1, 0) - v(-1, -1, 0) + v(-1, 1, 0) - v(1, -1, 0) + v(0, 1, 1) - v(0, -1, -1) + v(0, 1, -1) - v(0, -1, 1));
}
template <typename F>
__device__ __forceinline__ double ddz(F v) {
    return ISO4_1 * (v(0, 0, 1) - v(0, 0, -1)) +
           ISO4_2 * (v(1, 0, 1) - v(-1, 0, -1) + v(-1, 0, 1) - v(1, 0, -1) + v(0, 1, 1) - v(0, -1, -1) + v(0, -1, 1) - v(0, 1, -1));
}

__device__ __forceinline__ double w_equ(int n) { return n <= 6 ? 1.0 / 18.0 : 1.0 / 36.0; }

// K3: phi on solid boundary nodes = weighted mean over listed fluid neighbours (MP/Phase_gradient.F90:16-29)
__global__ void k_phi_solid(const Dev P) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= P.num_solid) return;
    const int c = P.solid_cell[n];
    const unsigned m = P.solid_mask[n];
    double acc = 0.0;
#pragma unroll
    for (int q = 1; q <= 18; q++)
        if (m & (1u << q)) acc = acc + P.phi[c + P.g.off(q)] * w_equ(q);
    P.phi[c] = acc / P.solid_law[n];
}

// K4: ISO4 gradient of phi, norm, normalise; zero on walls / below 1e-6 (MP/Phase_gradient.F90:36-78)
__global__ void __launch_bounds__(128) k_gradient(const Dev P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x - 1;
    const int j = (int)blockIdx.y - 1;
    const int k = (int)blockIdx.z - 1;
    if (i > P.g.nx + 2) return;
    const int c = P.g.cell(i, j, k);
    // solid nodes: the reference zeroes n and |grad phi| there every call; they are zero-initialised and only K6
    // ever writes n at solid-boundary nodes (fully, after this kernel), so skipping the store is unobservable.
    if (P.walls[c] == 1) return;
    const int sx = P.g.sx, sxy = P.g.sxy;
    const double *__restrict__ ph = P.phi;
    auto v = [&](int a, int b, int d) { return ph[c + a + sx * b + sxy * d]; };
    const double gx = ddx(v), gy = ddy(v), gz = ddz(v);
    const double cn = sqrt(gx * gx + gy * gy + gz * gz);
    if (cn < 1e-6) {
        P.cn_x[c] = 0.0; P.cn_y[c] = 0.0; P.cn_z[c] = 0.0; P.c_norm[c] = 0.0;
    } else {
        P.cn_x[c] = gx / cn; P.cn_y[c] = gy / cn; P.cn_z[c] = gz / cn; P.c_norm[c] = cn;
    }
}

// K5: geometric wetting (Akai et al. 2018), MP/Phase_gradient.F90:225-261; cos/sin(theta) precomputed on the host
__global__ void k_alter(const Dev P) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= P.num_fluid) return;
    const int c = P.fluid_cell[n];
    if (!(P.c_norm[c] > 1e-6)) return;
    const double nwx = P.fluid_nw[5 * n + 0], nwy = P.fluid_nw[5 * n + 1], nwz = P.fluid_nw[5 * n + 2];
    const double tcos = P.fluid_nw[5 * n + 3], tsin = P.fluid_nw[5 * n + 4];
    const double x0 = P.cn_x[c], y0 = P.cn_y[c], z0 = P.cn_z[c];
    const double t1 = nwx * x0 + nwy * y0 + nwz * z0;
    const double t2 = 1.0 / sqrt(1 - t1 * t1);
    const double coe1 = tsin * t1 * t2;
    const double coe2 = tsin * t2;
    const double xp = (tcos - coe1) * nwx + coe2 * x0;
    const double yp = (tcos - coe1) * nwy + coe2 * y0;
    const double zp = (tcos - coe1) * nwz + coe2 * z0;
    const double xm = (tcos + coe1) * nwx - coe2 * x0;
    const double ym = (tcos + coe1) * nwy - coe2 * y0;
    const double zm = (tcos + coe1) * nwz - coe2 * z0;
    const double dP = (xp - x0) * (xp - x0) + (yp - y0) * (yp - y0) + (zp - z0) * (zp - z0);
    const double dM = (xm - x0) * (xm - x0) + (ym - y0) * (ym - y0) + (zm - z0) * (zm - z0);
    if (dP <= dM) {
        P.cn_x[c] = xp; P.cn_y[c] = yp; P.cn_z[c] = zp;
    } else {
        P.cn_x[c] = xm; P.cn_y[c] = ym; P.cn_z[c] = zm;
    }
}

// K6: normal on solid boundary nodes inside the 0..n+1 box (MP/Phase_gradient.F90:88-109)
__global__ void k_cn_solid(const Dev P) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= P.num_solid) return;
    const unsigned m = P.solid_mask[n];
    if (!(m & 0x80000000u)) return;
    const int c = P.solid_cell[n];
    double ax = 0.0, ay = 0.0, az = 0.0;
#pragma unroll
    for (int q = 1; q <= 18; q++)
        if (m & (1u << q)) {
            const int cq = c + P.g.off(q);
            ax = ax + P.cn_x[cq] * w_equ(q);
            ay = ay + P.cn_y[cq] * w_equ(q);
            az = az + P.cn_z[cq] * w_equ(q);
        }
    const double law = P.solid_law[n];
    P.cn_x[c] = ax / law;
    P.cn_y[c] = ay / law;
    P.cn_z[c] = az / law;
}

// K7: curvature from the nine ISO4 derivatives of n, at all nodes incl. solids (MP/Phase_gradient.F90:116-200)
__global__ void __launch_bounds__(128) k_curvature(const Dev P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y + 1;
    const int k = blockIdx.z + 1;
    if (i > P.g.nx) return;
    const int c = P.g.cell(i, j, k);
    // the reference evaluates the curvature at solid nodes too (MP/Phase_gradient.F90:121, test commented out);
    // nothing on the hot path consumes it there, so it is only produced when full_curv is set.
    if (!P.full_curv && P.walls[c] != 0) return;
    const int sx = P.g.sx, sxy = P.g.sxy;
    const double *__restrict__ px = P.cn_x;
    const double *__restrict__ py = P.cn_y;
    const double *__restrict__ pz = P.cn_z;
    auto vx = [&](int a, int b, int d) { return px[c + a + sx * b + sxy * d]; };
    auto vy = [&](int a, int b, int d) { return py[c + a + sx * b + sxy * d]; };
    auto vz = [&](int a, int b, int d) { return pz[c + a + sx * b + sxy * d]; };
    const double kxx = ddx(vx), kyy = ddy(vy), kzz = ddz(vz);
    const double kxy = ddy(vx), kxz = ddz(vx);
    const double kyx = ddx(vy), kyz = ddz(vy);
    const double kzx = ddx(vz), kzy = ddy(vz);
    const double nx_ = px[c], ny_ = py[c], nz_ = pz[c];
    P.curv[c] = (nx_ * nx_ - 1.0) * kxx + (ny_ * ny_ - 1.0) * kyy + (nz_ * nz_ - 1.0) * kzz + nx_ * ny_ * (kxy + kyx) +
                nx_ * nz_ * (kxz + kzx) + ny_ * nz_ * (kzy + kyz);
}

void launch_color_gradient(mflbm_ctx *c, cudaStream_t st) {
    const Dev &P = c->d;
    if (!P.multiphase) return;
    if (P.num_solid > 0) {
        k_phi_solid<<<(P.num_solid + 127) / 128, 128, 0, st>>>(P);
        c->launches++;
    }
    {
        dim3 grid((P.g.nx + 4 + 127) / 128, P.g.ny + 4, P.g.nz + 4);
        k_gradient<<<grid, 128, 0, st>>>(P);
        c->launches++;
    }
    if (P.num_fluid > 0) {
        k_alter<<<(P.num_fluid + 127) / 128, 128, 0, st>>>(P);
        c->launches++;
    }
    if (P.num_solid > 0) {
        k_cn_solid<<<(P.num_solid + 127) / 128, 128, 0, st>>>(P);
        c->launches++;
    }
    {
        dim3 grid((P.g.nx + 127) / 128, P.g.ny, P.g.nz);
        k_curvature<<<grid, 128, 0, st>>>(P);
        c->launches++;
    }
}

}  // namespace mflbm
