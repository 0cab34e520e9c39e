// kernels_gradient.cu -- colour gradient / interface normal / curvature, reference-order dataflow.
//
// Replaces color_gradient (MP/Phase_gradient.F90:5-204) and alter_color_gradient_solid_surface
// (MP/Phase_gradient.F90:210-265): five launches K3..K7 exactly like the reference's five loop nests.
#include "gradient.cuh"

namespace mflbm {

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

// K4: ISO4 gradient of phi, norm, normalise; zero below 1e-6 (MP/Phase_gradient.F90:36-78).
// Solid nodes: the reference zeroes n and |grad phi| there on every call; the arrays are zero-initialised and only K6
// ever writes n at solid-boundary nodes (fully, after this kernel), so not touching solid nodes is unobservable.
template <bool LAZY>
__device__ __forceinline__ void gradient_at(const Dev &P, int c) {
    const int sx = P.g.sx, sxy = P.g.sxy;
    const double *__restrict__ ph = P.phi;
    auto v = [&](int a, int b, int d) { return ph[c + a + sx * b + sxy * d]; };
    const double gx = ddx(v), gy = ddy(v), gz = ddz(v);
    const double cn = sqrt(gx * gx + gy * gy + gz * gz);
    if (cn < 1e-6) {
        // LAZY (sparse layout): bulk nodes already hold zeros; c_norm == 0 implies n == 0 at non-solid nodes
        if (LAZY && P.c_norm[c] == 0.0) return;
        P.cn_x[c] = 0.0; P.cn_y[c] = 0.0; P.cn_z[c] = 0.0; P.c_norm[c] = 0.0;
    } else {
        P.cn_x[c] = gx / cn; P.cn_y[c] = gy / cn; P.cn_z[c] = gz / cn; P.c_norm[c] = cn;
    }
}

// dense traversal of the (-1:n+2)^3 box, like the reference's loop nest
__global__ void __launch_bounds__(128) k_gradient(const Dev P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x - 1;
    const int j = (int)blockIdx.y - 1;
    const int k = (int)blockIdx.z - 1;
    if (i > P.g.nx + 2) return;
    const int c = P.g.cell(i, j, k);
    if (P.walls[c] == 1) return;
    gradient_at<false>(P, c);
}

// traversal of the list of non-solid cells of the same box (sparse layout: work scales with the pore space)
__global__ void __launch_bounds__(128) k_gradient_list(const Dev P) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= P.nG) return;
    gradient_at<true>(P, P.gcell[n]);
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
    if (P.sparse) {
        // lazy path: if no listed fluid neighbour carries an interface (c_norm == 0 => n == 0) the result is exactly 0
        bool any = false;
#pragma unroll
        for (int q = 1; q <= 18; q++)
            if (m & (1u << q)) any |= (P.c_norm[c + P.g.off(q)] != 0.0);
        if (!any) {
            if (P.cn_x[c] != 0.0 || P.cn_y[c] != 0.0 || P.cn_z[c] != 0.0) {
                P.cn_x[c] = 0.0; P.cn_y[c] = 0.0; P.cn_z[c] = 0.0;
            }
            return;
        }
    }
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

// K7: curvature at all nodes incl. solids (MP/Phase_gradient.F90:116-200; the wall test at :121 is commented out).
// On the sparse layout nothing consumes the curvature at solid nodes and the collision kernel evaluates it on the
// fly, so this kernel only runs there when the field itself is requested (mflbm_download of curv).
__global__ void __launch_bounds__(128) k_curvature(const Dev P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y + 1;
    const int k = blockIdx.z + 1;
    if (i > P.g.nx) return;
    const int c = P.g.cell(i, j, k);
    if (!P.full_curv && P.walls[c] != 0) return;
    P.curv[c] = curvature_at(P, c);
}

void launch_curvature(mflbm_ctx *c, cudaStream_t st) {
    const Dev &P = c->d;
    dim3 grid((P.g.nx + 127) / 128, P.g.ny, P.g.nz);
    k_curvature<<<grid, 128, 0, st>>>(P);
    c->launches++;
}

void launch_color_gradient(mflbm_ctx *c, cudaStream_t st) {
    const Dev &P = c->d;
    if (!P.multiphase) return;
    if (P.num_solid > 0) {
        k_phi_solid<<<(P.num_solid + 127) / 128, 128, 0, st>>>(P);
        c->launches++;
    }
    if (P.sparse) {
        if (P.nG > 0) {
            k_gradient_list<<<(P.nG + 127) / 128, 128, 0, st>>>(P);
            c->launches++;
        }
    } else {
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
    if (!P.sparse) launch_curvature(c, st);  // sparse layout: evaluated inside the collision kernel
}

}  // namespace mflbm
