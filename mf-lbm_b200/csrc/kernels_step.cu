// kernels_step.cu -- collision + AA-pattern in-place streaming (reference-order dataflow).
//
// Replaces kernel_odd_color / kernel_even_color (MP/Kernel_multiphase.F90:6-362, :371-725) and
// kernel_odd / kernel_even (SP/Kernel.F90:5-200, :206-400).  One thread per lattice node, x fastest
// (rows are 128-byte aligned so the even step is perfectly coalesced; the odd step touches the +-1
// neighbours in x/y/z).  Solid nodes are skipped but their slots stay live storage for bounced
// populations exactly as in the reference.
#include "collide.cuh"

namespace mflbm {

template <bool MP, bool ODD>
__global__ void __launch_bounds__(128) k_collide(const Dev P, int k0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y + 1;
    const int k = blockIdx.z + k0;
    if (i > P.g.nx) return;
    const int c = P.g.cell(i, j, k);
    if (P.walls[c] != 0) return;

    double a[19], b[19];
    if (ODD) {  // pull f_q from x - e_q (MP/Kernel_multiphase.F90:46-84)
#pragma unroll
        for (int q = 0; q < 19; q++) {
            const int cq = c - P.g.off(q);
            a[q] = P.f[q][cq];
            if (MP) b[q] = P.gg[q][cq];
        }
    } else {  // node-local, direction-swapped slots (MP/Kernel_multiphase.F90:410-448)
#pragma unroll
        for (int q = 0; q < 19; q++) {
            a[q] = P.f[OPC(q)][c];
            if (MP) b[q] = P.gg[OPC(q)][c];
        }
    }

    if (MP) {
        const double phi = collide_mp(P, a, b, P.cn_x[c], P.cn_y[c], P.cn_z[c], P.curv[c], P.c_norm[c]);
        P.phi[c] = phi;
    } else {
        collide_sp(P, a);
    }

    if (ODD) {  // push q into slot opc(q) of x + e_q (MP/Kernel_multiphase.F90:318-354)
        P.f[0][c] = a[0];
        if (MP) P.gg[0][c] = b[0];
#pragma unroll
        for (int q = 1; q < 19; q++) {
            const int cq = c + P.g.off(q);
            P.f[OPC(q)][cq] = a[q];
            if (MP) P.gg[OPC(q)][cq] = b[q];
        }
    } else {
#pragma unroll
        for (int q = 0; q < 19; q++) {
            P.f[q][c] = a[q];
            if (MP) P.gg[q][c] = b[q];
        }
    }
}

void launch_collide(mflbm_ctx *c, cudaStream_t st, bool odd, int k0, int k1) {
    if (k1 < k0) return;
    const Dev &P = c->d;
    dim3 block(128);
    dim3 grid((P.g.nx + 127) / 128, P.g.ny, k1 - k0 + 1);
    if (P.multiphase) {
        if (odd) k_collide<true, true><<<grid, block, 0, st>>>(P, k0);
        else k_collide<true, false><<<grid, block, 0, st>>>(P, k0);
    } else {
        if (odd) k_collide<false, true><<<grid, block, 0, st>>>(P, k0);
        else k_collide<false, false><<<grid, block, 0, st>>>(P, k0);
    }
    c->launches++;
}

// ---------------------------------------------------------------------------------------------------
// periodic z on one GPU: the reference's self send/recv through the periodic Cartesian communicator
// (MP/Mpi.F90:101-346 pull, :354-598 push, :608-867 phi).  One launch moves all planes.
// ---------------------------------------------------------------------------------------------------
template <bool MP>
__global__ void k_wrap_z(const Dev P, int push) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y + 1;
    if (i > P.g.nx) return;
    const int nz = P.g.nz;
    const int c0 = P.g.cell(i, j, 0), c1 = P.g.cell(i, j, 1), cn = P.g.cell(i, j, nz), cn1 = P.g.cell(i, j, nz + 1);
    constexpr int qM[5] = {6, 14, 13, 18, 17};  // e_z = -1
    constexpr int qP[5] = {5, 11, 12, 15, 16};  // e_z = +1
#pragma unroll
    for (int m = 0; m < 5; m++) {
        if (!push) {
            P.f[qM[m]][cn1] = P.f[qM[m]][c1];
            P.f[qP[m]][c0] = P.f[qP[m]][cn];
            if (MP) {
                P.gg[qM[m]][cn1] = P.gg[qM[m]][c1];
                P.gg[qP[m]][c0] = P.gg[qP[m]][cn];
            }
        } else {
            P.f[qP[m]][cn] = P.f[qP[m]][c0];
            P.f[qM[m]][c1] = P.f[qM[m]][cn1];
            if (MP) {
                P.gg[qP[m]][cn] = P.gg[qP[m]][c0];
                P.gg[qM[m]][c1] = P.gg[qM[m]][cn1];
            }
        }
    }
    if (MP) {
#pragma unroll
        for (int kk = 1; kk <= 4; kk++) {
            const double lo = P.phi[P.g.cell(i, j, kk)];
            const double hi = P.phi[P.g.cell(i, j, nz + kk - 4)];
            P.phi[P.g.cell(i, j, kk - 4)] = hi;
            P.phi[P.g.cell(i, j, kk + nz)] = lo;
        }
    }
}

void launch_wrap_z(mflbm_ctx *c, cudaStream_t st, bool push) {
    const Dev &P = c->d;
    dim3 block(128);
    dim3 grid((P.g.nx + 127) / 128, P.g.ny);
    if (P.multiphase) k_wrap_z<true><<<grid, block, 0, st>>>(P, push ? 1 : 0);
    else k_wrap_z<false><<<grid, block, 0, st>>>(P, push ? 1 : 0);
    c->launches++;
}

}  // namespace mflbm
