// kernels_bc.cu -- open inlet/outlet boundary kernels and the porous-plate kernel on x-y planes.
//
// Replaces MP/Boundary_multiphase_inlet.F90 (:6-102 velocity, :115-294 Zou-He pressure),
// MP/Boundary_multiphase_outlet.F90 (:7-127 convective, :139-311 Zou-He pressure),
// MP/Boundary_multiphase_other.F90 (:8-184 porous plate) and SP/Boundary.F90 (:5-390).
// One thread per (i,j) column; every written address has exactly one writer and the planes read are
// never written by the same kernel, as in the reference's loops.
#include "mflbm_internal.cuh"

namespace mflbm {

#define W1 (1.0 / 18.0)
#define W2 (1.0 / 36.0)
#define C13 0.333333333333333333
#define C16 0.166666666666666667

// population index of dense cell c: identity on the dense layout, active-node index on the sparse layout.
// Every cell touched for a column whose boundary node is fluid is a D3Q19 neighbour of that fluid node,
// hence active; columns whose boundary node is solid are skipped on the sparse layout (the reference's
// blend new*(1-w)+old*w leaves them unchanged).
// y-periodic lattice (sparse layout only): a ghost-row cell (j = 0 or ny+1) of an interior plane stands for its periodic
// image -- the reference's y exchange keeps such rows equal to the image rows for k = 1..nz (MP/Mpi.F90:147-180), here the
// image row is the only storage.  Ghost-PLANE cells behind the seam are real storage when z is not periodic.
__device__ __forceinline__ int bc_index(const Dev &P, int c) {
    if (!P.sparse) return c;
    if (P.jper) {
        unsigned ix, jy, kz;
        P.g.coords3(c, ix, jy, kz);
        const int j = (int)jy - 3, k = (int)kz - 3;
        if (k >= 1 && k <= P.g.nz) {
            if (j == 0) c += P.g.sx * P.g.ny;
            else if (j == P.g.ny + 1) c -= P.g.sx * P.g.ny;
        }
    }
    return P.smap[c];
}
#define X(c) bc_index(P, (c))

__device__ __forceinline__ void phi_inlet_ghost(const Dev &P, int i, int j, int wi) {
    const int c0 = P.g.cell(i, j, 0);
    const double v = P.phi_inlet * (1 - wi) + P.phi[c0] * wi;
    P.phi[c0] = v;
    P.phi[c0 - P.g.sxy] = v;
    P.phi[c0 - 2 * P.g.sxy] = v;
    P.phi[c0 - 3 * P.g.sxy] = v;
    if (P.bc_lo_dyn && P.use_tiles) tile_record(P, c0, v);  // k = 0..-3 share one tile layer (kz = 3..0)
}

template <bool MP, bool AFTER>
__global__ void k_inlet_velocity(const Dev P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > P.g.nx) return;
    const int sx = P.g.sx;
    const int c0 = P.g.cell(i, j, 0), c1 = P.g.cell(i, j, 1);
    const int wi = P.walls[c1];
    if (MP) phi_inlet_ghost(P, i, j, wi);
    if (P.sparse && wi) return;
    double tmp2 = P.w_in[P.g.cell2(i, j)] * P.relaxation;
    const double tmp1 = MP ? tmp2 * P.sa_inject : tmp2;
    tmp2 = tmp2 - tmp1;
#pragma unroll
    for (int fl = 0; fl < (MP ? 2 : 1); fl++) {
        double *const *F = fl == 0 ? P.f : P.gg;
        const double t = fl == 0 ? tmp1 : tmp2;
        if (!AFTER) {
            F[5][X(c0)] = (F[6][X(c1)] + 6.0 * W1 * t) * (1 - wi) + F[5][X(c0)] * wi;
            F[11][X(c0 - 1)] = (F[14][X(c1)] + 6.0 * W2 * t) * (1 - wi) + F[11][X(c0 - 1)] * wi;
            F[12][X(c0 + 1)] = (F[13][X(c1)] + 6.0 * W2 * t) * (1 - wi) + F[12][X(c0 + 1)] * wi;
            F[15][X(c0 - sx)] = (F[18][X(c1)] + 6.0 * W2 * t) * (1 - wi) + F[15][X(c0 - sx)] * wi;
            F[16][X(c0 + sx)] = (F[17][X(c1)] + 6.0 * W2 * t) * (1 - wi) + F[16][X(c0 + sx)] * wi;
        } else {
            F[6][X(c1)] = (F[5][X(c0)] + 6.0 * W1 * t) * (1 - wi) + F[6][X(c1)] * wi;
            F[13][X(c1)] = (F[12][X(c0 + 1)] + 6.0 * W2 * t) * (1 - wi) + F[13][X(c1)] * wi;
            F[14][X(c1)] = (F[11][X(c0 - 1)] + 6.0 * W2 * t) * (1 - wi) + F[14][X(c1)] * wi;
            F[17][X(c1)] = (F[16][X(c0 + sx)] + 6.0 * W2 * t) * (1 - wi) + F[17][X(c1)] * wi;
            F[18][X(c1)] = (F[15][X(c0 - sx)] + 6.0 * W2 * t) * (1 - wi) + F[18][X(c1)] * wi;
        }
    }
}

template <bool MP, bool AFTER>
__global__ void k_inlet_pressure(const Dev P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > P.g.nx) return;
    const int sx = P.g.sx, sxy = P.g.sxy;
    const int c0 = P.g.cell(i, j, 0), c1 = c0 + sxy, c2 = c1 + sxy;
    const int wi = P.walls[c1];
    if (MP) phi_inlet_ghost(P, i, j, wi);
    if (P.sparse && wi) return;
    double r2 = P.rho_in;
    const double r1 = MP ? P.rho_in * P.sa_inject : P.rho_in;
    r2 = r2 - r1;
#pragma unroll
    for (int fl = 0; fl < (MP ? 2 : 1); fl++) {
        double *const *F = fl == 0 ? P.f : P.gg;
        const double rin = fl == 0 ? r1 : r2;
        if (!AFTER) {
            const double t = (rin - (F[0][X(c1)] + F[1][X(c1 - 1)] + F[2][X(c1 + 1)] + F[3][X(c1 - sx)] + F[4][X(c1 + sx)] + F[7][X(c1 - 1 - sx)] +
                                     F[8][X(c1 + 1 - sx)] + F[9][X(c1 - 1 + sx)] + F[10][X(c1 + 1 + sx)] +
                                     2.0 * (F[6][X(c2)] + F[14][X(c2 + 1)] + F[13][X(c2 - 1)] + F[18][X(c2 + sx)] + F[17][X(c2 - sx)]))) *
                             P.relaxation;
            const double tnx = 0.5 * (F[1][X(c1 - 1)] + F[7][X(c1 - 1 - sx)] + F[9][X(c1 - 1 + sx)] - (F[2][X(c1 + 1)] + F[8][X(c1 + 1 - sx)] + F[10][X(c1 + 1 + sx)]));
            const double tny = 0.5 * (F[3][X(c1 - sx)] + F[7][X(c1 - 1 - sx)] + F[8][X(c1 + 1 - sx)] - (F[4][X(c1 + sx)] + F[10][X(c1 + 1 + sx)] + F[9][X(c1 - 1 + sx)]));
            F[5][X(c0)] = (F[6][X(c2)] + C13 * t) * (1 - wi) + F[5][X(c0)] * wi;
            F[11][X(c0 - 1)] = (F[14][X(c2 + 1)] + C16 * t - tnx) * (1 - wi) + F[11][X(c0 - 1)] * wi;
            F[12][X(c0 + 1)] = (F[13][X(c2 - 1)] + C16 * t + tnx) * (1 - wi) + F[12][X(c0 + 1)] * wi;
            F[15][X(c0 - sx)] = (F[18][X(c2 + sx)] + C16 * t - tny) * (1 - wi) + F[15][X(c0 - sx)] * wi;
            F[16][X(c0 + sx)] = (F[17][X(c2 - sx)] + C16 * t + tny) * (1 - wi) + F[16][X(c0 + sx)] * wi;
        } else {
            const double t = (rin - (F[0][X(c1)] + F[2][X(c1)] + F[1][X(c1)] + F[4][X(c1)] + F[3][X(c1)] + F[8][X(c1)] + F[7][X(c1)] + F[10][X(c1)] + F[9][X(c1)] +
                                     2.0 * (F[5][X(c1)] + F[11][X(c1)] + F[12][X(c1)] + F[15][X(c1)] + F[16][X(c1)]))) *
                             P.relaxation;
            const double tnx = 0.5 * (F[2][X(c1)] + F[8][X(c1)] + F[10][X(c1)] - (F[1][X(c1)] + F[7][X(c1)] + F[9][X(c1)]));
            const double tny = 0.5 * (F[4][X(c1)] + F[9][X(c1)] + F[10][X(c1)] - (F[3][X(c1)] + F[8][X(c1)] + F[7][X(c1)]));
            const double n6 = (F[5][X(c1)] + C13 * t) * (1 - wi) + F[6][X(c1)] * wi;
            const double n13 = (F[12][X(c1)] + C16 * t + tnx) * (1 - wi) + F[13][X(c1)] * wi;
            const double n14 = (F[11][X(c1)] + C16 * t - tnx) * (1 - wi) + F[14][X(c1)] * wi;
            const double n17 = (F[16][X(c1)] + C16 * t + tny) * (1 - wi) + F[17][X(c1)] * wi;
            const double n18 = (F[15][X(c1)] + C16 * t - tny) * (1 - wi) + F[18][X(c1)] * wi;
            F[6][X(c1)] = n6; F[13][X(c1)] = n13; F[14][X(c1)] = n14; F[17][X(c1)] = n17; F[18][X(c1)] = n18;
        }
    }
}

template <bool MP, bool AFTER>
__global__ void k_outlet_convective(const Dev P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > P.g.nx) return;
    const int sx = P.g.sx, sxy = P.g.sxy, nz = P.g.nz;
    const int cn = P.g.cell(i, j, nz), cp = cn + sxy, cm = cn - sxy;
    const int c2 = P.g.cell2(i, j);
    const int wi = P.walls[cn];
    const double uc = P.uin_avg;
    const double temp = 1.0 / (1.0 + uc);
    if (MP) {
        const double v = ((P.phi_convec[c2] + uc * P.phi[cn]) * temp) * (1 - wi) + P.phi[cp] * wi;
        P.phi[cp] = v;
        P.phi_convec[c2] = v;
        P.phi[cp + sxy] = v;
        P.phi[cp + 2 * sxy] = v;
        P.phi[cp + 3 * sxy] = v;
        if (P.bc_hi_dyn && P.use_tiles) {  // k = nz+1 .. nz+4 lie in one or two tile layers
            tile_record(P, cp, v);
            if (nz & 3) tile_record(P, cp + 3 * sxy, v);
        }
    }
    if (P.sparse && wi) return;
#pragma unroll
    for (int fl = 0; fl < (MP ? 2 : 1); fl++) {
        double *const *F = fl == 0 ? P.f : P.gg;
        double *cb = fl == 0 ? P.f_convec : P.g_convec;
#define CB(q) cb[(q)*sxy + c2]
        if (!AFTER) {
            double v;
            v = ((CB(6) + uc * F[6][X(cn)]) * temp) * (1 - wi) + F[6][X(cp)] * wi;                 F[6][X(cp)] = v;       CB(6) = v;
            v = ((CB(13) + uc * F[13][X(cn - 1)]) * temp) * (1 - wi) + F[13][X(cp - 1)] * wi;      F[13][X(cp - 1)] = v;  CB(13) = v;
            v = ((CB(14) + uc * F[14][X(cn + 1)]) * temp) * (1 - wi) + F[14][X(cp + 1)] * wi;      F[14][X(cp + 1)] = v;  CB(14) = v;
            v = ((CB(17) + uc * F[17][X(cn - sx)]) * temp) * (1 - wi) + F[17][X(cp - sx)] * wi;    F[17][X(cp - sx)] = v; CB(17) = v;
            v = ((CB(18) + uc * F[18][X(cn + sx)]) * temp) * (1 - wi) + F[18][X(cp + sx)] * wi;    F[18][X(cp + sx)] = v; CB(18) = v;
        } else {
            double v;
            v = ((CB(6) + uc * F[5][X(cm)]) * temp) * (1 - wi) + F[5][X(cn)] * wi;     F[5][X(cn)] = v;  CB(6) = v;
            v = ((CB(14) + uc * F[11][X(cm)]) * temp) * (1 - wi) + F[11][X(cn)] * wi;  F[11][X(cn)] = v; CB(14) = v;
            v = ((CB(13) + uc * F[12][X(cm)]) * temp) * (1 - wi) + F[12][X(cn)] * wi;  F[12][X(cn)] = v; CB(13) = v;
            v = ((CB(18) + uc * F[15][X(cm)]) * temp) * (1 - wi) + F[15][X(cn)] * wi;  F[15][X(cn)] = v; CB(18) = v;
            v = ((CB(17) + uc * F[16][X(cm)]) * temp) * (1 - wi) + F[16][X(cn)] * wi;  F[16][X(cn)] = v; CB(17) = v;
        }
#undef CB
    }
}

template <bool MP, bool AFTER>
__global__ void k_outlet_pressure(const Dev P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > P.g.nx) return;
    const int sx = P.g.sx, sxy = P.g.sxy, nz = P.g.nz;
    const int cn = P.g.cell(i, j, nz), cp = cn + sxy, cm = cn - sxy;
    const int wi = P.walls[cn];
    const double dwi = 1.0 - wi;
    double phin = 0.0;
    if (MP) {
        phin = P.phi[cn];
        P.phi[cp] = phin; P.phi[cp + sxy] = phin; P.phi[cp + 2 * sxy] = phin; P.phi[cp + 3 * sxy] = phin;
        if (P.bc_hi_dyn && P.use_tiles) {
            tile_record(P, cp, phin);
            if (nz & 3) tile_record(P, cp + 3 * sxy, phin);
        }
    }
    if (P.sparse && wi) return;
    double *const *F = P.f;
    double *const *G = P.gg;
    double tmp1, tmp2 = 0.0;
    if (!AFTER) {
        if (MP)
            tmp1 = (F[0][X(cn)] + F[1][X(cn - 1)] + F[2][X(cn + 1)] + F[3][X(cn - sx)] + F[4][X(cn + sx)] + F[7][X(cn - 1 - sx)] + F[8][X(cn + 1 - sx)] +
                    F[9][X(cn - 1 + sx)] + F[10][X(cn + 1 + sx)] +
                    2.0 * (F[5][X(cm)] + F[11][X(cm - 1)] + F[12][X(cm + 1)] + F[15][X(cm - sx)] + F[16][X(cm + sx)]) + G[0][X(cn)] + G[1][X(cn - 1)] +
                    G[2][X(cn + 1)] + G[3][X(cn - sx)] + G[4][X(cn + sx)] + G[7][X(cn - 1 - sx)] + G[8][X(cn + 1 - sx)] + G[9][X(cn - 1 + sx)] +
                    G[10][X(cn + 1 + sx)] + 2.0 * (G[5][X(cm)] + G[11][X(cm - 1)] + G[12][X(cm + 1)] + G[15][X(cm - sx)] + G[16][X(cm + sx)])) -
                   P.rho_out;
        else
            tmp1 = (F[0][X(cn)] + F[1][X(cn - 1)] + F[2][X(cn + 1)] + F[3][X(cn - sx)] + F[4][X(cn + sx)] + F[7][X(cn - 1 - sx)] + F[8][X(cn + 1 - sx)] +
                    F[9][X(cn - 1 + sx)] + F[10][X(cn + 1 + sx)] +
                    2.0 * (F[5][X(cm)] + F[11][X(cm - 1)] + F[12][X(cm + 1)] + F[15][X(cm - sx)] + F[16][X(cm + sx)])) -
                   P.rho_out;
    } else {
        if (MP)
            tmp1 = (F[0][X(cn)] + F[2][X(cn)] + F[1][X(cn)] + F[4][X(cn)] + F[3][X(cn)] + F[8][X(cn)] + F[7][X(cn)] + F[10][X(cn)] + F[9][X(cn)] +
                    2.0 * (F[6][X(cn)] + F[14][X(cn)] + F[13][X(cn)] + F[18][X(cn)] + F[17][X(cn)]) + G[0][X(cn)] + G[2][X(cn)] + G[1][X(cn)] + G[4][X(cn)] +
                    G[3][X(cn)] + G[8][X(cn)] + G[7][X(cn)] + G[10][X(cn)] + G[9][X(cn)] +
                    2.0 * (G[6][X(cn)] + G[14][X(cn)] + G[13][X(cn)] + G[18][X(cn)] + G[17][X(cn)])) -
                   P.rho_out;
        else
            tmp1 = (F[0][X(cn)] + F[2][X(cn)] + F[1][X(cn)] + F[4][X(cn)] + F[3][X(cn)] + F[8][X(cn)] + F[7][X(cn)] + F[10][X(cn)] + F[9][X(cn)] +
                    2.0 * (F[6][X(cn)] + F[14][X(cn)] + F[13][X(cn)] + F[18][X(cn)] + F[17][X(cn)])) -
                   P.rho_out;
    }
    if (MP) {
        tmp2 = tmp1 * 0.5 * (1.0 - phin);
        tmp1 = tmp1 - tmp2;
    }
#pragma unroll
    for (int fl = 0; fl < (MP ? 2 : 1); fl++) {
        double *const *H = fl == 0 ? P.f : P.gg;
        const double t = fl == 0 ? tmp1 : tmp2;
        if (!AFTER) {
            const double tnx = 0.5 * (H[1][X(cn - 1)] + H[7][X(cn - 1 - sx)] + H[9][X(cn - 1 + sx)] - (H[2][X(cn + 1)] + H[8][X(cn + 1 - sx)] + H[10][X(cn + 1 + sx)]));
            const double tny = 0.5 * (H[3][X(cn - sx)] + H[7][X(cn - 1 - sx)] + H[8][X(cn + 1 - sx)] - (H[4][X(cn + sx)] + H[10][X(cn + 1 + sx)] + H[9][X(cn - 1 + sx)]));
            H[6][X(cp)] = (H[5][X(cm)] - C13 * t) * dwi + H[6][X(cp)] * wi;
            H[13][X(cp - 1)] = (H[12][X(cm + 1)] - C16 * t - tnx) * dwi + H[13][X(cp - 1)] * wi;
            H[14][X(cp + 1)] = (H[11][X(cm - 1)] - C16 * t + tnx) * dwi + H[14][X(cp + 1)] * wi;
            H[17][X(cp - sx)] = (H[16][X(cm + sx)] - C16 * t - tny) * dwi + H[17][X(cp - sx)] * wi;
            H[18][X(cp + sx)] = (H[15][X(cm - sx)] - C16 * t + tny) * dwi + H[18][X(cp + sx)] * wi;
        } else {
            const double tnx = 0.5 * (H[2][X(cn)] + H[8][X(cn)] + H[10][X(cn)] - (H[1][X(cn)] + H[7][X(cn)] + H[9][X(cn)]));
            const double tny = 0.5 * (H[4][X(cn)] + H[10][X(cn)] + H[9][X(cn)] - (H[3][X(cn)] + H[7][X(cn)] + H[8][X(cn)]));
            const double n5 = (H[6][X(cn)] - C13 * t) * dwi + H[5][X(cn)] * wi;
            const double n11 = (H[14][X(cn)] - C16 * t + tnx) * dwi + H[11][X(cn)] * wi;
            const double n12 = (H[13][X(cn)] - C16 * t - tnx) * dwi + H[12][X(cn)] * wi;
            const double n15 = (H[18][X(cn)] - C16 * t + tny) * dwi + H[15][X(cn)] * wi;
            const double n16 = (H[17][X(cn)] - C16 * t - tny) * dwi + H[16][X(cn)] * wi;
            H[5][X(cn)] = n5; H[11][X(cn)] = n11; H[12][X(cn)] = n12; H[15][X(cn)] = n15; H[16][X(cn)] = n16;
        }
    }
}

// porous plate: B = blocked fluid (bounced), T = passing fluid (copied through); zp = local plane
template <bool AFTER>
__global__ void k_porous_plate(const Dev P, int zp, int block_fluid1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1;
    if (i > P.g.nx) return;
    const int sx = P.g.sx, sxy = P.g.sxy;
    const int c = P.g.cell(i, j, zp), cm = c - sxy, cp = c + sxy;
    double *const *B = block_fluid1 ? P.f : P.gg;
    double *const *T = block_fluid1 ? P.gg : P.f;
    if (!AFTER) {
        B[6][X(c)] = B[5][X(cm)];
        B[13][X(c - 1)] = B[12][X(cm)];
        B[14][X(c + 1)] = B[11][X(cm)];
        B[17][X(c - sx)] = B[16][X(cm)];
        B[18][X(c + sx)] = B[15][X(cm)];
        B[5][X(c)] = B[6][X(cp)];
        B[12][X(c + 1)] = B[13][X(cp)];
        B[11][X(c - 1)] = B[14][X(cp)];
        B[16][X(c + sx)] = B[17][X(cp)];
        B[15][X(c - sx)] = B[18][X(cp)];
        T[6][X(c)] = T[6][X(cp)]; T[13][X(c)] = T[13][X(cp)]; T[14][X(c)] = T[14][X(cp)]; T[17][X(c)] = T[17][X(cp)]; T[18][X(c)] = T[18][X(cp)];
        T[5][X(c)] = T[5][X(cm)]; T[12][X(c)] = T[12][X(cm)]; T[11][X(c)] = T[11][X(cm)]; T[16][X(c)] = T[16][X(cm)]; T[15][X(c)] = T[15][X(cm)];
    } else {
        B[5][X(cm)] = B[6][X(c)];
        B[11][X(cm)] = B[14][X(c + 1)];
        B[12][X(cm)] = B[13][X(c - 1)];
        B[15][X(cm)] = B[18][X(c + sx)];
        B[16][X(cm)] = B[17][X(c - sx)];
        B[6][X(cp)] = B[5][X(c)];
        B[14][X(cp)] = B[11][X(c - 1)];
        B[13][X(cp)] = B[12][X(c + 1)];
        B[18][X(cp)] = B[15][X(c - sx)];
        B[17][X(cp)] = B[16][X(c + sx)];
        T[5][X(cm)] = T[5][X(c)]; T[11][X(cm)] = T[11][X(c)]; T[12][X(cm)] = T[12][X(c)]; T[15][X(cm)] = T[15][X(c)]; T[16][X(cm)] = T[16][X(c)];
        T[6][X(cp)] = T[6][X(c)]; T[14][X(cp)] = T[14][X(c)]; T[13][X(cp)] = T[13][X(c)]; T[18][X(cp)] = T[18][X(c)]; T[17][X(cp)] = T[17][X(c)];
    }
}

template <bool MP, bool AFTER>
static void launch_bc_t(mflbm_ctx *c, cudaStream_t st) {
    const Dev &P = c->d;
    const mflbm_config &cfg = c->cfg;
    dim3 block(128), grid((P.g.nx + 127) / 128, P.g.ny);
    if (c->open_z) {  // MP/Main_multiphase.F90:399-410, :462-473
        if (cfg.idz == 0) {
            if (cfg.inlet_BC == 1) { k_inlet_velocity<MP, AFTER><<<grid, block, 0, st>>>(P); c->launches++; }
            else if (cfg.inlet_BC == 2) { k_inlet_pressure<MP, AFTER><<<grid, block, 0, st>>>(P); c->launches++; }
        }
        if (cfg.idz == cfg.npz - 1) {
            if (cfg.outlet_BC == 1) { k_outlet_convective<MP, AFTER><<<grid, block, 0, st>>>(P); c->launches++; }
            else if (cfg.outlet_BC == 2) { k_outlet_pressure<MP, AFTER><<<grid, block, 0, st>>>(P); c->launches++; }
        }
    }
    if (MP && cfg.porous_plate_cmd != 0) {  // MP/Main_multiphase.F90:411-413, :474-476
        const int zmin = cfg.idz * cfg.nz + 1, zmax = cfg.idz * cfg.nz + cfg.nz;
        if (cfg.Z_porous_plate >= zmin && cfg.Z_porous_plate <= zmax && (cfg.porous_plate_cmd == 1 || cfg.porous_plate_cmd == 2)) {
            k_porous_plate<AFTER><<<grid, block, 0, st>>>(P, cfg.Z_porous_plate - cfg.idz * cfg.nz, cfg.porous_plate_cmd == 1);
            c->launches++;
        }
    }
}

void launch_bc(mflbm_ctx *c, cudaStream_t st, bool after_odd) {
    if (c->d.multiphase) {
        if (after_odd) launch_bc_t<true, true>(c, st);
        else launch_bc_t<true, false>(c, st);
    } else {
        if (after_odd) launch_bc_t<false, true>(c, st);
        else launch_bc_t<false, false>(c, st);
    }
}

}  // namespace mflbm
