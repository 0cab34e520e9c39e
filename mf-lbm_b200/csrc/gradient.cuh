// gradient.cuh -- ISO4 finite differences of the colour-gradient model, shared by the gradient kernels and
// by the collision kernel (which evaluates the curvature on the fly on the sparse layout).
// Term order is the reference's source order (MP/Phase_gradient.F90:40-62, :122-195).
#pragma once
#include "mflbm_internal.cuh"

namespace mflbm {

#define MFLBM_ISO4_1 (1.0 / 6.0)
#define MFLBM_ISO4_2 (1.0 / 12.0)

template <typename F>
__device__ __forceinline__ double ddx(F v) {
    return MFLBM_ISO4_1 * (v(1, 0, 0) - v(-1, 0, 0)) +
           MFLBM_ISO4_2 * (v(1, 1, 0) - v(-1, -1, 0) + v(1, -1, 0) - v(-1, 1, 0) + v(1, 0, 1) - v(-1, 0, -1) + v(1, 0, -1) - v(-1, 0, 1));
}
template <typename F>
__device__ __forceinline__ double ddy(F v) {
    return MFLBM_ISO4_1 * (v(0, 1, 0) - v(0, -1, 0)) +
           MFLBM_ISO4_2 * (v(1, 1, 0) - v(-1, -1, 0) + v(-1, 1, 0) - v(1, -1, 0) + v(0, 1, 1) - v(0, -1, -1) + v(0, 1, -1) - v(0, -1, 1));
}
template <typename F>
__device__ __forceinline__ double ddz(F v) {
    return MFLBM_ISO4_1 * (v(0, 0, 1) - v(0, 0, -1)) +
           MFLBM_ISO4_2 * (v(1, 0, 1) - v(-1, 0, -1) + v(-1, 0, 1) - v(1, 0, -1) + v(0, 1, 1) - v(0, -1, -1) + v(0, -1, 1) - v(0, 1, -1));
}

#ifndef MARCH_EMU  // (the CPU emulation of march.cuh only needs the derivative templates above)
// The 19 stencil values of one field around cell c, ALL loaded before the first use.  The ISO4 sums are sequential
// dependency chains in the reference's term order; written as "load where used", ptxas keeps few registers and issues the
// loads in dependent batches (measured, r02 ncu: K4 / K7 kernels latency-bound at 25 % issue utilisation with every memory
// pipe below 50 %).  Index of offset (a, b, d), each in -1..1: (a+1) + 3 (b+1) + 9 (d+1); the eight corners are unused.
struct Stencil19 {
    double v[27];
    // sink: any int in global memory.  The never-taken branch below depends on EVERY loaded value and sits between the
    // loads and the arithmetic, which keeps the whole batch of loads in flight together (ptxas otherwise interleaves
    // them with the dependent DADD chain to stay at 32 registers: three to four memory round trips per node instead of one).
    __device__ __forceinline__ void load(const double *__restrict__ p, int c, int sx, int sxy, int *sink) {
        unsigned long long x = 0;
#pragma unroll
        for (int d = -1; d <= 1; d++)
#pragma unroll
            for (int b = -1; b <= 1; b++)
#pragma unroll
                for (int a = -1; a <= 1; a++)
                    if (a * a + b * b + d * d <= 2) {
                        const double t = p[c + a + sx * b + sxy * d];
                        v[(a + 1) + 3 * (b + 1) + 9 * (d + 1)] = t;
                        x ^= (unsigned long long)__double_as_longlong(t);
                    }
        if (x == 0x7ff8b200dead5eedULL) atomicAdd(sink, 0);  // a NaN pattern no stencil produces; harmless if it ever did
    }
    __device__ __forceinline__ double operator()(int a, int b, int d) const { return v[(a + 1) + 3 * (b + 1) + 9 * (d + 1)]; }
};

// K7 for one node: curvature from the nine ISO4 derivatives of the interface normal (MP/Phase_gradient.F90:116-200).
// The 18 neighbours' normals are each loaded once (54 loads, 18 in flight at a time) and reused by the three derivative shapes.
__device__ __forceinline__ double curvature_at(const Dev &P, int c) {
    const int sx = P.g.sx, sxy = P.g.sxy;
    Stencil19 s;
    s.load(P.cn_x, c, sx, sxy, P.tcount + 7);
    const double kxx = ddx(s), kxy = ddy(s), kxz = ddz(s), nx_ = s(0, 0, 0);
    s.load(P.cn_y, c, sx, sxy, P.tcount + 7);
    const double kyx = ddx(s), kyy = ddy(s), kyz = ddz(s), ny_ = s(0, 0, 0);
    s.load(P.cn_z, c, sx, sxy, P.tcount + 7);
    const double kzx = ddx(s), kzy = ddy(s), kzz = ddz(s), nz_ = s(0, 0, 0);
    return (nx_ * nx_ - 1.0) * kxx + (ny_ * ny_ - 1.0) * kyy + (nz_ * nz_ - 1.0) * kzz + nx_ * ny_ * (kxy + kyx) +
           nx_ * nz_ * (kxz + kzx) + ny_ * nz_ * (kzy + kyz);
}

#endif  // MARCH_EMU

}  // namespace mflbm
