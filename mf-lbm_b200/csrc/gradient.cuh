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

// K7 for one node: curvature from the nine ISO4 derivatives of the interface normal (MP/Phase_gradient.F90:116-200).
// The 18 neighbours' normals are each loaded once (54 loads) and reused by the three derivative shapes.
__device__ __forceinline__ double curvature_at(const Dev &P, int c) {
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
    return (nx_ * nx_ - 1.0) * kxx + (ny_ * ny_ - 1.0) * kyy + (nz_ * nz_ - 1.0) * kzz + nx_ * ny_ * (kxy + kyx) +
           nx_ * nz_ * (kxz + kzx) + ny_ * nz_ * (kzy + kyz);
}

}  // namespace mflbm
