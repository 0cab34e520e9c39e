// collide.cuh -- FP64 MRT collision + colour-gradient recolouring for one lattice node.
//
// Arithmetic of kernel_odd_color / kernel_even_color (MP/Kernel_multiphase.F90:86-315, :450-678) and of
// kernel_odd / kernel_even (SP/Kernel.F90:54-178).  The expression order below is the reference's source
// order, so the strict build (-fmad=false) is bit-comparable with the CPU oracle; the default build lets
// ptxas contract a*b+c into DFMA (results then agree to ~1e-15 relative).
#pragma once
#include "mflbm_internal.cuh"

namespace mflbm {

struct Rates {
    double s_e, s_e2, s_q, s_nu, s_pi, s_t;
};

// 19 populations -> moments -> relaxation with Guo-type forcing -> back to populations
__device__ __forceinline__ void mrt_core(double (&ft)[19], double den, double fx, double fy, double fz, const Rates &r) {
    const double ft0 = ft[0], ft1 = ft[1], ft2 = ft[2], ft3 = ft[3], ft4 = ft[4], ft5 = ft[5], ft6 = ft[6], ft7 = ft[7],
                 ft8 = ft[8], ft9 = ft[9], ft10 = ft[10], ft11 = ft[11], ft12 = ft[12], ft13 = ft[13], ft14 = ft[14],
                 ft15 = ft[15], ft16 = ft[16], ft17 = ft[17], ft18 = ft[18];
    const double ux = ft1 - ft2 + ft7 - ft8 + ft9 - ft10 + ft11 - ft12 + ft13 - ft14 + 0.5 * fx;
    const double uy = ft3 - ft4 + ft7 + ft8 - ft9 - ft10 + ft15 - ft16 + ft17 - ft18 + 0.5 * fy;
    const double uz = ft5 - ft6 + ft11 + ft12 - ft13 - ft14 + ft15 + ft16 - ft17 - ft18 + 0.5 * fz;
    const double u2 = ux * ux + uy * uy + uz * uz;
    double sum1 = ft1 + ft2 + ft3 + ft4 + ft5 + ft6;
    double sum2 = ft7 + ft8 + ft9 + ft10 + ft11 + ft12 + ft13 + ft14 + ft15 + ft16 + ft17 + ft18;
    double sum3 = ft7 - ft8 + ft9 - ft10 + ft11 - ft12 + ft13 - ft14;
    double sum4 = ft7 + ft8 - ft9 - ft10 + ft15 - ft16 + ft17 - ft18;
    double sum5 = ft11 + ft12 - ft13 - ft14 + ft15 + ft16 - ft17 - ft18;
    double sum6 = 2.0 * (ft1 + ft2) - ft3 - ft4 - ft5 - ft6;
    double sum7 = ft7 + ft8 + ft9 + ft10 + ft11 + ft12 + ft13 + ft14 - 2.0 * (ft15 + ft16 + ft17 + ft18);
    double sum8 = ft3 + ft4 - ft5 - ft6;
    double sum9 = ft7 + ft8 + ft9 + ft10 - ft11 - ft12 - ft13 - ft14;

    double m_rho = den;
    double m_e = -30.0 * ft0 - 11.0 * sum1 + 8.0 * sum2;
    double m_e2 = 12.0 * ft0 - 4.0 * sum1 + sum2;
    double m_jx = ft1 - ft2 + sum3;
    double m_qx = -4.0 * (ft1 - ft2) + sum3;
    double m_jy = ft3 - ft4 + sum4;
    double m_qy = -4.0 * (ft3 - ft4) + sum4;
    double m_jz = ft5 - ft6 + sum5;
    double m_qz = -4.0 * (ft5 - ft6) + sum5;
    double m_3pxx = sum6 + sum7;
    double m_3pixx = -2.0 * sum6 + sum7;
    double m_pww = sum8 + sum9;
    double m_piww = -2.0 * sum8 + sum9;
    double m_pxy = ft7 - ft8 - ft9 + ft10;
    double m_pyz = ft15 - ft16 - ft17 + ft18;
    double m_pzx = ft11 - ft12 - ft13 + ft14;
    double m_tx = ft7 - ft8 + ft9 - ft10 - ft11 + ft12 - ft13 + ft14;
    double m_ty = -ft7 - ft8 + ft9 + ft10 + ft15 - ft16 + ft17 - ft18;
    double m_tz = ft11 + ft12 - ft13 - ft14 - ft15 - ft16 + ft17 + ft18;

    // MP/Module.F90:120-122: mrt_e2_coef1 = 0, mrt_e2_coef2 = -475/63, mrt_omega_xx = 0
    constexpr double e2c1 = 0.0, e2c2 = -475.0 / 63.0, oxx = 0.0;
    constexpr double c23 = 0.666666666666666667, c13 = 0.333333333333333333;
    const double fu = fx * ux + fy * uy + fz * uz;
    m_e = m_e - r.s_e * (m_e - (-11.0 * den + 19.0 * u2)) + (38.0 - 19.0 * r.s_e) * fu;
    m_e2 = m_e2 - r.s_e2 * (m_e2 - (e2c1 * den + e2c2 * u2)) + (-11.0 + 5.5 * r.s_e2) * fu;
    m_jx = m_jx + fx;
    m_qx = m_qx - r.s_q * (m_qx - (-c23 * ux)) + (-c23 + c13 * r.s_q) * fx;
    m_jy = m_jy + fy;
    m_qy = m_qy - r.s_q * (m_qy - (-c23 * uy)) + (-c23 + c13 * r.s_q) * fy;
    m_jz = m_jz + fz;
    m_qz = m_qz - r.s_q * (m_qz - (-c23 * uz)) + (-c23 + c13 * r.s_q) * fz;
    m_3pxx = m_3pxx - r.s_nu * (m_3pxx - (3.0 * ux * ux - u2)) + (2.0 - r.s_nu) * (2.0 * fx * ux - fy * uy - fz * uz);
    m_3pixx = m_3pixx - r.s_pi * (m_3pixx - oxx * (3.0 * ux * ux - u2)) + (1.0 - 0.5 * r.s_pi) * (-2.0 * fx * ux + fy * uy + fz * uz);
    m_pww = m_pww - r.s_nu * (m_pww - (uy * uy - uz * uz)) + (2.0 - r.s_nu) * (fy * uy - fz * uz);
    m_piww = m_piww - r.s_pi * (m_piww - oxx * (uy * uy - uz * uz)) + (1.0 - 0.5 * r.s_pi) * (-fy * uy + fz * uz);
    m_pxy = m_pxy - r.s_nu * (m_pxy - (ux * uy)) + (1.0 - 0.5 * r.s_nu) * (fx * uy + fy * ux);
    m_pyz = m_pyz - r.s_nu * (m_pyz - (uy * uz)) + (1.0 - 0.5 * r.s_nu) * (fy * uz + fz * uy);
    m_pzx = m_pzx - r.s_nu * (m_pzx - (ux * uz)) + (1.0 - 0.5 * r.s_nu) * (fx * uz + fz * ux);
    m_tx = m_tx - r.s_t * (m_tx);
    m_ty = m_ty - r.s_t * (m_ty);
    m_tz = m_tz - r.s_t * (m_tz);

    constexpr double k1 = 1.0 / 19.0, k2 = 1.0 / 2394.0, k3 = 1.0 / 252.0, k4 = 1.0 / 72.0;
    m_rho = k1 * m_rho;
    m_e = k2 * m_e;
    m_e2 = k3 * m_e2;
    m_jx = 0.1 * m_jx;
    m_qx = 0.025 * m_qx;
    m_jy = 0.1 * m_jy;
    m_qy = 0.025 * m_qy;
    m_jz = 0.1 * m_jz;
    m_qz = 0.025 * m_qz;
    m_3pxx = 2.0 * k4 * m_3pxx;
    m_3pixx = k4 * m_3pixx;
    m_pww = 6.0 * k4 * m_pww;
    m_piww = 3.0 * k4 * m_piww;
    m_pxy = 0.25 * m_pxy;
    m_pyz = 0.25 * m_pyz;
    m_pzx = 0.25 * m_pzx;
    m_tx = 0.125 * m_tx;
    m_ty = 0.125 * m_ty;
    m_tz = 0.125 * m_tz;
    sum1 = m_rho - 11.0 * m_e - 4.0 * m_e2;
    sum2 = 2.0 * m_3pxx - 4.0 * m_3pixx;
    sum3 = m_pww - 2.0 * m_piww;
    sum4 = m_rho + 8.0 * m_e + m_e2;
    sum5 = m_jx + m_qx;
    sum6 = m_jy + m_qy;
    sum7 = m_jz + m_qz;
    sum8 = m_3pxx + m_3pixx;
    sum9 = m_pww + m_piww;
    ft[0] = m_rho - 30.0 * m_e + 12.0 * m_e2;
    ft[1] = sum1 + m_jx - 4.0 * m_qx + sum2;
    ft[2] = sum1 - m_jx + 4.0 * m_qx + sum2;
    ft[3] = sum1 + m_jy - 4.0 * m_qy - 0.5 * sum2 + sum3;
    ft[4] = sum1 - m_jy + 4.0 * m_qy - 0.5 * sum2 + sum3;
    ft[5] = sum1 + m_jz - 4.0 * m_qz - 0.5 * sum2 - sum3;
    ft[6] = sum1 - m_jz + 4.0 * m_qz - 0.5 * sum2 - sum3;
    ft[7] = sum4 + sum5 + sum6 + sum8 + sum9 + m_pxy + m_tx - m_ty;
    ft[8] = sum4 - sum5 + sum6 + sum8 + sum9 - m_pxy - m_tx - m_ty;
    ft[9] = sum4 + sum5 - sum6 + sum8 + sum9 - m_pxy + m_tx + m_ty;
    ft[10] = sum4 - sum5 - sum6 + sum8 + sum9 + m_pxy - m_tx + m_ty;
    ft[11] = sum4 + sum5 + sum7 + sum8 - sum9 + m_pzx - m_tx + m_tz;
    ft[12] = sum4 - sum5 + sum7 + sum8 - sum9 - m_pzx + m_tx + m_tz;
    ft[13] = sum4 + sum5 - sum7 + sum8 - sum9 - m_pzx - m_tx - m_tz;
    ft[14] = sum4 - sum5 - sum7 + sum8 - sum9 + m_pzx + m_tx - m_tz;
    ft[15] = sum4 + sum6 + sum7 - sum8 * 2.0 + m_pyz + m_ty - m_tz;
    ft[16] = sum4 - sum6 + sum7 - sum8 * 2.0 - m_pyz - m_ty - m_tz;
    ft[17] = sum4 + sum6 - sum7 - sum8 * 2.0 - m_pyz + m_ty + m_tz;
    ft[18] = sum4 - sum6 - sum7 - sum8 * 2.0 + m_pyz - m_ty + m_tz;
}

// Multiphase node update.  a = fluid-1 incoming populations, b = fluid-2; both are overwritten with
// the recoloured post-collision populations.  tmp0 = 0.5*gamma*curv*c_norm evaluated in the reference's order
// (MP/Kernel_multiphase.F90:118) by the caller.  Returns phi.
__device__ __forceinline__ double collide_mp(const Dev &P, double (&a)[19], double (&b)[19], double cnx, double cny, double cnz,
                                             double tmp0) {
    double ft[19];
#pragma unroll
    for (int q = 0; q < 19; q++) ft[q] = a[q] + b[q];
    const double rho1 = a[0] + a[1] + a[2] + a[3] + a[4] + a[5] + a[6] + a[7] + a[8] + a[9] + a[10] + a[11] + a[12] + a[13] +
                        a[14] + a[15] + a[16] + a[17] + a[18];
    const double rho2 = b[0] + b[1] + b[2] + b[3] + b[4] + b[5] + b[6] + b[7] + b[8] + b[9] + b[10] + b[11] + b[12] + b[13] +
                        b[14] + b[15] + b[16] + b[17] + b[18];
    const double phi = (rho1 - rho2) / (rho1 + rho2);
    double tmp = tmp0;
    const double fx = tmp * cnx, fy = tmp * cny, fz = tmp * cnz + P.force_Z;
    const double omega = 1.0 / (6.0 / ((1.0 + phi) * P.la_nui1 + (1.0 - phi) * P.la_nui2) + 0.5);
    Rates r;
    r.s_nu = omega;
    if (P.mrt == 2) {  // MP/Kernel_multiphase.F90:134-140 (shipped)
        r.s_e = 1.19; r.s_e2 = 1.4; r.s_pi = 1.4; r.s_q = 1.2; r.s_t = 1.98;
    } else if (P.mrt == 1) {
        r.s_e = omega; r.s_e2 = omega; r.s_pi = omega; r.s_q = 8.0 * (2.0 - omega) / (8.0 - omega); r.s_t = r.s_q;
    } else if (P.mrt == 4) {
        r.s_e = omega; r.s_e2 = omega; r.s_pi = omega; r.s_q = (6.0 - 3.0 * omega) / (3.0 - omega); r.s_t = omega;
    } else {
        r.s_e = omega; r.s_e2 = omega; r.s_pi = omega; r.s_q = omega; r.s_t = omega;
    }
    const double den = rho1 + rho2;
    mrt_core(ft, den, fx, fy, fz, r);
    // R-K recolouring, MP/Kernel_multiphase.F90:272-315
    const double tmp1 = rho1 / den;
    a[0] = tmp1 * ft[0];
    b[0] = ft[0] * (1.0 - tmp1);
    tmp = rho1 * rho2 * P.beta / den;
    constexpr double w1 = 1.0 / 18.0;
    const double rk = P.rk_weight2;
    a[1] = tmp1 * ft[1] + w1 * tmp * (cnx);
    a[2] = tmp1 * ft[2] + w1 * tmp * (-cnx);
    a[3] = tmp1 * ft[3] + w1 * tmp * (cny);
    a[4] = tmp1 * ft[4] + w1 * tmp * (-cny);
    a[5] = tmp1 * ft[5] + w1 * tmp * (cnz);
    a[6] = tmp1 * ft[6] + w1 * tmp * (-cnz);
    a[7] = tmp1 * ft[7] + rk * tmp * (cnx + cny);
    a[8] = tmp1 * ft[8] + rk * tmp * (-cnx + cny);
    a[9] = tmp1 * ft[9] + rk * tmp * (cnx - cny);
    a[10] = tmp1 * ft[10] + rk * tmp * (-cnx - cny);
    a[11] = tmp1 * ft[11] + rk * tmp * (cnx + cnz);
    a[12] = tmp1 * ft[12] + rk * tmp * (-cnx + cnz);
    a[13] = tmp1 * ft[13] + rk * tmp * (cnx - cnz);
    a[14] = tmp1 * ft[14] + rk * tmp * (-cnx - cnz);
    a[15] = tmp1 * ft[15] + rk * tmp * (cny + cnz);
    a[16] = tmp1 * ft[16] + rk * tmp * (-cny + cnz);
    a[17] = tmp1 * ft[17] + rk * tmp * (cny - cnz);
    a[18] = tmp1 * ft[18] + rk * tmp * (-cny - cnz);
#pragma unroll
    for (int q = 1; q < 19; q++) b[q] = ft[q] - a[q];
    return phi;
}

__device__ __forceinline__ void collide_sp(const Dev &P, double (&ft)[19]) {
    const double den = ft[0] + ft[1] + ft[2] + ft[3] + ft[4] + ft[5] + ft[6] + ft[7] + ft[8] + ft[9] + ft[10] + ft[11] + ft[12] +
                       ft[13] + ft[14] + ft[15] + ft[16] + ft[17] + ft[18];
    Rates r{P.s_e, P.s_e2, P.s_q, P.s_nu, P.s_pi, P.s_t};
    mrt_core(ft, den, 0.0, 0.0, P.force_Z, r);
}

}  // namespace mflbm
