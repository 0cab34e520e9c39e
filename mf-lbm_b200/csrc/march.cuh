// march.cuh -- the colour-gradient chain K3 -> K4 -> K5 -> K6 -> K7 in ONE kernel, on chip (DESIGN.md "March kernel").
//
// Replaces, on the sparse multiphase layout, the five list kernels of kernels_gradient.cu (color_gradient,
// MP/Phase_gradient.F90:5-204, and alter_color_gradient_solid_surface, :210-265) and the packing kernel: phi is read
// ONCE per cell, the normals never touch HBM, only the packed G[4][nA] the collision kernel reads is written.
//
// One block owns a column of MARCH_TX x MARCH_TY cells and marches through its planes k = kA..kB.  The five loop nests
// of the reference depend on each other through a one-cell neighbourhood each (K3: phi of the fluid neighbours of a
// solid node; K4: phi around a non-solid node; K5: the node itself; K6: normals of the fluid neighbours of a solid node;
// K7: normals around a fluid node), so plane q of stage s+1 can run once plane q+1 of stage s is done: the stages follow
// each other one plane apart, staggered over the two phases of a "tick" (one tick per plane, two barriers per tick):
//
//   tick t   top      plane t of phi / cell codes has landed in shared memory (cp.async issued two ticks earlier)
//            phase A  K3(t-1)  K6(t-4)   + task lists of phase B: K4 cells of plane t-2, K7 nodes of plane t-5
//            phase B  K4+K5(t-2)  K7(t-5) -> G   + task lists of the next phase A + cp.async of plane t+2
//
// Halo: K7 on the TX x TY cells needs normals one cell out (K6 there), K6 needs K4/K5 two cells out, K4 needs phi three
// cells out (K3 there), K3 reads raw phi four cells out -- exactly the reference's ghost widths (phi 4, normals 2).
// Work inside a plane is COMPACTED: ~40 % of a sphere pack's cells are pore space, so every phase first builds the list of
// cells that have a task (one ballot per 32 cells) and the tasks then run with full warps out of shared memory.
// What a cell is comes from one 32-bit code per cell (Dev::mcode, built once per upload by kernels_march.cu):
//   bits 1..0  type: 0 non-solid cell of the (-1:n+2)^3 box, 1 listed solid boundary node, 2 listed fluid boundary node,
//              3 anything else (solid without list entry, cell outside the box, row padding)
//   type 1: bits 2..19 = neighbour mask (bit 1+q <=> e_q listed, q = 1..18), bit 20 = inside the 0..n+1 box (K6 applies);
//           la_weight is recomputed from the mask (checked against the caller's value at build time)
//   type 2: bits 2..31 = index into the fluid boundary list (wall normal, cos / sin of the contact angle)
// Expressions and their order are those of kernels_gradient.cu / gradient.cuh, hence the same bits.
//
// The same source is compiled by g++ as a sequential emulation (MARCH_EMU: every phase loops over the thread index, barriers
// vanish) -- tests/test_march_emu.py runs it against the oracle on the CPU, where no GPU exists.
#pragma once
#include <math.h>

#include "gradient.cuh"

namespace mflbm {

#define MARCH_TX MFLBM_MARCH_TX
#define MARCH_TY MFLBM_MARCH_TY
#define MARCH_NT 512
#define M_PX4 (MARCH_TX + 8)
#define M_PY4 (MARCH_TY + 8)
#define M_N4 (M_PX4 * M_PY4)  // 960: phi / code region, halo 4
#define M_PX2 (MARCH_TX + 4)
#define M_PY2 (MARCH_TY + 4)
#define M_N2 (M_PX2 * M_PY2)  // 720: normals, halo 2
#define M_N0 (MARCH_TX * MARCH_TY)
#define M_N3 ((MARCH_TX + 6) * (MARCH_TY + 6))  // 836: K3 cells, halo 3
#define M_N1 ((MARCH_TX + 2) * (MARCH_TY + 2))  // 612: K6 cells, halo 1
#define M_RPHI 6   // ring depths in planes (see the lifetimes in march_block)
#define M_RCODE 6
#define M_RSMAP 3
#define M_RCN 5
#define M_RNORM 4

#define MCODE_FLUID 0u
#define MCODE_SOLID_LISTED 1u
#define MCODE_FLUID_LISTED 2u
#define MCODE_NONE 3u

struct alignas(16) MarchSmem {  // (the cp.async targets phi, code, smap must be 16-byte aligned)
    alignas(16) double phi[M_RPHI][M_N4];
    double cn[M_RCN][3][M_N2];
    double cnorm[M_RNORM][M_N0];
    double lawtab[7 * 13];  // la_weight of a node with a axis and b diagonal fluid neighbours, summed like the reference does
    alignas(16) unsigned code[M_RCODE][M_N4];
    alignas(16) int smap[M_RSMAP][M_N0];
    int cnt[4][2];  // task counts of the four lists, double-buffered by tick parity
    unsigned short l3[M_N3 + 28], l6[M_N1 + 28], l4[M_N2 + 16], l7[M_N0];
};

__host__ __device__ __forceinline__ double m_wequ(int n) { return n <= 6 ? 1.0 / 18.0 : 1.0 / 36.0; }

// la_weight from the neighbour counts: the reference adds w_equ(n) for n = 1..18 in order (MP/Geometry_preprocessing.F90:203-211),
// i.e. first the axis neighbours (1/18 each), then the diagonal ones (1/36 each)
__host__ __device__ __forceinline__ double m_law_from_counts(int a, int b) {
    double s = 0.0;
    for (int n = 0; n < a; n++) s = s + 1.0 / 18.0;
    for (int n = 0; n < b; n++) s = s + 1.0 / 36.0;
    return s;
}
// code payload of a listed solid node from the chain's mask word (bits 1..18 neighbours, bit 31 box flag)
__host__ __device__ __forceinline__ unsigned m_code_solid(unsigned mask) {
    return MCODE_SOLID_LISTED | (((mask >> 1) & 0x3ffffu) << 2) | ((mask >> 31) << 20);
}
__host__ __device__ __forceinline__ unsigned m_code_fluid(unsigned idx) { return MCODE_FLUID_LISTED | (idx << 2); }
// code of a cell before the node lists are scattered over it: non-solid cells of the (-1:n+2)^3 box are K4 cells
// (MP/Phase_gradient.F90:36-38 with the wall test of :39), everything else has no task
__host__ __device__ __forceinline__ unsigned m_code_base(const Grid &g, const int8_t *walls, long long c) {
    if (c < g.base - 4 || c >= (long long)(g.base - 4) + (long long)g.sxy * (g.nz + 8)) return MCODE_NONE;
    unsigned ix, jy, kz;
    g.coords3((int)c, ix, jy, kz);
    const int i = (int)ix - 3, j = (int)jy - 3, k = (int)kz - 3;
    if (i < -1 || i > g.nx + 2 || j < -1 || j > g.ny + 2 || k < -1 || k > g.nz + 2) return MCODE_NONE;
    return walls[c] != 1 ? MCODE_FLUID : MCODE_NONE;
}

#ifdef MARCH_EMU
#define MH_FN inline
#ifdef MARCH_EMU_REVERSE  // threads of a phase in the opposite order: the results must not depend on it
#define MH_FOR_TID(tid) for (int tid = MARCH_NT - 1; tid >= 0; tid--)
#else
#define MH_FOR_TID(tid) for (int tid = 0; tid < MARCH_NT; tid++)
#endif
#define MH_SYNC()
#define MH_POPC(x) __builtin_popcount(x)
#else
#define MH_FN __device__ __forceinline__
#define MH_FOR_TID(tid) for (int tid = threadIdx.x, once_ = 1; once_; once_ = 0)
#define MH_SYNC() __syncthreads()
#define MH_POPC(x) __popc(x)
#endif

// append `value` to a task list when pred; called by whole warps in lockstep (device: one shared-memory atomic per warp)
MH_FN void m_append(bool pred, int value, int *cnt, unsigned short *list) {
#ifdef MARCH_EMU
    if (pred) list[(*cnt)++] = (unsigned short)value;
#else
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(cnt, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (pred) list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)value;
#endif
}

// 16-byte asynchronous copy global -> shared (plain copy in the emulation)
MH_FN void m_copy16(void *dst, const void *src) {
#if defined(MARCH_EMU) || defined(MARCH_NO_CPASYNC)
    *(reinterpret_cast<uint4 *>(dst)) = *(reinterpret_cast<const uint4 *>(src));
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
#endif
}
MH_FN void m_commit() {
#if !defined(MARCH_EMU) && !defined(MARCH_NO_CPASYNC)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
MH_FN void m_wait1() {  // all but the most recent group have landed
#if !defined(MARCH_EMU) && !defined(MARCH_NO_CPASYNC)
    asm volatile("cp.async.wait_group 1;" ::: "memory");
#endif
}

__host__ __device__ __forceinline__ int m_slot(int q, int ring) { return (q + 10 * ring) % ring; }  // q >= -4 always
// slot of plane t - k (k in -2..6) given r = slot of plane t: the modulo is taken once per ring and tick (r is carried along),
// not once per use (r02 profile: the ring-index arithmetic was 11 % of the instructions)
__host__ __device__ __forceinline__ int m_rel(int r, int k, int ring) {
    const int kk = ((k % ring) + ring) % ring;  // constants: folded at compile time
    const int v = r - kk;
    return v < 0 ? v + ring : v;
}

// One block: column (bx, by), planes kA..kB (1 <= kA <= kB <= nz).  stamp > 0: stamp the warps of the nodes written (Dev::wstamp).
MH_FN void march_block(const Dev &P, MarchSmem &S, const int bx, const int by, const int kA, const int kB,
                                            const int stamp) {
    const Grid &g = P.g;
    const int i0 = 1 + MARCH_TX * bx, j0 = 1 + MARCH_TY * by;
    const int sx = g.sx, sxy = g.sxy;
    // cell of local (li, lj) = (0, 0) in plane q: g.cell(i0 - 4, j0 - 4, q)
    const long long cell00 = (long long)g.base + (i0 - 5) + (long long)sx * (j0 - 1);
    const int rows = g.ny + 8;  // allocated rows per plane
    const double gamma_half = 0.5 * P.gamma;

    // asynchronous loads of plane q: phi and codes (halo 4), and the active indices of plane qs (the column itself)
    auto issue_loads = [&](const int tid, const int q, const bool want, const int qs, const bool want_s, const int sl_phi, const int sl_code,
                           const int sl_smap) {
        if (want) {
            const long long pbase = cell00 + (long long)sxy * (q + 3);
            if (tid < M_PY4 * (M_PX4 / 2)) {  // 480 chunks of two doubles
                const int lj = tid / (M_PX4 / 2), li = 2 * (tid % (M_PX4 / 2));
                double *dst = &S.phi[sl_phi][li + M_PX4 * lj];
                if (i0 - 1 + li < sx && j0 - 1 + lj < rows) m_copy16(dst, P.phi + pbase + li + (long long)sx * lj);
                else { dst[0] = 0.0; dst[1] = 0.0; }
            }
            if (tid < M_PY4 * (M_PX4 / 4)) {  // 240 chunks of four codes
                const int lj = tid / (M_PX4 / 4), li = 4 * (tid % (M_PX4 / 4));
                unsigned *dst = &S.code[sl_code][li + M_PX4 * lj];
                if (i0 - 1 + li < sx && j0 - 1 + lj < rows) m_copy16(dst, P.mcode + pbase + li + (long long)sx * lj);
                else { dst[0] = MCODE_NONE; dst[1] = MCODE_NONE; dst[2] = MCODE_NONE; dst[3] = MCODE_NONE; }
            }
        }
        if (want_s && tid >= 256 && tid < 256 + MARCH_TY * (MARCH_TX / 4)) {  // 128 chunks of four indices
            const int ch = tid - 256;
            const int lj = ch / (MARCH_TX / 4), li = 4 * (ch % (MARCH_TX / 4));
            int *dst = &S.smap[sl_smap][li + MARCH_TX * lj];
            const long long c = cell00 + (long long)sxy * (qs + 3) + (li + 4) + (long long)sx * (lj + 4);
            if (i0 + 3 + li < sx && j0 + 3 + lj < rows) m_copy16(dst, P.smap + c);
            else { dst[0] = -1; dst[1] = -1; dst[2] = -1; dst[3] = -1; }
        }
    };

    // ---- prologue: law table, counters, the first two planes ----
    MH_FOR_TID(tid) {
        if (tid < 91) S.lawtab[tid] = m_law_from_counts(tid / 13, tid % 13);
        if (tid < 8) S.cnt[tid >> 1][tid & 1] = 0;
        issue_loads(tid, kA - 4, true, 0, false, m_slot(kA - 4, M_RPHI), m_slot(kA - 4, M_RCODE), 0);
        m_commit();
        issue_loads(tid, kA - 3, true, 0, false, m_slot(kA - 3, M_RPHI), m_slot(kA - 3, M_RCODE), 0);
        m_commit();
    }

    int r_phi = m_slot(kA - 4, M_RPHI), r_code = m_slot(kA - 4, M_RCODE), r_cn = m_slot(kA - 4, M_RCN), r_norm = m_slot(kA - 4, M_RNORM),
        r_smap = m_slot(kA - 4, M_RSMAP);  // ring slots of plane t
    for (int t = kA - 4; t <= kB + 5; t++, r_phi = m_rel(r_phi, -1, M_RPHI), r_code = m_rel(r_code, -1, M_RCODE),
             r_cn = m_rel(r_cn, -1, M_RCN), r_norm = m_rel(r_norm, -1, M_RNORM), r_smap = m_rel(r_smap, -1, M_RSMAP)) {
        const int par = t & 1;
        MH_FOR_TID(tid) { (void)tid; m_wait1(); }
        MH_SYNC();
        // ================= phase A =================
        const bool doK3 = t - 1 >= kA - 3 && t - 1 <= kB + 3;
        const bool doK6 = t - 4 >= kA - 1 && t - 4 <= kB + 1;
        const bool doL4 = t - 2 >= kA - 2 && t - 2 <= kB + 2;
        const bool doL7 = t - 5 >= kA && t - 5 <= kB;
        MH_FOR_TID(tid) {
            // K3 (plane t-1): phi on listed solid nodes from the raw phi of their listed fluid neighbours
            if (doK3) {
                const int n3 = S.cnt[0][par];
                double *p0 = S.phi[m_rel(r_phi, 1, M_RPHI)];
                const double *pm = S.phi[m_rel(r_phi, 2, M_RPHI)], *pp = S.phi[m_rel(r_phi, 0, M_RPHI)];
                const unsigned *cd = S.code[m_rel(r_code, 1, M_RCODE)];
                for (int idx = tid; idx < n3; idx += MARCH_NT) {
                    const int e = S.l3[idx];
                    const unsigned m = cd[e] >> 1;  // bit q <=> neighbour q listed (bits 1..18)
                    double acc = 0.0;
#pragma unroll
                    for (int q = 1; q <= 18; q++)
                        if (m & (1u << q)) {
                            const double *pl = EZ(q) < 0 ? pm : (EZ(q) > 0 ? pp : p0);
                            acc = acc + pl[e + EX(q) + M_PX4 * EY(q)] * m_wequ(q);
                        }
                    const int na = MH_POPC(m & 0x7eu), nb = MH_POPC(m & 0x7ff80u);
                    p0[e] = acc / S.lawtab[na * 13 + nb];
                }
            }
            // K6 (plane t-4): normal on listed solid nodes of the 0..n+1 box from the normals of their fluid neighbours
            if (doK6) {
                const int n6 = S.cnt[1][par];
                const int s0 = m_rel(r_cn, 4, M_RCN), sm = m_rel(r_cn, 5, M_RCN), sp = m_rel(r_cn, 3, M_RCN);
                const unsigned *cd = S.code[m_rel(r_code, 4, M_RCODE)];
                for (int idx = MARCH_NT - 1 - tid; idx < n6; idx += MARCH_NT) {
                    const int e = S.l6[idx];
                    const unsigned m = cd[e] >> 1;
                    const int e2 = (e % M_PX4 - 2) + M_PX2 * (e / M_PX4 - 2);
                    double ax = 0.0, ay = 0.0, az = 0.0;
#pragma unroll
                    for (int q = 1; q <= 18; q++)
                        if (m & (1u << q)) {
                            const int sl = EZ(q) < 0 ? sm : (EZ(q) > 0 ? sp : s0);
                            const int o = e2 + EX(q) + M_PX2 * EY(q);
                            ax = ax + S.cn[sl][0][o] * m_wequ(q);
                            ay = ay + S.cn[sl][1][o] * m_wequ(q);
                            az = az + S.cn[sl][2][o] * m_wequ(q);
                        }
                    const int na = MH_POPC(m & 0x7eu), nb = MH_POPC(m & 0x7ff80u);
                    const double law = S.lawtab[na * 13 + nb];
                    S.cn[s0][0][e2] = ax / law;
                    S.cn[s0][1][e2] = ay / law;
                    S.cn[s0][2][e2] = az / law;
                }
            }
            // task list of K4 (plane t-2): non-solid cells two cells out; every other cell of the region holds n = 0
            if (doL4) {
                const unsigned *cd = S.code[m_rel(r_code, 2, M_RCODE)];
                const int s2 = m_rel(r_cn, 2, M_RCN);
                for (int e2 = tid; e2 < ((M_N2 + 31) & ~31); e2 += MARCH_NT) {
                    bool task = false;
                    int e = 0;
                    if (e2 < M_N2) {
                        e = (e2 % M_PX2 + 2) + M_PX4 * (e2 / M_PX2 + 2);
                        const unsigned ty = cd[e] & 3u;
                        task = ty == MCODE_FLUID || ty == MCODE_FLUID_LISTED;
                        if (!task) { S.cn[s2][0][e2] = 0.0; S.cn[s2][1][e2] = 0.0; S.cn[s2][2][e2] = 0.0; }
                    }
                    m_append(task, e, &S.cnt[2][par], S.l4);
                }
            }
            // task list of K7 (plane t-5): the column's fluid nodes
            if (doL7) {
                const int *sm = S.smap[m_rel(r_smap, 5, M_RSMAP)];
                for (int e0 = tid; e0 < M_N0; e0 += MARCH_NT) {
                    const int a = sm[e0];
                    m_append(a >= 0 && a < P.nA, e0, &S.cnt[3][par], S.l7);
                }
            }
            if (tid == 0) { S.cnt[0][par ^ 1] = 0; S.cnt[1][par ^ 1] = 0; }  // lists the coming phase B builds
        }
        MH_SYNC();
        // ================= phase B =================
        const bool doL3 = t >= kA - 3 && t <= kB + 3;          // for K3(t) of the next tick
        const bool doL6 = t - 3 >= kA - 1 && t - 3 <= kB + 1;  // for K6(t-3) of the next tick
        MH_FOR_TID(tid) {
            // K4 + K5 (plane t-2)
            if (doL4) {
                const int n4 = S.cnt[2][par];
                const double *p0 = S.phi[m_rel(r_phi, 2, M_RPHI)], *pm = S.phi[m_rel(r_phi, 3, M_RPHI)], *pp = S.phi[m_rel(r_phi, 1, M_RPHI)];
                const unsigned *cd = S.code[m_rel(r_code, 2, M_RCODE)];
                const int s2 = m_rel(r_cn, 2, M_RCN), sn = m_rel(r_norm, 2, M_RNORM);
                const size_t nf = (size_t)P.num_fluid;
                for (int idx = tid; idx < n4; idx += MARCH_NT) {
                    const int e = S.l4[idx];
                    const int li = e % M_PX4, lj = e / M_PX4;
                    const unsigned code = cd[e];
                    const bool wet = (code & 3u) == MCODE_FLUID_LISTED;
                    double nwx = 0.0, nwy = 0.0, nwz = 0.0, tcos = 0.0, tsin = 0.0;
                    if (wet) {  // issued ahead of the stencil arithmetic
                        const size_t n = code >> 2;
                        nwx = P.fluid_nw[n]; nwy = P.fluid_nw[nf + n]; nwz = P.fluid_nw[2 * nf + n];
                        tcos = P.fluid_nw[3 * nf + n]; tsin = P.fluid_nw[4 * nf + n];
                    }
                    auto v = [&](int a, int b, int d) { return (d < 0 ? pm : (d > 0 ? pp : p0))[e + a + M_PX4 * b]; };
                    const double gx = ddx(v), gy = ddy(v), gz = ddz(v);
                    const double q2 = gx * gx + gy * gy + gz * gz;
                    const double cn = q2 < 0.99e-12 ? 0.0 : sqrt(q2);  // see gradient_at
                    double x0 = 0.0, y0 = 0.0, z0 = 0.0, cnorm = 0.0;
                    if (!(cn < 1e-6)) { x0 = gx / cn; y0 = gy / cn; z0 = gz / cn; cnorm = cn; }
                    if (wet && cnorm > 1e-6) {  // K5, alter_at
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
                        if (dP <= dM) { x0 = xp; y0 = yp; z0 = zp; }
                        else { x0 = xm; y0 = ym; z0 = zm; }
                    }
                    const int e2 = (li - 2) + M_PX2 * (lj - 2);
                    S.cn[s2][0][e2] = x0; S.cn[s2][1][e2] = y0; S.cn[s2][2][e2] = z0;
                    if (li >= 4 && li < 4 + MARCH_TX && lj >= 4 && lj < 4 + MARCH_TY) S.cnorm[sn][(li - 4) + MARCH_TX * (lj - 4)] = cnorm;
                }
            }
            // K7 (plane t-5) and the packed output
            if (doL7) {
                const int n7 = S.cnt[3][par];
                const int s0 = m_rel(r_cn, 5, M_RCN), sm = m_rel(r_cn, 6, M_RCN), sp = m_rel(r_cn, 4, M_RCN);
                const double *nrm = S.cnorm[m_rel(r_norm, 5, M_RNORM)];
                const int *smp = S.smap[m_rel(r_smap, 5, M_RSMAP)];
                for (int idx = MARCH_NT - 1 - tid; idx < n7; idx += MARCH_NT) {
                    const int e0 = S.l7[idx];
                    const int n = smp[e0];
                    const double cnorm = nrm[e0];
                    double cnx = 0.0, cny = 0.0, cnz = 0.0, tmp = 0.0;
                    if (cnorm != 0.0) {
                        const int e2 = (e0 % MARCH_TX + 2) + M_PX2 * (e0 / MARCH_TX + 2);
                        auto vx = [&](int a, int b, int d) { return S.cn[d < 0 ? sm : (d > 0 ? sp : s0)][0][e2 + a + M_PX2 * b]; };
                        auto vy = [&](int a, int b, int d) { return S.cn[d < 0 ? sm : (d > 0 ? sp : s0)][1][e2 + a + M_PX2 * b]; };
                        auto vz = [&](int a, int b, int d) { return S.cn[d < 0 ? sm : (d > 0 ? sp : s0)][2][e2 + a + M_PX2 * b]; };
                        const double kxx = ddx(vx), kxy = ddy(vx), kxz = ddz(vx), nx_ = vx(0, 0, 0);
                        const double kyx = ddx(vy), kyy = ddy(vy), kyz = ddz(vy), ny_ = vy(0, 0, 0);
                        const double kzx = ddx(vz), kzy = ddy(vz), kzz = ddz(vz), nz_ = vz(0, 0, 0);
                        const double curv = (nx_ * nx_ - 1.0) * kxx + (ny_ * ny_ - 1.0) * kyy + (nz_ * nz_ - 1.0) * kzz + nx_ * ny_ * (kxy + kyx) +
                                            nx_ * nz_ * (kxz + kzx) + ny_ * nz_ * (kzy + kyz);
                        cnx = nx_; cny = ny_; cnz = nz_;
                        tmp = gamma_half * curv * cnorm;
                    }
                    P.G[0][n] = cnx; P.G[1][n] = cny; P.G[2][n] = cnz; P.G[3][n] = tmp;
                    if (stamp > 0) P.wstamp[n >> 5] = stamp;
                }
            }
            // task lists of the next phase A
            if (doL3) {
                const unsigned *cd = S.code[m_rel(r_code, 0, M_RCODE)];
                for (int e3 = tid; e3 < ((M_N3 + 31) & ~31); e3 += MARCH_NT) {
                    bool task = false;
                    int e = 0;
                    if (e3 < M_N3) {
                        e = (e3 % (MARCH_TX + 6) + 1) + M_PX4 * (e3 / (MARCH_TX + 6) + 1);
                        task = (cd[e] & 3u) == MCODE_SOLID_LISTED;
                    }
                    m_append(task, e, &S.cnt[0][par ^ 1], S.l3);
                }
            }
            if (doL6) {
                const unsigned *cd = S.code[m_rel(r_code, 3, M_RCODE)];
                for (int e1 = tid; e1 < ((M_N1 + 31) & ~31); e1 += MARCH_NT) {
                    bool task = false;
                    int e = 0;
                    if (e1 < M_N1) {
                        e = (e1 % (MARCH_TX + 2) + 3) + M_PX4 * (e1 / (MARCH_TX + 2) + 3);
                        const unsigned c = cd[e];
                        task = (c & 3u) == MCODE_SOLID_LISTED && ((c >> 20) & 1u);
                    }
                    m_append(task, e, &S.cnt[1][par ^ 1], S.l6);
                }
            }
            if (tid == 0) { S.cnt[2][par ^ 1] = 0; S.cnt[3][par ^ 1] = 0; }  // lists the next phase A builds
            issue_loads(tid, t + 2, t + 2 <= kB + 4, t - 3, t - 3 >= kA && t - 3 <= kB, m_rel(r_phi, -2, M_RPHI), m_rel(r_code, -2, M_RCODE),
                        m_rel(r_smap, 3, M_RSMAP));
            m_commit();
        }
    }
}

}  // namespace mflbm
