// kernels_gradient.cu -- colour gradient / interface normal / curvature, reference-order dataflow.
//
// Replaces color_gradient (MP/Phase_gradient.F90:5-204) and alter_color_gradient_solid_surface
// (MP/Phase_gradient.F90:210-265): five launches K3..K7 exactly like the reference's five loop nests.
#include <algorithm>

#include "gradient.cuh"

namespace mflbm {

__device__ __forceinline__ double w_equ(int n) { return n <= 6 ? 1.0 / 18.0 : 1.0 / 36.0; }

// ---------------------------------------------------------------------------------------------------
// Per-entry device functions of the five loop nests; the launch shapes below (raster lists / dense loops for the
// reference-order variant, tile-driven for the sparse layout) only differ in how entries are enumerated.
// ---------------------------------------------------------------------------------------------------

// K3: phi on solid boundary nodes = weighted mean over listed fluid neighbours (MP/Phase_gradient.F90:16-29)
// flat = true: entry n of the flat-order copies of the lists (Dev::solid_cell_r ...)
__device__ __forceinline__ void phi_solid_at(const Dev &P, int n, bool flat = false) {
    const int c = (flat ? P.solid_cell_r : P.solid_cell)[n];
    const unsigned m = (flat ? P.solid_mask_r : P.solid_mask)[n];
    double acc = 0.0;
#pragma unroll
    for (int q = 1; q <= 18; q++)
        if (m & (1u << q)) acc = acc + P.phi[c + P.g.off(q)] * w_equ(q);
    P.phi[c] = acc / (flat ? P.solid_law_r : P.solid_law)[n];
}

// K4: ISO4 gradient of phi, norm, normalise; zero below 1e-6 (MP/Phase_gradient.F90:36-78).
// Solid nodes: the reference zeroes n and |grad phi| there on every call; the arrays are zero-initialised and only K6
// ever writes n at solid-boundary nodes (fully, after this kernel), so not touching solid nodes is unobservable.
template <bool LAZY>
__device__ __forceinline__ void gradient_at(const Dev &P, int c) {
    // (loading the 19 values as one pinned batch like curvature_at does -- Stencil19 -- was measured: flat K4 unchanged at
    // 1.19 ms on C3 with random phi, the tile-driven K4 slower because of its 56 registers; r02_m6)
    const int sx = P.g.sx, sxy = P.g.sxy;
    const double *__restrict__ ph = P.phi;
    auto v = [&](int a, int b, int d) { return ph[c + a + sx * b + sxy * d]; };
    const double gx = ddx(v), gy = ddy(v), gz = ddz(v);
    const double s2 = gx * gx + gy * gy + gz * gz;
    // sqrt is monotonic and correctly rounded: s2 < 0.99e-12 implies sqrt(s2) < 1e-6, so the (slow, FP64) square root is
    // only taken where the outcome of the reference's cut-off test can depend on it.  Most cells of an active tile are bulk.
    const double cn = s2 < 0.99e-12 ? 0.0 : sqrt(s2);
    if (cn < 1e-6) {
        // LAZY (sparse layout): bulk nodes already hold zeros; c_norm == 0 implies n == 0 at non-solid nodes
        if (LAZY && P.lazy_ok && P.c_norm[c] == 0.0) return;
        P.cn_x[c] = 0.0; P.cn_y[c] = 0.0; P.cn_z[c] = 0.0; P.c_norm[c] = 0.0;
    } else {
        P.cn_x[c] = gx / cn; P.cn_y[c] = gy / cn; P.cn_z[c] = gz / cn; P.c_norm[c] = cn;
    }
}

// K5: geometric wetting (Akai et al. 2018), MP/Phase_gradient.F90:225-261; cos/sin(theta) precomputed on the host
// the altered normal of one node: (x0, y0, z0) in, the closer of the two candidate directions out
__device__ __forceinline__ void alter_normal(const double nwx, const double nwy, const double nwz, const double tcos, const double tsin,
                                             double &x0, double &y0, double &z0) {
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

// K4 + K5 in one pass over a cell list (kf = entry of the cell in the fluid boundary list of the same order -- flat or grouped
// by tile -- or -1): the
// normal goes from the registers of K4 straight into K5 instead of through four dense arrays (K5 alone moved 3.2 GB on C3)
__device__ __forceinline__ void gradient_alter_at(const Dev &P, int c, int kf, bool flat) {
    const int sx = P.g.sx, sxy = P.g.sxy;
    const double *__restrict__ ph = P.phi;
    double nwx = 0.0, nwy = 0.0, nwz = 0.0, tcos = 0.0, tsin = 0.0;
    if (kf >= 0) {  // issued together with the stencil loads
        const size_t nf = (size_t)P.num_fluid;
        const double *__restrict__ nw = flat ? P.fluid_nw_r : P.fluid_nw;
        nwx = nw[kf]; nwy = nw[nf + kf]; nwz = nw[2 * nf + kf]; tcos = nw[3 * nf + kf]; tsin = nw[4 * nf + kf];
    }
    auto v = [&](int a, int b, int d) { return ph[c + a + sx * b + sxy * d]; };
    const double gx = ddx(v), gy = ddy(v), gz = ddz(v);
    const double s2 = gx * gx + gy * gy + gz * gz;
    const double cn = s2 < 0.99e-12 ? 0.0 : sqrt(s2);  // see gradient_at
    if (cn < 1e-6) {
        if (P.lazy_ok && P.c_norm[c] == 0.0) return;
        P.cn_x[c] = 0.0; P.cn_y[c] = 0.0; P.cn_z[c] = 0.0; P.c_norm[c] = 0.0;
    } else {
        double x0 = gx / cn, y0 = gy / cn, z0 = gz / cn;
        if (kf >= 0 && cn > 1e-6) alter_normal(nwx, nwy, nwz, tcos, tsin, x0, y0, z0);
        P.cn_x[c] = x0; P.cn_y[c] = y0; P.cn_z[c] = z0; P.c_norm[c] = cn;
    }
}

__device__ __forceinline__ void alter_at(const Dev &P, int n, bool flat = false) {
    const int c = (flat ? P.fluid_cell_r : P.fluid_cell)[n];
    const size_t nf = (size_t)P.num_fluid;
    if (!(P.c_norm[c] > 1e-6)) return;
    const double *__restrict__ nw = flat ? P.fluid_nw_r : P.fluid_nw;
    const double nwx = nw[n], nwy = nw[nf + n], nwz = nw[2 * nf + n];
    const double tcos = nw[3 * nf + n], tsin = nw[4 * nf + n];
    const double x0 = P.cn_x[c], y0 = P.cn_y[c], z0 = P.cn_z[c];
    double x1 = x0, y1 = y0, z1 = z0;
    alter_normal(nwx, nwy, nwz, tcos, tsin, x1, y1, z1);
    P.cn_x[c] = x1; P.cn_y[c] = y1; P.cn_z[c] = z1;
}

// K6: normal on solid boundary nodes inside the 0..n+1 box (MP/Phase_gradient.F90:88-109)
__device__ __forceinline__ void cn_solid_at(const Dev &P, int n, bool lazy = true, bool flat = false) {
    const unsigned m = (flat ? P.solid_mask_r : P.solid_mask)[n];
    if (!(m & 0x80000000u)) return;
    const int c = (flat ? P.solid_cell_r : P.solid_cell)[n];
    if (P.sparse && lazy) {
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
    const double law = (flat ? P.solid_law_r : P.solid_law)[n];
    P.cn_x[c] = ax / law;
    P.cn_y[c] = ay / law;
    P.cn_z[c] = az / law;
}

// ---- raster launch shapes (dense layout, sparse layout without tiles) ----
__global__ void k_phi_solid(const Dev P) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < P.num_solid) phi_solid_at(P, n);
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
template <bool LAZY>
__global__ void __launch_bounds__(128) k_gradient_list(const Dev P) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < P.nG) gradient_at<LAZY>(P, P.gcell[n]);
}

__global__ void k_alter(const Dev P) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < P.num_fluid) alter_at(P, n);
}

__global__ void k_cn_solid(const Dev P) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < P.num_solid) cn_solid_at(P, n);
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

// ---------------------------------------------------------------------------------------------------
// Quiet tiles (sparse multiphase layout).  Exactness argument (DESIGN.md "Quiet tiles"): the ISO4 gradient of values
// that all lie within +-1e-7 of the same constant is below 1.8e-7 < 1e-6, which the reference zeroes
// (MP/Phase_gradient.F90:64-73); tiles are at least 4 cells wide, so every phi value that can influence K3..K7 at a
// cell of tile T (radius 3) lies in T's 27-tile neighbourhood.  A tile turns quiet only after one full evaluation
// under uniform phi (two consecutive steps with the same single class), which leaves n = |grad phi| = 0 stored.
// The node lists are sorted by tile at upload (CSR ranges t*_start), so the chain only touches ACTIVE tiles:
// its cost follows the interfacial region, not the lattice.
// ---------------------------------------------------------------------------------------------------
#define TILE_P 1
#define TILE_M 2
#define TILE_X 4

__device__ __forceinline__ unsigned tile_class(double phi) {
    if (fabs(phi - 1.0) <= 1e-7) return TILE_P;
    if (fabs(phi + 1.0) <= 1e-7) return TILE_M;
    return TILE_X;  // also NaN
}

__device__ __forceinline__ void tile_or(unsigned char *arr, int tile, unsigned bits) {
    unsigned *w = (unsigned *)(arr + (tile & ~3));
    atomicOr(w, bits << (8 * (tile & 3)));
}

// one-time: mark[c] = 1 on solid boundary nodes (their phi is derived from fluid neighbours by K3)
__global__ void k_tile_mark_solid(const Dev P, unsigned char *mark) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < P.num_solid) mark[P.solid_cell[n]] = 1;
}

// one-time: class bits of every cell whose phi is read by the gradient stencils but is neither written by the collision
// kernel (A nodes) nor derived by K3: non-solid ghost cells, solid cells missing from the list.  Tiles that
// touch the z ghost planes (inlet/outlet values, periodic wrap, halo exchange) are always X.
__global__ void __launch_bounds__(128) k_tile_static(const Dev P, const unsigned char *mark) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x - 2;
    const int j = (int)blockIdx.y - 2;
    const int k = (int)blockIdx.z - 2;
    const int nx = P.g.nx, ny = P.g.ny, nz = P.g.nz;
    if (i > nx + 3) return;
    const int c = P.g.cell(i, j, k);
    const int t = P.g.tile_of(c, P.ntx, P.nty);
    if (P.jper && (j <= 0 || j >= ny + 1)) {  // y ghost rows are rewritten by the periodic wrap every step: always unknown
        tile_or(P.tstat, t, TILE_X);
        return;
    }
    if (k <= 0 || k >= nz + 1) {
        // ghost planes filled by the periodic wrap / the halo exchange: always X.  Ghost planes rewritten every step by
        // an inlet / outlet kernel (columns 1..nx x 1..ny): that kernel records the class of what it writes
        // (tile_record); the cells beside those columns are constant and classified below like any other unlisted cell.
        if (!(k <= 0 ? P.bc_lo_dyn : P.bc_hi_dyn)) {
            tile_or(P.tstat, t, TILE_X);
            return;
        }
        if (i >= 1 && i <= nx && j >= 1 && j <= ny) return;
    }
    const int a = P.smap[c];
    if (a >= 0 && a < P.nA) return;  // fluid node: dynamic class
    if (mark[c]) return;             // K3 node: derived from fluid nodes of the neighbourhood
    auto inG = [&](int ii, int jj, int kk) {
        return ii >= -1 && ii <= nx + 2 && jj >= -1 && jj <= ny + 2 && kk >= -1 && kk <= nz + 2 && P.walls[P.g.cell(ii, jj, kk)] != 1;
    };
    bool rel = inG(i, j, k);
#pragma unroll
    for (int q = 1; q < 19 && !rel; q++) rel = inG(i + EX(q), j + EY(q), k + EZ(q));
    if (rel) tile_or(P.tstat, t, tile_class(P.phi[c]));
}

// per step: U = OR of the class bits over the 27-tile neighbourhood; quiet iff U is one single class and equals the
// previous step's U.  Active tiles are appended to tact; every tile within one tile of an active tile is appended
// (once, guarded by a per-step stamp) to tk3: K3 must refresh phi on every solid node an active evaluation can read.
// Also clears the class buffer the next step will write.
// tz_lo..tz_hi, inside: only the tiles whose layer tz lies inside (inside = 1) / outside (inside = 0) that range -- the two
// halves of a speculative step (mflbm_api.cu step_impl); the whole lattice is (0, ntz - 1, 1).
__global__ void k_tile_update(const Dev P, int cur, int stamp, int tz_lo, int tz_hi, int inside) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ int s_cnt[4], s_base;
    const int tx = t % P.ntx, ty = (t / P.ntx) % P.nty, tz = t / (P.ntx * P.nty);
    const bool mine = t < P.ntiles && ((tz >= tz_lo && tz <= tz_hi) == (inside != 0));
    bool active = false;
    if (mine) {
    unsigned U = 0;
    for (int dz = -1; dz <= 1; dz++) {
        const int z = tz + dz;
        if (z < 0 || z >= P.ntz) continue;
        for (int dy = -1; dy <= 1; dy++) {
            const int y = ty + dy;
            if (y < 0 || y >= P.nty) continue;
            for (int dx = -1; dx <= 1; dx++) {
                const int x = tx + dx;
                if (x < 0 || x >= P.ntx) continue;
                const int o = x + P.ntx * (y + P.nty * z);
                U |= (unsigned)P.tcls[cur][o] | (unsigned)P.tstat[o];
            }
        }
    }
    const unsigned prev = P.tU[cur ^ 1][t];
    const bool quiet = (U == prev && (U == 0 || U == TILE_P || U == TILE_M));
    P.tU[cur][t] = (unsigned char)U;
    P.tquiet[t] = quiet ? 1 : 0;
    P.tcls[cur ^ 1][t] = 0;
    active = !quiet;
    }
    // Ordered append: the active tiles of this block (128 consecutive tile indices) go to tact as one run in index order,
    // one atomic per block.  With one atomic per tile the list came out scrambled, and the chain kernels, which walk it
    // block by block, lost the L2 reuse between neighbouring tiles' phi / normal boxes.
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, active);
    if (lane == 0) s_cnt[wib] = __popc(bal);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int tot = s_cnt[0] + s_cnt[1] + s_cnt[2] + s_cnt[3];
        s_base = tot ? atomicAdd(&P.tcount[0], tot) : 0;
    }
    __syncthreads();
    if (!active) return;
    int pos = s_base + __popc(bal & ((1u << lane) - 1u));
    for (int w = 0; w < wib; w++) pos += s_cnt[w];
    P.tact[pos] = t;
    if (P.mlist) {
        // march kernel: every (column, chunk) work item that holds a cell of 1..n of this tile (tile cells in lattice
        // coordinates: i = 8 tx - 3 .. 8 tx + 4, j = 4 ty - 3 .. 4 ty, k = 4 tz - 3 .. 4 tz), once per step
        const int ia = max(1, 8 * tx - 3), ib = min(P.g.nx, 8 * tx + 4);
        const int ja = max(1, 4 * ty - 3), jb = min(P.g.ny, 4 * ty);
        const int ka = max(1, 4 * tz - 3), kb = min(P.g.nz, 4 * tz);
        if (ia <= ib && ja <= jb && ka <= kb)
            for (int bz = (ka - 1) / P.march_lz; bz <= (kb - 1) / P.march_lz; bz++)
                for (int by = (ja - 1) / MFLBM_MARCH_TY; by <= (jb - 1) / MFLBM_MARCH_TY; by++)
                    for (int bx = (ia - 1) / MFLBM_MARCH_TX; bx <= (ib - 1) / MFLBM_MARCH_TX; bx++) {
                        const int item = bx + P.mcols_x * (by + P.mcols_y * bz);
                        if (atomicExch(&P.mflag[item], stamp) != stamp) P.mlist[atomicAdd(&P.tcount[10], 1)] = item;
                    }
    }
    atomicMax(&P.tcount[5], P.ntz - tz);  // layer range of the active tiles (zero-initialised: min as ntz - tz, max as tz + 1)
    atomicMax(&P.tcount[6], tz + 1);
    for (int dz = -1; dz <= 1; dz++) {
        const int z = tz + dz;
        if (z < 0 || z >= P.ntz) continue;
        for (int dy = -1; dy <= 1; dy++) {
            const int y = ty + dy;
            if (y < 0 || y >= P.nty) continue;
            for (int dx = -1; dx <= 1; dx++) {
                const int x = tx + dx;
                if (x < 0 || x >= P.ntx) continue;
                const int o = x + P.ntx * (y + P.nty * z);
                if (atomicExch(&P.tk3stamp[o], stamp) != stamp) P.tk3[atomicAdd(&P.tcount[1], 1)] = o;
            }
        }
    }
}

// Stamp every warp of 32 consecutive A nodes that owns a fluid node of ACTIVE tile `tile` (called by the 64 threads of
// a K4 block, two cells per thread).  Warps left with a stale stamp lie entirely in quiet tiles: the collision kernel
// knows |grad phi| = 0 there without reading anything (see Dev::wstamp).
__device__ __forceinline__ void tile_stamp_warps(const Dev &P, int tile, int stamp) {
    const int tx = tile % P.ntx, ty = (tile / P.ntx) % P.nty, tz = tile / (P.ntx * P.nty);
    for (int e = threadIdx.x; e < 128; e += blockDim.x) {  // cell of the 8x4x4 tile
        const int ix = 8 * tx + (e & 7), jy = 4 * ty + ((e >> 3) & 3), kz = 4 * tz + (e >> 5);
        if (jy >= P.g.ny + 8 || kz >= P.g.nz + 8) continue;
        const int a = P.smap[P.g.base - 4 + ix + P.g.sx * jy + P.g.sxy * kz];
        if (a >= 0 && a < P.nA) P.wstamp[a >> 5] = stamp;
    }
}

// K4 on the active tiles with phi staged through shared memory: the 128 threads of a block own the 128 cells of one
// 8x4x4 tile; the tile's phi box with a one-cell halo (10 x 6 x 6 values, rows of 10 contiguous doubles) is loaded once,
// cooperatively, and the 18-neighbour ISO4 stencils of all cells read it from there -- 360 loads per tile instead of
// 19 gathers per non-solid cell.  Cells are enumerated geometrically (non-solid cells of the (-1:n+2)^3 box, the
// reference's loop range, MP/Phase_gradient.F90:36-38), so neither the cell list nor its per-tile CSR is read.
// Same expressions in the same order as gradient_at (ddx / ddy / ddz), hence the same bits (parity suite green with it).
// MEASURED (r01_v11, profiles/r01_v11_k4.txt): slower than the list gathers of k_chain_tiles<4> -- 1050 vs 775 us on the
// 1536x1536x192 slab, 8600 vs 8620 MLUPS on C3: only ~36 % of a tile's cells are pore space, the list version packs
// them into full warps and its 19 gathers hit L1, while this one pays 360 + 128 loads and two barriers per tile
// whatever the porosity.  Kept selectable (MFLBM_K4_SMEM=1) as the evidence; the list version is the default.
#define MFLBM_K4_ROW 24  // shared-memory row pitch in doubles (10 used): 8 mod 16 keeps half-warps conflict-free
__global__ void __launch_bounds__(128) k_gradient_tiles(const Dev P, int stamp) {
    __shared__ double sphi[6 * 6 * MFLBM_K4_ROW];
    if (!P.tcount[3]) return;  // most tiles active: the flat kernels run instead
    const int count = P.tcount[0];
    const int sx = P.g.sx, sxy = P.g.sxy;
    const int tid = threadIdx.x;
    const int a = tid & 7, b = (tid >> 3) & 3, d = tid >> 5;
    for (int t = blockIdx.x; t < count; t += gridDim.x) {
        const int tile = P.tact[t];
        const int tx = tile % P.ntx, ty = (tile / P.ntx) % P.nty, tz = tile / (P.ntx * P.nty);
        const int ix0 = 8 * tx, jy0 = 4 * ty, kz0 = 4 * tz;  // padded coordinates (i+3, j+3, k+3) of the tile origin
        if (stamp > 0) tile_stamp_warps(P, tile, stamp);
        for (int e = tid; e < 360; e += 128) {
            const int ha = e % 10, hb = (e / 10) % 6, hd = e / 60;
            const long long c = (long long)(P.g.base - 4) + (ix0 + ha - 1) + (long long)sx * (jy0 + hb - 1) + (long long)sxy * (kz0 + hd - 1);
            sphi[ha + MFLBM_K4_ROW * (hb + 6 * hd)] = (c >= 0 && c < P.g.ntot) ? P.phi[c] : 0.0;
        }
        __syncthreads();
        const int ix = ix0 + a, jy = jy0 + b, kz = kz0 + d;  // this thread's cell
        const bool inbox = ix >= 2 && ix <= P.g.nx + 5 && jy >= 2 && jy <= P.g.ny + 5 && kz >= 2 && kz <= P.g.nz + 5;
        const int c = P.g.base - 4 + ix + sx * jy + sxy * kz;
        if (inbox && P.walls[c] != 1) {
            const double *__restrict__ sp = sphi + (a + 1) + MFLBM_K4_ROW * ((b + 1) + 6 * (d + 1));
            auto v = [&](int da, int db, int dd) { return sp[da + MFLBM_K4_ROW * (db + 6 * dd)]; };
            const double gx = ddx(v), gy = ddy(v), gz = ddz(v);
            const double cn = sqrt(gx * gx + gy * gy + gz * gz);
            if (cn < 1e-6) {
                if (P.c_norm[c] != 0.0) {  // lazy: bulk nodes already hold zeros (see gradient_at)
                    P.cn_x[c] = 0.0; P.cn_y[c] = 0.0; P.cn_z[c] = 0.0; P.c_norm[c] = 0.0;
                }
            } else {
                P.cn_x[c] = gx / cn; P.cn_y[c] = gy / cn; P.cn_z[c] = gz / cn; P.c_norm[c] = cn;
            }
        }
        __syncthreads();
    }
}

// MEASURED AND REJECTED (r02_k4, the 1536x1536x192 slab, 560 K active tiles): K4 on the active tiles without lists (one
// thread per tile cell, wall flag / active index / 19 phi values all independent loads after the tile index): 993 us
// against 992 us for k_chain_tiles<4>; neither deduplicating the warp-stamp stores (1025 us) nor appending the active
// tiles in index order (1019 us) moves it.  The kernel is bound by fetching ~3.5 KB of scattered 80-byte phi rows per
// tile, not by its dependent loads.
// every tile active (explicit mflbm_color_gradient)
__global__ void k_tile_all(const Dev P) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.ntiles) return;
    P.tact[t] = t;
    P.tk3[t] = t;
    if (t == 0) {
        P.tcount[0] = P.ntiles;
        P.tcount[1] = P.ntiles;
        P.tcount[2] = 1;
        P.tcount[3] = 0;
        P.tcount[5] = P.ntz;
        P.tcount[6] = P.ntz;
    }
}

// tile-driven launch shape: persistent blocks walk a tile list; the threads of a block share the entries of one tile
template <int K>
__global__ void __launch_bounds__(64) k_chain_tiles(const Dev P, int stamp, int force) {
    if (!force && !P.tcount[3]) return;  // most tiles active: k_chain_flat<K> runs instead (or nothing is left to do)
    if (K == 5 && P.gk5) return;         // done by K4
    const int *__restrict__ list = K == 3 ? P.tk3 : P.tact;
    const int count = P.tcount[K == 3 ? 1 : 0];
    const int *__restrict__ start = (K == 3 || K == 6) ? P.ts_start : (K == 4 ? P.tg_start : P.tf_start);
    for (int t = blockIdx.x; t < count; t += gridDim.x) {
        const int tile = list[t];
        if (K == 4 && stamp > 0) tile_stamp_warps(P, tile, stamp);
        const int e1 = start[tile + 1];
        for (int e = start[tile] + threadIdx.x; e < e1; e += 64) {
            if (K == 3) phi_solid_at(P, e);
            else if (K == 4) {
                if (P.gk5) gradient_alter_at(P, P.gcell[e], P.gk5[e], false);
                else gradient_at<true>(P, P.gcell[e]);
            } else if (K == 5) alter_at(P, e);
            else cn_solid_at(P, e);
        }
    }
}

// Which shape runs the chain of this step, decided on the device right after k_tile_update: when more than a quarter of
// the tiles are active (interface-rich states, e.g. the reference's benchmark case 6 with random phi) walking tile lists
// with one 64-thread block per tile wastes most lanes on short CSR ranges (measured on C3, random phi: K3..K6 5.8 ms
// tile-driven); the flat kernels below then sweep the whole node lists with full warps.  Both shapes are always
// launched, the one that is not selected returns at once (fixed grid-stride grids, so an idle launch costs microseconds).
// Evaluating a quiet tile anyway is exact: it reproduces the zeros / the K3 means the skipped evaluation would give.
// tcount[2] = 1: flat sweep (every warp then counts as active in the next collision), tcount[3] = 1: tile-driven pass.
// spec = 1 (second half of a speculative step): tcount[4] holds the number of active tiles the early pass already
// evaluated; nothing more runs unless the rest of the lattice turned out to hold active tiles too (a "miss"), in which
// case the whole list is evaluated again -- the chain is idempotent (K4 rewrites what K5 alters).
__global__ void k_tile_mode(const Dev P, int spec) {
    const bool many = (long long)P.tcount[0] * 4 > (long long)P.ntiles;
    const bool need = !spec || P.tcount[0] > P.tcount[4];
    P.tcount[2] = (need && many) ? 1 : 0;
    P.tcount[3] = (need && !many) ? 1 : 0;
}
__global__ void k_tile_mark(const Dev P) { P.tcount[4] = P.tcount[0]; }

template <int K>
__global__ void __launch_bounds__(256) k_chain_flat(const Dev P) {
    if (!P.tcount[2]) return;
    const int count = (K == 3 || K == 6) ? P.num_solid : (K == 4 ? P.nG : P.num_fluid);
    // K6: the 18 extra |grad phi| reads of the lazy path only pay while a good part of the lattice has no interface
    const bool lazy = P.lazy_ok && (long long)P.tcount[0] * 2 < (long long)P.ntiles;
    if (K == 5 && P.gk5_r) return;  // done by the K4 sweep
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
        if (K == 3) phi_solid_at(P, e, true);
        else if (K == 4) {
            if (P.gk5_r) gradient_alter_at(P, P.gcell_r[e], P.gk5_r[e], true);
            else gradient_at<true>(P, P.gcell_r[e]);
        } else if (K == 5) alter_at(P, e, true);
        else cn_solid_at(P, e, lazy, true);
    }
}

// Dev::gk5_r: scatter the flat-order fluid-list entries over a dense map (-1 = none), then read it back per K4 cell
__global__ void k_gk5_scatter(const int *__restrict__ fluid_cell, int num_fluid, int *map, int *dup) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= num_fluid) return;
    if (atomicExch(&map[fluid_cell[n]], n) != -1) atomicAdd(dup, 1);
}
__global__ void k_gk5_gather(const int *__restrict__ gcell, int nG, const int *map, int *gk5) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nG) gk5[e] = map[gcell[e]];
}
// flat = true: gk5_r from (gcell_r, fluid_cell_r); false: gk5 from the lists grouped by tile
void launch_build_gk5(mflbm_ctx *c, cudaStream_t st, int *map, int *dup, bool flat) {
    const Dev &P = c->d;
    if (P.num_fluid > 0) k_gk5_scatter<<<(P.num_fluid + 255) / 256, 256, 0, st>>>(flat ? P.fluid_cell_r : P.fluid_cell, P.num_fluid, map, dup);
    k_gk5_gather<<<(P.nG + 255) / 256, 256, 0, st>>>(flat ? P.gcell_r : P.gcell, P.nG, map, flat ? P.gk5_r : P.gk5);
    c->launches += 2;
}

// K7 + packing (Dev::G): one thread per fluid node of every ACTIVE warp (32 consecutive A nodes).  The blocks scan the
// warp stamps 32 at a time (one per lane, ballot) so that quiet warps cost one coalesced word each; all = 1 skips the scan.
__global__ void __launch_bounds__(256) k_gradient_pack(const Dev P, int all) {
    const int lane = threadIdx.x & 31;
    const int nW = (P.nA + 31) >> 5;
    if (!all && P.use_tiles && P.tcount[2]) return;  // k_gradient_pack_all did it
    // every warp draws groups of 32 node-warps from a ticket counter: the groups in flight stay inside one moving window
    // of the lattice (see k_gradient_pack_all), whatever the share of quiet warps in each group
    for (;;) {
        int w0 = 0;
        if (lane == 0) w0 = atomicAdd(&P.tcount[9], 1) * 32;
        w0 = __shfl_sync(0xffffffffu, w0, 0);
        if (w0 >= nW) return;
        unsigned m = 0xffffffffu;
        if (!all) m = __ballot_sync(0xffffffffu, w0 + lane < nW && P.wstamp[w0 + lane] == P.wq_stamp);
        else if (w0 + 32 > nW) m = nW - w0 >= 32 ? 0xffffffffu : ((1u << (nW - w0)) - 1u);
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const int n = ((w0 + b) << 5) + lane;
            if (n >= P.nA) continue;
            const int c = P.cellA[n];
            const double cnorm = P.c_norm[c];
            double cnx = 0.0, cny = 0.0, cnz = 0.0, tmp = 0.0;
            if (cnorm != 0.0) {  // c_norm == 0: n = 0 and F = 0.5*gamma*curv*0 = 0 whatever the curvature (exact zeros)
                cnx = P.cn_x[c]; cny = P.cn_y[c]; cnz = P.cn_z[c];
                tmp = 0.5 * P.gamma * curvature_at(P, c) * cnorm;  // MP/Kernel_multiphase.F90:118, MP/Phase_gradient.F90:116-200
            }
            P.G[0][n] = cnx; P.G[1][n] = cny; P.G[2][n] = cnz; P.G[3][n] = tmp;
        }
    }
}

// MEASURED AND REJECTED (r02_m4, C3 with random phi): the same with the normals staged through shared memory tile by
// tile (128 threads per 8x4x4 tile, 10x6x6 halo box per component, two barriers): 5.6 ms against 2.6 ms for the gathers
// above -- an 8x4x4 tile holds only ~46 fluid nodes, so most lanes idle through the stencil phase and every tile pays
// four dependent memory round trips plus the barriers; the gathers of the list version run with full warps and hit L1 /
// L2.  Same finding as for K4 in round 1 (k_gradient_tiles).
// Grid of a grid-stride kernel = exactly the blocks that are resident at once.  With more blocks than that, the first
// resident set strides through the WHOLE list before the next set starts, i.e. the lattice is swept several times and
// the stencil neighbourhoods fall out of L2 between the sweeps (measured, r02 ncu: k_gradient_pack with 4x oversubscribed
// grid read 13.7 GB from DRAM for 4.5 GB of operands).
template <typename K>
static int resident_grid(K kernel, int block) {
    int per_sm = 0, dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return per_sm * sms;
}

// Every fluid node.  Resident blocks draw chunks of 256 consecutive nodes from a ticket counter, so the chunks in flight
// always form one compact window of ~2 lattice planes and the three planes a node's stencil reads stay in L2 between
// their uses, like with an ordinary one-block-per-chunk launch -- but a launch that has nothing to do (tile mode) costs a
// few microseconds instead of one empty block per chunk (measured: 117 us on C3, 0.6 ns per block).  The grid-stride
// scan above lets blocks drift apart over its ~170 iterations: 13.7 GB DRAM reads for 4.5 GB of operands on C3 with random
// phi, every plane fetched three times (r02 ncu).  force = 0: runs only when the chain of this step swept every tile.
#ifndef MFLBM_PACK_MINB
#define MFLBM_PACK_MINB 3  // resident 256-thread blocks per SM the register allocation is bounded for (3 -> 80 registers; 2, 4, 5 measured slower: r02_pack)
#endif
__global__ void __launch_bounds__(256, MFLBM_PACK_MINB) k_gradient_pack_all(const Dev P, int force) {
    if (!force && !P.tcount[2]) return;
    __shared__ int s_chunk;
    const int nchunk = (P.nA + 255) >> 8;
    for (;;) {
        if (threadIdx.x == 0) s_chunk = atomicAdd(&P.tcount[8], 1);
        __syncthreads();
        const int chunk = s_chunk;
        __syncthreads();
        if (chunk >= nchunk) return;
        const int e = (chunk << 8) + threadIdx.x;
        if (e >= P.nA) continue;
        // brick order (MFLBM_BRICK7): the chunk's stencils overlap in y and z as well; cell and node index are independent loads
        const int n = P.aorder ? P.aorder[e] : e;
        const int c = P.aorder ? P.acell[e] : P.cellA[e];
        const double cnorm = P.c_norm[c];
        double cnx = 0.0, cny = 0.0, cnz = 0.0, tmp = 0.0;
        if (cnorm != 0.0) {
            cnx = P.cn_x[c]; cny = P.cn_y[c]; cnz = P.cn_z[c];
            tmp = 0.5 * P.gamma * curvature_at(P, c) * cnorm;
        }
        P.G[0][n] = cnx; P.G[1][n] = cny; P.G[2][n] = cnz; P.G[3][n] = tmp;
    }
}

void launch_gradient_pack(mflbm_ctx *c, cudaStream_t st) {
    const Dev &P = c->d;
    if (!P.multiphase || !P.sparse || P.nA <= 0) return;
    const int all = (!P.use_tiles || P.wq_all) ? 1 : 0;
    static int cap = 0, cap_all = 0;
    if (!cap) {
        cap = resident_grid(k_gradient_pack, 256);
        cap_all = resident_grid(k_gradient_pack_all, 256);
    }
    cudaMemsetAsync(P.tcount + 8, 0, 2 * sizeof(int), st);  // ticket counters of the two kernels
    k_gradient_pack_all<<<std::min(cap_all, (P.nA + 255) / 256), 256, 0, st>>>(P, all);
    c->launches++;
    if (all) return;
    int nb = (P.nA + 255) / 256;
    if (nb > cap) nb = cap;
    k_gradient_pack<<<nb, 256, 0, st>>>(P, 0);  // returns at once when tcount[2] is set
    c->launches++;
}

// every tile active: after create / upload / compute_macro_vars / an explicit mflbm_color_gradient
void launch_tiles_reset(mflbm_ctx *c, cudaStream_t st) {
    Dev &P = c->d;
    if (!P.use_tiles) return;
    P.wq_all = 1;
    const size_t n = (size_t)P.ntiles + 4;
    cudaMemsetAsync(P.tcls[0], 0, n, st);
    cudaMemsetAsync(P.tcls[1], 0, n, st);
    cudaMemsetAsync(P.tU[0], TILE_X, n, st);
    cudaMemsetAsync(P.tU[1], TILE_X, n, st);
    cudaMemsetAsync(P.tquiet, 0, n, st);
}

// (re)computes the static class bits; needs walls, smap, the solid list and phi on the device
int tiles_prepare(mflbm_ctx *c, cudaStream_t st) {
    const Dev &P = c->d;
    if (!P.use_tiles || c->tiles_static_ready) return 0;
    unsigned char *mark = nullptr;
    if (cudaMalloc((void **)&mark, (size_t)P.g.ntot) != cudaSuccess) return -1;
    cudaMemsetAsync(mark, 0, (size_t)P.g.ntot, st);
    cudaMemsetAsync(P.tstat, 0, (size_t)P.ntiles + 4, st);
    cudaMemsetAsync(P.tk3stamp, 0, (size_t)P.ntiles * sizeof(int), st);
    cudaMemsetAsync(P.wstamp, 0, ((size_t)(P.nA + 31) / 32 + 1) * sizeof(int), st);
    if (P.num_solid > 0) {
        k_tile_mark_solid<<<(P.num_solid + 255) / 256, 256, 0, st>>>(P, mark);
        c->launches++;
    }
    dim3 grid((P.g.nx + 6 + 127) / 128, P.g.ny + 6, P.g.nz + 6);
    k_tile_static<<<grid, 128, 0, st>>>(P, mark);
    c->launches++;
    cudaStreamSynchronize(st);
    cudaFree(mark);
    launch_tiles_reset(c, st);
    c->tile_stamp = 0;
    c->tiles_static_ready = true;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// phi on ALL listed solid boundary nodes (download / monitors want the reference's values everywhere)
void launch_phi_solid_refresh(mflbm_ctx *c, cudaStream_t st) {
    const Dev &P = c->d;
    if (!P.multiphase || P.num_solid <= 0) return;
    k_phi_solid<<<(P.num_solid + 127) / 128, 128, 0, st>>>(P);
    c->launches++;
    c->solid_phi_stale = false;
}

// stepping = true: called from mflbm_step right after the collision kernel, which recorded the phi classes of this
// step; otherwise (explicit mflbm_color_gradient, e.g. before the first step) every tile is evaluated.
// K3..K6: force = 1 runs the tile-driven kernels unconditionally over the lists as they are now (early half of a
// speculative step), otherwise the gated tile-driven AND flat kernels (whichever k_tile_mode / k_tile_all selected)
static void launch_chain_kernels(mflbm_ctx *c, cudaStream_t st, bool tiles, bool force) {
    const Dev &P = c->d;
    const int grid = P.ntiles < 148 * 32 ? P.ntiles : 148 * 32;
    static int fcap[4] = {0, 0, 0, 0};
    if (!fcap[0]) {
        fcap[0] = resident_grid(k_chain_flat<3>, 256); fcap[1] = resident_grid(k_chain_flat<4>, 256);
        fcap[2] = resident_grid(k_chain_flat<5>, 256); fcap[3] = resident_grid(k_chain_flat<6>, 256);
    }
    const int f = force ? 1 : 0;
    int n = 0;
    // K4 (tile-driven) also stamps the warps of the active tiles (nG > 0 whenever there is a fluid node)
    if (P.num_solid > 0) {
        if (tiles) { k_chain_tiles<3><<<grid, 64, 0, st>>>(P, 0, f); n++; }
        if (!force) { k_chain_flat<3><<<fcap[0], 256, 0, st>>>(P); n++; }
    }
    if (P.nG > 0) {
        if (tiles) {
            if (P.k4_smem == 1 && !force) k_gradient_tiles<<<P.ntiles < 148 * 16 ? P.ntiles : 148 * 16, 128, 0, st>>>(P, c->tile_stamp);
            else k_chain_tiles<4><<<grid, 64, 0, st>>>(P, c->tile_stamp, f);
            n++;
        }
        if (!force) { k_chain_flat<4><<<fcap[1], 256, 0, st>>>(P); n++; }
    }
    if (P.num_fluid > 0) {
        if (tiles) { k_chain_tiles<5><<<grid, 64, 0, st>>>(P, 0, f); n++; }
        if (!force) { k_chain_flat<5><<<fcap[2], 256, 0, st>>>(P); n++; }
    }
    if (P.num_solid > 0) {
        if (tiles) { k_chain_tiles<6><<<grid, 64, 0, st>>>(P, 0, f); n++; }
        if (!force) { k_chain_flat<6><<<fcap[3], 256, 0, st>>>(P); n++; }
    }
    c->launches += n;
}

// Speculative step, early half (stream st runs beside the collision of the far planes): tile update of the layers
// tz_lo..tz_hi only and the chain K3..K6 over the active tiles found there.  Everything these kernels read -- phi and tile
// classes of the layers tz_lo-2..tz_hi+2 -- has been written by the collision of the near planes already (step_impl).
void launch_chain_early(mflbm_ctx *c, cudaStream_t st, int tz_lo, int tz_hi) {
    Dev &P = c->d;
    const int nb = (P.ntiles + 127) / 128;
    // tcount[2] is left alone: the far-plane collision running beside this still reads it
    cudaMemsetAsync(P.tcount, 0, 2 * sizeof(int), st);
    cudaMemsetAsync(P.tcount + 3, 0, 5 * sizeof(int), st);
    ++c->tile_stamp;
    k_tile_update<<<nb, 128, 0, st>>>(P, P.tile_cur, c->tile_stamp, tz_lo, tz_hi, 1);
    k_tile_mark<<<1, 1, 0, st>>>(P);
    c->launches += 2;
    launch_chain_kernels(c, st, true, true);
}

// ... late half, after the far planes: the rest of the tile update; when it finds active tiles the early half did not
// know of, the whole chain runs again on the complete lists (exactly what a non-speculative step does); then K7 + packing.
void launch_chain_late(mflbm_ctx *c, cudaStream_t st, int tz_lo, int tz_hi) {
    Dev &P = c->d;
    const int nb = (P.ntiles + 127) / 128;
    k_tile_update<<<nb, 128, 0, st>>>(P, P.tile_cur, c->tile_stamp, tz_lo, tz_hi, 0);
    k_tile_mode<<<1, 1, 0, st>>>(P, 1);
    c->launches += 2;
    P.wq_stamp = c->tile_stamp;
    P.wq_all = 0;
    P.tile_cur ^= 1;
    c->solid_phi_stale = true;
    launch_chain_kernels(c, st, true, false);
    launch_gradient_pack(c, st);
}

// The reference-order chain K3..K6 over the whole lists into the DENSE arrays (phi on every listed solid node, n and
// |grad phi| everywhere), whatever they held before: contexts that run the march kernel keep these arrays only for the
// callers that ask for them (mflbm_download of the normals / the curvature, compute_macro_vars).
void launch_dense_gradient(mflbm_ctx *c, cudaStream_t st) {
    const Dev &P = c->d;
    if (!P.multiphase || !P.sparse || c->cn_dense_valid) return;
    if (P.num_solid > 0) k_phi_solid<<<(P.num_solid + 127) / 128, 128, 0, st>>>(P);
    if (P.nG > 0) k_gradient_list<false><<<(P.nG + 127) / 128, 128, 0, st>>>(P);
    if (P.num_fluid > 0) k_alter<<<(P.num_fluid + 127) / 128, 128, 0, st>>>(P);
    if (P.num_solid > 0) k_cn_solid<<<(P.num_solid + 127) / 128, 128, 0, st>>>(P);
    c->launches += 4;
    c->solid_phi_stale = false;
    c->cn_dense_valid = true;
}

// Self-check of the march kernel against the list kernels on the current state: evaluates the reference-order chain into
// the dense arrays, packs it into scratch buffers and counts the entries of G that differ bit for bit (0 = identical).
__global__ void k_count_diff(const unsigned long long *a, const unsigned long long *b, int n, unsigned long long *out) {
    unsigned long long bad = 0;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) bad += a[e] != b[e];
    if (bad) atomicAdd(out, bad);
}

long long chain_selfcheck(mflbm_ctx *c, cudaStream_t st) {
    const Dev &P = c->d;
    if (!P.multiphase || !P.sparse || P.nA <= 0) return 0;
    c->cn_dense_valid = false;
    launch_dense_gradient(c, st);
    Dev Q = P;
    double *scratch = nullptr;
    unsigned long long *cnt = nullptr;
    if (cudaMalloc((void **)&scratch, (size_t)4 * P.nA * sizeof(double)) != cudaSuccess) return -1;
    if (cudaMalloc((void **)&cnt, sizeof(unsigned long long)) != cudaSuccess) { cudaFree(scratch); return -1; }
    for (int m = 0; m < 4; m++) Q.G[m] = scratch + (size_t)m * P.nA;
    cudaMemsetAsync(P.tcount + 8, 0, 2 * sizeof(int), st);
    cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), st);
    k_gradient_pack_all<<<std::min(resident_grid(k_gradient_pack_all, 256), (P.nA + 255) / 256), 256, 0, st>>>(Q, 1);
    for (int m = 0; m < 4; m++)
        k_count_diff<<<1024, 256, 0, st>>>((const unsigned long long *)P.G[m], (const unsigned long long *)Q.G[m], P.nA, cnt);
    c->launches += 5;
    unsigned long long h = 0;
    const bool ok = cudaMemcpyAsync(&h, cnt, sizeof(h), cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess;
    cudaFree(scratch);
    cudaFree(cnt);
    return ok ? (long long)h : -1;
}

void launch_color_gradient(mflbm_ctx *c, cudaStream_t st, bool stepping) {
    Dev &P = c->d;
    if (!P.multiphase) return;
    if (c->march_on && c->march_ready && c->march_hybrid && P.use_tiles) {
        // hybrid: the fused kernel where few tiles are active (the work items around them), the flat sweeps of the list kernels
        // where most are -- whichever the tile update selects on the device.  The dense arrays are current only after a flat
        // evaluation, which the host cannot know: they count as stale, and the flat sweeps run without lazy shortcuts.
        c->solid_phi_stale = true;
        c->cn_dense_valid = false;
        cudaMemsetAsync(P.tcount, 0, 12 * sizeof(int), st);
        const int nb = (P.ntiles + 127) / 128;
        if (stepping) {
            k_tile_update<<<nb, 128, 0, st>>>(P, P.tile_cur, ++c->tile_stamp, 0, P.ntz - 1, 1);
            k_tile_mode<<<1, 1, 0, st>>>(P, 0);
            c->launches += 2;
            P.wq_stamp = c->tile_stamp;
            P.wq_all = 0;
            P.tile_cur ^= 1;
        } else {
            launch_tiles_reset(c, st);
            k_tile_all<<<nb, 128, 0, st>>>(P);  // tcount[2] = 1: the flat sweeps run
            c->launches++;
        }
        launch_chain_kernels(c, st, false, false);  // flat sweeps, gated by tcount[2]
        static int cap_all = 0;
        if (!cap_all) cap_all = resident_grid(k_gradient_pack_all, 256);
        k_gradient_pack_all<<<std::min(cap_all, (P.nA + 255) / 256), 256, 0, st>>>(P, 0);
        c->launches++;
        if (stepping) launch_march(c, st, 2, c->tile_stamp);  // gated by tcount[3]
        return;
    }
    if (c->march_on && c->march_ready) {
        // one kernel (march.cuh); phi on the solid nodes and the dense normal arrays are not written
        c->solid_phi_stale = true;
        c->cn_dense_valid = false;
        cudaMemsetAsync(P.tcount, 0, 12 * sizeof(int), st);
        if (P.use_tiles && stepping) {
            const int nb = (P.ntiles + 127) / 128;
            k_tile_update<<<nb, 128, 0, st>>>(P, P.tile_cur, ++c->tile_stamp, 0, P.ntz - 1, 1);
            k_tile_mode<<<1, 1, 0, st>>>(P, 0);
            c->launches += 2;
            P.wq_stamp = c->tile_stamp;
            P.wq_all = 0;
            P.tile_cur ^= 1;
            launch_march(c, st, 1, 0);                 // most tiles active: everything
            launch_march(c, st, 2, c->tile_stamp);     // else: the work items around the active tiles
        } else {
            if (P.use_tiles) {
                launch_tiles_reset(c, st);
                k_tile_all<<<(P.ntiles + 127) / 128, 128, 0, st>>>(P);
                c->launches++;
            }
            launch_march(c, st, 0, 0);
        }
        return;
    }
    if (P.use_tiles) {
        const int nb = (P.ntiles + 127) / 128;
        if (stepping) {
            cudaMemsetAsync(P.tcount, 0, 8 * sizeof(int), st);
            k_tile_update<<<nb, 128, 0, st>>>(P, P.tile_cur, ++c->tile_stamp, 0, P.ntz - 1, 1);
            k_tile_mode<<<1, 1, 0, st>>>(P, 0);
            c->launches += 2;
            P.wq_stamp = c->tile_stamp;
            P.wq_all = 0;
            P.tile_cur ^= 1;
            c->solid_phi_stale = true;
        } else {
            launch_tiles_reset(c, st);
            k_tile_all<<<nb, 128, 0, st>>>(P);  // tcount[2] = 1: the flat kernels run
            c->launches++;
            c->solid_phi_stale = false;
        }
        launch_chain_kernels(c, st, stepping, false);
        launch_gradient_pack(c, st);
        return;
    }
    if (P.num_solid > 0) {
        k_phi_solid<<<(P.num_solid + 127) / 128, 128, 0, st>>>(P);
        c->launches++;
    }
    c->solid_phi_stale = false;
    if (P.sparse) {
        if (P.nG > 0) {
            k_gradient_list<true><<<(P.nG + 127) / 128, 128, 0, st>>>(P);
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
    if (!P.sparse) launch_curvature(c, st);  // sparse layout: evaluated by k_gradient_pack for the fluid nodes
    else launch_gradient_pack(c, st);
}

}  // namespace mflbm
