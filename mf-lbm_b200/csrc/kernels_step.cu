// kernels_step.cu -- collision + AA-pattern in-place streaming.
//
// Replaces kernel_odd_color / kernel_even_color (MP/Kernel_multiphase.F90:6-362, :371-725) and
// kernel_odd / kernel_even (SP/Kernel.F90:5-200, :206-400).  One thread per fluid node.
//
// Two population layouts share this kernel (template parameter SPARSE):
//  * dense  : the reference's direct addressing on the padded grid; solid nodes are skipped but their slots
//             are live storage for bounced populations.
//  * sparse : populations exist only for ACTIVE nodes (fluid nodes in raster order, a thin zone of ghost / solid nodes
//             at the slab ends, and one compact link slot per fluid node and wall direction).  Warps are fully
//             populated with fluid nodes, which is what the FP64 pipe and the HBM sectors want in porous media
//             (MLUPS counts fluid nodes only, MP/Main_multiphase.F90:540).  The even step is node-local and needs
//             no addressing at all; the odd step decodes the 18 neighbour indices of the node from the warp's
//             compressed adjacency records (mflbm_internal.cuh), staged through shared memory.  Same 38 addresses
//             are read and written per node, so the update stays race-free in any order, like the reference.
#include "collide.cuh"
#include "gradient.cuh"

namespace mflbm {

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute (launch_pdl below) may
// start while the previous kernel of the stream is still draining; everything above this wait must only touch data no kernel
// writes (adjacency records, cell lists, index maps).  A no-op in a kernel launched the ordinary way.
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ double ldpop(const double *p) { return *p; }
__device__ __forceinline__ void stpop(double *p, double v) { *p = v; }

// resident blocks of 128 threads per SM the register allocation is bounded for (4 -> 128 registers, 5 -> 96, 6 -> 80)
#ifndef MFLBM_MP_BLOCKS
#define MFLBM_MP_BLOCKS 4
#endif
#ifndef MFLBM_SP_BLOCKS
#define MFLBM_SP_BLOCKS 5  // measured on C2: 5 blocks (96 registers, 16-40 B of spills) +1 % over 4, 6 blocks -6 %
#endif

// Collision of one node in registers (a = fluid 1 / the single fluid, b = fluid 2), including what surrounds it on the
// multiphase path: interface normal / curvature inputs, the phi store and the per-tile phi classes.
template <bool MP, bool SPARSE>
__device__ __forceinline__ void collide_node(const Dev &P, const int n, const int c, const int wst, double (&a)[19], double (&b)[19]) {
    if (MP) {
        double cnx = 0.0, cny = 0.0, cnz = 0.0, tmp = 0.0;
        if (!SPARSE) {
            cnx = P.cn_x[c]; cny = P.cn_y[c]; cnz = P.cn_z[c];
            tmp = 0.5 * P.gamma * P.curv[c] * P.c_norm[c];
        } else {
            // Quiet warp (every node in a tile with uniform phi around, see kernels_gradient.cu): n = 0 and F = 0 are known
            // without reading anything.  Otherwise four coalesced reads of the packed gradient (Dev::G); the skipped terms
            // of a node without interface are exact zeros there.
            const bool quiet = P.use_tiles && !P.wq_all && wst != P.wq_stamp;  // warp-uniform
            if (!quiet) {
                cnx = P.G[0][n]; cny = P.G[1][n]; cnz = P.G[2][n]; tmp = P.G[3][n];
            }
        }
        const double phi = collide_mp(P, a, b, cnx, cny, cnz, tmp);
        P.phi[c] = phi;
        if (SPARSE && P.use_tiles) {
            // record the phi class of this node's tile: consecutive lanes mostly share (tile, class), so the first lane
            // of every run of equal keys issues one fire-and-forget atomic (a handful per warp)
            const int tile = P.g.tile_of(c, P.ntx, P.nty);
            const unsigned bits = fabs(phi - 1.0) <= 1e-7 ? 1u : (fabs(phi + 1.0) <= 1e-7 ? 2u : 4u);
            const unsigned key = ((unsigned)tile << 3) | bits;
            const unsigned act = __activemask();
            const unsigned prev = __shfl_up_sync(act, key, 1);
            const int lane = threadIdx.x & 31;
            if (lane == 0 || !((act >> (lane - 1)) & 1u) || prev != key)
                atomicOr((unsigned *)(P.tcls[P.tile_cur] + (tile & ~3)), bits << (8 * (tile & 3)));
        }
    } else {
        collide_sp(P, a);
    }
}

// One node: gather the incoming populations, collide, scatter.  n = active index (sparse layout), c = dense cell,
// wst = this warp's quiet stamp (Dev::wstamp), rec = this warp's adjacency records in shared memory (sparse odd step).
template <bool MP, bool ODD, bool SPARSE, int PF>
__device__ __forceinline__ void node_update(const Dev &P, const int n, const int c, const int wst, const uint4 *rec) {
    double a[19], b[19];
    int nb[19];  // sparse odd step: active index of x+e_q
    if (ODD) {   // pull f_q from x - e_q (MP/Kernel_multiphase.F90:46-84)
        if (SPARSE) {
            nb[0] = n;
            a[0] = ldpop(&P.f[0][n]);
            if (MP) b[0] = ldpop(&P.gg[0][n]);
            const int *reci = reinterpret_cast<const int *>(rec);
            const int lane = n & 31;
            const unsigned irr = rec[0].x;
            if (irr == 0) {
                // decode direction d and issue its population loads right away, so that the memory system is busy while
                // the remaining directions are decoded (pure LDS + ALU, ~10 instructions each)
#pragma unroll
                for (int d = 1; d < 19; d++) {
                    const int cq = adj_index_fast(reci, d, lane, P.nAct);
                    nb[d] = cq;
                    a[OPC(d)] = ldpop(&P.f[OPC(d)][cq]);
                    if (MP) b[OPC(d)] = ldpop(&P.gg[OPC(d)][cq]);
                    // software prefetch into L2 for the warps pf_dist nodes ahead: their neighbour indices differ from
                    // ours by ~pf_dist (same offsets), one request per 128-byte line
                    if (PF == 1 && (lane & 15) == 0) {
                        const int pq = min(cq + P.pf_dist, P.nAct - 1);
                        prefetch_l2(P.f[OPC(d)] + pq);
                        if (MP) prefetch_l2(P.gg[OPC(d)] + pq);
                    }
                }
                if (PF == 2) {
                    // metadata only: the adjacency records (five 128-byte lines per warp) and the cell list of the warp
                    // pf_dist nodes ahead -- the two loads every other load of that warp depends on -- become L2 hits
                    const int wn = min((n + P.pf_dist) >> 5, (P.nA - 1) >> 5);
                    if (lane < 5) prefetch_l2(reinterpret_cast<const char *>(P.adj + (size_t)wn * MFLBM_ADJ_REC) + 128 * lane);
                    if (lane == 5) prefetch_l2(P.cellA + min(n + P.pf_dist, P.nA - 1));
                }
                if (PF == 1) {
                    const int wn = min((n + P.pf_dist) >> 5, (P.nA - 1) >> 5);
                    if (lane < 5) prefetch_l2(reinterpret_cast<const char *>(P.adj + (size_t)wn * MFLBM_ADJ_REC) + 128 * lane);
                    if (lane == 5) prefetch_l2(P.cellA + min(n + P.pf_dist, P.nA - 1));
                    if (lane == 6) {
                        const int pq = min(n + P.pf_dist, P.nA - 1);
                        prefetch_l2(P.f[0] + pq);
                        if (MP) prefetch_l2(P.gg[0] + pq);
                    }
                }
            } else {  // warp-uniform and rare: directions with more than five index runs read their indices verbatim
                const int *__restrict__ row = P.adjfull + (size_t)rec[0].y * 32 + lane;
#pragma unroll
                for (int d = 1; d < 19; d++) {
                    int cq = adj_index(rec[1 + 2 * (d - 1)], rec[2 + 2 * (d - 1)], lane, P.nAct);
                    if ((irr >> (d - 1)) & 1u) cq = __ldg(row + 32 * __popc(irr & ((1u << (d - 1)) - 1u)));
                    nb[d] = cq;
                }
#pragma unroll
                for (int d = 1; d < 19; d++) {
                    a[OPC(d)] = P.f[OPC(d)][nb[d]];
                    if (MP) b[OPC(d)] = P.gg[OPC(d)][nb[d]];
                }
            }
        } else {
#pragma unroll
            for (int q = 0; q < 19; q++) {
                const int cq = c - P.g.off(q);
                a[q] = ldpop(&P.f[q][cq]);
                if (MP) b[q] = ldpop(&P.gg[q][cq]);
            }
        }
    } else {  // node-local, direction-swapped slots (MP/Kernel_multiphase.F90:410-448)
        const int cl = SPARSE ? n : c;
#pragma unroll
        for (int q = 0; q < 19; q++) {
            a[q] = ldpop(&P.f[OPC(q)][cl]);
            if (MP) b[q] = ldpop(&P.gg[OPC(q)][cl]);
        }
    }

    collide_node<MP, SPARSE>(P, n, c, wst, a, b);

    if (ODD) {  // push q into slot opc(q) of x + e_q (MP/Kernel_multiphase.F90:318-354)
        const int cl = SPARSE ? n : c;
        stpop(&P.f[0][cl], a[0]);
        if (MP) stpop(&P.gg[0][cl], b[0]);
#pragma unroll
        for (int q = 1; q < 19; q++) {
            const int cq = SPARSE ? nb[q] : c + P.g.off(q);
            stpop(&P.f[OPC(q)][cq], a[q]);
            if (MP) stpop(&P.gg[OPC(q)][cq], b[q]);
        }
    } else {
        const int cl = SPARSE ? n : c;
#pragma unroll
        for (int q = 0; q < 19; q++) {
            stpop(&P.f[q][cl], a[q]);
            if (MP) stpop(&P.gg[q][cl], b[q]);
        }
    }
}

// Threads per block of the collision kernel.  A block's registers are released when its LAST warp exits, so smaller
// blocks leave fewer idle warp slots behind early finishers; measured (r01_v6): singlephase C2 +3 % with one warp per
// block, multiphase C3 -1.6 % (more blocks to launch per byte moved) -> 32 for the singlephase kernels, 128 otherwise.
__host__ __device__ constexpr int collide_block(bool mp) { return mp ? 128 : 32; }
__host__ __device__ constexpr int collide_resident(bool mp) { return (mp ? MFLBM_MP_BLOCKS : MFLBM_SP_BLOCKS) * 128 / collide_block(mp); }

// PF: L2 software prefetch of the sparse odd step (populations, adjacency records and cell list of the warps pf_dist
// nodes ahead)
template <bool MP, bool ODD, bool SPARSE, int PF = 0>
__global__ void __launch_bounds__(collide_block(MP), collide_resident(MP)) k_collide(const Dev P, int k0, int n0, int n1) {
    int c, n = 0, wst = 0;
    __shared__ uint4 s_adj[SPARSE && ODD ? collide_block(MP) / 32 : 1][SPARSE && ODD ? MFLBM_ADJ_REC : 1];
    if (SPARSE) {
        // n0 is rounded down to a multiple of 32 by the launcher so that lane == n & 31 (the adjacency is per warp of
        // 32 consecutive A nodes)
        n = (n0 & ~31) + blockIdx.x * blockDim.x + threadIdx.x;
        if (ODD) {
            // stage this warp's 37 adjacency records (592 contiguous bytes) in shared memory with one coalesced load
            // wave; decoding them straight from global memory makes ptxas chain 18 load->use round trips (measured)
            const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
            if ((n & ~31) < n1) {
                const uint4 *__restrict__ rec = P.adj + (size_t)(n >> 5) * MFLBM_ADJ_REC;
                s_adj[wib][lane] = __ldg(rec + lane);
                if (lane < MFLBM_ADJ_REC - 32) s_adj[wib][32 + lane] = __ldg(rec + 32 + lane);
            }
            __syncwarp();
        }
        if (n < n0 || n >= n1) return;
        c = P.cellA[n];
        grid_dep_wait();  // the adjacency records and the cell index above are static; everything below is not
        if (MP && P.use_tiles) {
            const int all = __ldg(P.tcount + 2);  // the last gradient chain ran over every tile: every warp is active
            const int w = P.wstamp[n >> 5];       // (two independent loads)
            wst = all ? P.wq_stamp : w;
        }
    } else {
        const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
        const int j = blockIdx.y + 1;
        const int k = blockIdx.z + k0;
        if (i > P.g.nx) return;
        c = P.g.cell(i, j, k);
        if (P.walls[c] != 0) return;
        grid_dep_wait();
    }

    node_update<MP, ODD, SPARSE, PF>(P, n, c, wst, SPARSE && ODD ? s_adj[threadIdx.x >> 5] : nullptr);
}

// ordinary launch, or (pdl) with the programmatic-stream-serialization attribute: the kernel's blocks may be scheduled as the
// previous kernel's blocks retire and run up to their grid_dep_wait() -- for a 0.2 ms singlephase step the two launch gaps
// per step are 2 % of the time
template <typename... KArgs, typename... Args>
static void launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args... args) {
    if (!pdl) {
        kernel<<<grid, block, 0, st>>>(args...);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <bool SPARSE>
static void launch_collide_t(mflbm_ctx *c, cudaStream_t st, bool odd, int k0, int k1) {
    const Dev &P = c->d;
    const int blk = collide_block(P.multiphase != 0);
    dim3 block(blk), grid;
    int n0 = 0, n1 = 0;
    if (SPARSE) {
        n0 = c->kstartA[k0];
        n1 = c->kstartA[k1 + 1];
        if (n1 <= n0) return;
        grid = dim3((n1 - (n0 & ~31) + blk - 1) / blk);
    } else {
        grid = dim3((P.g.nx + blk - 1) / blk, P.g.ny, k1 - k0 + 1);
    }
    const bool pf = SPARSE && odd && P.pf_dist > 0;  // L2 software prefetch: a separate instantiation, no dead issue slots
    const bool pdl = c->pdl;
    if (P.multiphase) {
        if (pf && P.pf_mode == 2) launch_pdl(pdl, k_collide<true, true, SPARSE, SPARSE ? 2 : 0>, grid, block, st, P, k0, n0, n1);
        else if (pf) launch_pdl(pdl, k_collide<true, true, SPARSE, SPARSE ? 1 : 0>, grid, block, st, P, k0, n0, n1);
        else if (odd) launch_pdl(pdl, k_collide<true, true, SPARSE>, grid, block, st, P, k0, n0, n1);
        else launch_pdl(pdl, k_collide<true, false, SPARSE>, grid, block, st, P, k0, n0, n1);
    } else {
        if (pf) launch_pdl(pdl, k_collide<false, true, SPARSE, SPARSE ? 1 : 0>, grid, block, st, P, k0, n0, n1);
        else if (odd) launch_pdl(pdl, k_collide<false, true, SPARSE>, grid, block, st, P, k0, n0, n1);
        else launch_pdl(pdl, k_collide<false, false, SPARSE>, grid, block, st, P, k0, n0, n1);
    }
    c->launches++;
}

void launch_collide(mflbm_ctx *c, cudaStream_t st, bool odd, int k0, int k1) {
    if (k1 < k0) return;
    if (c->d.sparse) launch_collide_t<true>(c, st, odd, k0, k1);
    else launch_collide_t<false>(c, st, odd, k0, k1);
}

// ---------------------------------------------------------------------------------------------------
// periodic z on one GPU: the reference's self send/recv through the periodic Cartesian communicator
// (MP/Mpi.F90:101-346 pull, :354-598 push, :608-867 phi).  One launch moves all planes.
// Sparse layout: a copy is skipped when either end is not an active node -- such a slot has no fluid
// consumer (the only consumer of slot q at y is y+e_q), so the skipped copy is unobservable.
// ---------------------------------------------------------------------------------------------------
template <bool MP>
__global__ void k_wrap_z(const Dev P, int push) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y + 1;
    if (i > P.g.nx) return;
    const int nz = P.g.nz;
    int c0 = P.g.cell(i, j, 0), c1 = P.g.cell(i, j, 1), cn = P.g.cell(i, j, nz), cn1 = P.g.cell(i, j, nz + 1);
    grid_dep_wait();
    if (MP) {
#pragma unroll
        for (int kk = 1; kk <= 4; kk++) {
            const double lo = P.phi[P.g.cell(i, j, kk)];
            const double hi = P.phi[P.g.cell(i, j, nz + kk - 4)];
            P.phi[P.g.cell(i, j, kk - 4)] = hi;
            P.phi[P.g.cell(i, j, kk + nz)] = lo;
        }
    }
    bool lo_ok = true, hi_ok = true;  // pull: (1 -> nz+1) and (nz -> 0); push: (nz+1 -> 1) and (0 -> nz)
    if (P.sparse) {
        c0 = P.smap[c0]; c1 = P.smap[c1]; cn = P.smap[cn]; cn1 = P.smap[cn1];
        lo_ok = c1 >= 0 && cn1 >= 0;
        hi_ok = c0 >= 0 && cn >= 0;
    }
    constexpr int qM[5] = {6, 14, 13, 18, 17};  // e_z = -1
    constexpr int qP[5] = {5, 11, 12, 15, 16};  // e_z = +1
#pragma unroll
    for (int m = 0; m < 5; m++) {
        if (!push) {
            if (lo_ok) {
                P.f[qM[m]][cn1] = P.f[qM[m]][c1];
                if (MP) P.gg[qM[m]][cn1] = P.gg[qM[m]][c1];
            }
            if (hi_ok) {
                P.f[qP[m]][c0] = P.f[qP[m]][cn];
                if (MP) P.gg[qP[m]][c0] = P.gg[qP[m]][cn];
            }
        } else {
            if (hi_ok) {
                P.f[qP[m]][cn] = P.f[qP[m]][c0];
                if (MP) P.gg[qP[m]][cn] = P.gg[qP[m]][c0];
            }
            if (lo_ok) {
                P.f[qM[m]][c1] = P.f[qM[m]][cn1];
                if (MP) P.gg[qM[m]][c1] = P.gg[qM[m]][cn1];
            }
        }
    }
}

void launch_wrap_z(mflbm_ctx *c, cudaStream_t st, bool push) {
    const Dev &P = c->d;
    dim3 block(128);
    dim3 grid((P.g.nx + 127) / 128, P.g.ny);
    if (P.multiphase) launch_pdl(c->pdl, k_wrap_z<true>, grid, block, st, P, push ? 1 : 0);
    else launch_pdl(c->pdl, k_wrap_z<false>, grid, block, st, P, push ? 1 : 0);
    c->launches++;
}

// y-periodic lattice: the four phi ghost rows on each side, MP/Mpi.F90:633-655 (pack) and :729-790 (update): rows
// k = 1..nz, and with z periodic too the x edges, which are the same copies applied to the z ghost planes k_wrap_z has
// just filled.  (The populations need no copies: the adjacency wraps, see build_active_host.)
__global__ void k_wrap_y_phi(const Dev P, int k0, int k1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int k = (int)blockIdx.y + k0;
    if (i > P.g.nx || k > k1) return;
    const int ny = P.g.ny;
#pragma unroll
    for (int jj = 1; jj <= 4; jj++) {
        const double lo = P.phi[P.g.cell(i, jj, k)];
        const double hi = P.phi[P.g.cell(i, ny + jj - 4, k)];
        P.phi[P.g.cell(i, jj - 4, k)] = hi;
        P.phi[P.g.cell(i, jj + ny, k)] = lo;
    }
}

void launch_wrap_y_phi(mflbm_ctx *c, cudaStream_t st) {
    const Dev &P = c->d;
    if (!P.multiphase) return;
    const int k0 = P.yw_lo ? -3 : 1, k1 = P.yw_hi ? P.g.nz + 4 : P.g.nz;
    dim3 block(128), grid((P.g.nx + 127) / 128, k1 - k0 + 1);
    k_wrap_y_phi<<<grid, block, 0, st>>>(P, k0, k1);
    c->launches++;
}

// ---------------------------------------------------------------------------------------------------
// sparse layout helpers
// ---------------------------------------------------------------------------------------------------
__global__ void k_fill_smap(const Dev P) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < P.nAct) P.smap[P.cellA[n]] = n;
}

void launch_fill_smap(mflbm_ctx *c, cudaStream_t st) {
    const Dev &P = c->d;
    cudaMemsetAsync(P.smap, 0xff, (size_t)P.g.ntot * sizeof(int), st);
    k_fill_smap<<<(P.nAct + 255) / 256, 256, 0, st>>>(P);
    c->launches++;
}

// caller's (0:nx+1,0:ny+1,0:nz+1) array <-> active-node list
// position of dense cell c in the caller's (0:nx+1,0:ny+1,0:nz+1) array
__device__ __forceinline__ size_t host_pos(const Dev &P, int c) {
    unsigned ix, jy, kz;
    P.g.coords3(c, ix, jy, kz);
    return (size_t)(ix - 3) + (size_t)(P.g.nx + 2) * ((size_t)(jy - 3) + (size_t)(P.g.ny + 2) * (size_t)(kz - 3));
}

// Population array q.  Node entries: n < nAct <-> cell cellA[n].  Link entries: the slot of A node n in direction
// d = opc(q) holds what the reference keeps in array q at the (non-fluid) cell x_n + e_d.
__global__ void k_repack_sparse(const Dev P, double *pdf, double *packed, int to_dev, int q) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= P.nAct) return;
    const int c = P.cellA[n];
    const size_t p = host_pos(P, c);
    if (to_dev) pdf[n] = packed[p];
    else packed[p] = pdf[n];
    if (q == 0 || n >= P.nA) return;
    const int d = OPC(q);
    const uint4 *__restrict__ rec = P.adj + (size_t)(n >> 5) * MFLBM_ADJ_REC;
    const int lane = n & 31;
    if (!((rec[1 + 2 * (d - 1)].x >> lane) & 1u)) return;
    const int slot = adj_lookup(rec, P.adjfull, d, lane, P.nAct);
    const size_t pl = host_pos(P, c + P.g.off(d));
    if (to_dev) pdf[slot] = packed[pl];
    else packed[pl] = pdf[slot];
}

void launch_repack_sparse(mflbm_ctx *c, cudaStream_t st, double *pdf, double *packed, bool to_dev, int q) {
    const Dev &P = c->d;
    k_repack_sparse<<<(P.nAct + 255) / 256, 256, 0, st>>>(P, pdf, packed, to_dev ? 1 : 0, q);
    c->launches++;
}

// NVLink halo exchange, sparse layout: gather/scatter the 5 (+5) z-crossing populations of the boundary
// planes into dense nx*ny send/recv buffers (same plane pairs as MP/Mpi.F90:121-143,244-266,374-396,497-519).
// A non-active source is encoded as a tagged NaN and ignored by the receiver.
#define MFLBM_HOLE 0x7FF8DEADBEEF0001LL
template <bool MP>
__global__ void k_halo_pack(const Dev P, double *buf_lo, double *buf_hi, int push, int unpack) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y + 1;
    if (i > P.g.nx) return;
    const int nz = P.g.nz, nxy = P.g.nx * P.g.ny;
    const int p = (i - 1) + P.g.nx * (j - 1);
    constexpr int qM[5] = {6, 14, 13, 18, 17};
    constexpr int qP[5] = {5, 11, 12, 15, 16};
    // plane used on the low / high side and which population set lives there
    //   pack  pull: lo = qM @ k=1     hi = qP @ k=nz        unpack pull: lo = qP @ k=0    hi = qM @ k=nz+1
    //   pack  push: lo = qP @ k=0     hi = qM @ k=nz+1      unpack push: lo = qM @ k=1    hi = qP @ k=nz
    const bool lo_is_M = (!push && !unpack) || (push && unpack);
    const int klo = lo_is_M ? 1 : 0;
    const int khi = lo_is_M ? nz : nz + 1;
    int clo = P.g.cell(i, j, klo), chi = P.g.cell(i, j, khi);
    if (P.sparse) {
        clo = P.smap[clo];
        chi = P.smap[chi];
    }
#pragma unroll
    for (int fl = 0; fl < (MP ? 2 : 1); fl++) {
        double *const *F = fl == 0 ? P.f : P.gg;
#pragma unroll
        for (int m = 0; m < 5; m++) {
            const int qlo = lo_is_M ? qM[m] : qP[m];
            const int qhi = lo_is_M ? qP[m] : qM[m];
            const int slot = (fl * 5 + m) * nxy + p;
            if (!unpack) {
                if (buf_lo) buf_lo[slot] = clo >= 0 ? F[qlo][clo] : __longlong_as_double(MFLBM_HOLE);
                if (buf_hi) buf_hi[slot] = chi >= 0 ? F[qhi][chi] : __longlong_as_double(MFLBM_HOLE);
            } else {
                if (buf_lo && clo >= 0) {
                    const double v = buf_lo[slot];
                    if (__double_as_longlong(v) != MFLBM_HOLE) F[qlo][clo] = v;
                }
                if (buf_hi && chi >= 0) {
                    const double v = buf_hi[slot];
                    if (__double_as_longlong(v) != MFLBM_HOLE) F[qhi][chi] = v;
                }
            }
        }
    }
}

// phi classes of the ghost planes the halo exchange just filled (k = -3..0 from the z-1 neighbour, nz+1..nz+4 from the
// z+1 neighbour), recorded like the collision kernel records the classes of the fluid nodes: without it the tiles along
// a slab interface would have to count as "unknown" (X) for ever and their gradient chain could never be skipped.
__global__ void __launch_bounds__(128) k_halo_phi_classes(const Dev P, int lo, int hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y + 1;
    if (i > P.g.nx) return;
    const int nz = P.g.nz;
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
        if (lo) {
            const int c = P.g.cell(i, j, -kk);
            tile_record(P, c, P.phi[c]);
        }
        if (hi) {
            const int c = P.g.cell(i, j, nz + 1 + kk);
            tile_record(P, c, P.phi[c]);
        }
    }
}

void launch_halo_phi_classes(mflbm_ctx *c, cudaStream_t st, bool lo, bool hi) {
    const Dev &P = c->d;
    if (!P.multiphase || !P.use_tiles || (!lo && !hi)) return;
    dim3 block(128), grid((P.g.nx + 127) / 128, P.g.ny);
    k_halo_phi_classes<<<grid, block, 0, st>>>(P, lo ? 1 : 0, hi ? 1 : 0);
    c->launches++;
}

void launch_halo_pack(mflbm_ctx *c, cudaStream_t st, double *buf_lo, double *buf_hi, bool push, bool unpack) {
    const Dev &P = c->d;
    dim3 block(128), grid((P.g.nx + 127) / 128, P.g.ny);
    if (P.multiphase) k_halo_pack<true><<<grid, block, 0, st>>>(P, buf_lo, buf_hi, push, unpack);
    else k_halo_pack<false><<<grid, block, 0, st>>>(P, buf_lo, buf_hi, push, unpack);
    c->launches++;
}

}  // namespace mflbm
