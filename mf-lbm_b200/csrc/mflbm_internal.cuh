// mflbm_internal.cuh -- device-side layout, parameter block and context of the MF-LBM hot path.
//
// HBM layout (DESIGN.md "Data layout"): every 3-D field lives on ONE padded grid so that a single
// linear cell index addresses all arrays:
//     cell(i,j,k) = base + (i-1) + sx*(j+3) + sxy*(k+3),   i in [-3,nx+4], j in [-3,ny+4], k in [-3,nz+4]
// with sx = round_up(nx+8,16) doubles (rows start 128-byte aligned at i=1; the 4 low-x ghosts of a row
// sit in the tail padding of the previous row), sxy = sx*(ny+8), base = 16.
// Structure of arrays: one such grid per population (f0..f18, g0..g18), phi, and the optional
// derived fields.  int32 cell indices are sufficient for one GPU (<= 2^31 cells is checked at create).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/mflbm.h"

namespace mflbm {

// D3Q19 lattice of the reference, MP/Module.F90:111-114
__host__ __device__ constexpr int EX(int q) {
    constexpr int t[19] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
    return t[q];
}
__host__ __device__ constexpr int EY(int q) {
    constexpr int t[19] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
    return t[q];
}
__host__ __device__ constexpr int EZ(int q) {
    constexpr int t[19] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
    return t[q];
}
__host__ __device__ constexpr int OPC(int q) {
    constexpr int t[19] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
    return t[q];
}

struct Grid {
    int nx, ny, nz;
    int sx;    // x stride (doubles), multiple of 16
    int sxy;   // sx*(ny+8)
    int base;  // 16
    int ntot;  // allocated cells per field
    __host__ __device__ __forceinline__ int cell(int i, int j, int k) const { return base + (i - 1) + sx * (j + 3) + sxy * (k + 3); }
    __host__ __device__ __forceinline__ int off(int q) const { return EX(q) + sx * EY(q) + sxy * EZ(q); }
    __host__ __device__ __forceinline__ int cell2(int i, int j) const { return base + (i - 1) + sx * (j + 3); }  // 2-D plane fields
    int plane_cells() const { return sxy; }
    // exact division of r < 2^31 by sx / sxy with a multiply-high (Granlund-Montgomery: m = ceil(2^(31+L)/d),
    // L = ceil(log2 d), fits 32 bits): the generic 32-bit division costs ~20 instructions and two MUFU round trips
    unsigned m_sx, h_sx, m_sxy, h_sxy;  // magic multiplier and post-shift (applied to the high word)
    void set_magic() {
        auto mk = [](unsigned d, unsigned &m, unsigned &h) {
            unsigned L = 0;
            while ((1ull << L) < d) L++;
            const unsigned long long p2 = 1ull << (31 + L);
            m = (unsigned)((p2 + d - 1) / d);
            h = L - 1;  // (r*m) >> (31+L) == umulhi(r,m) >> (L-1);  d >= 2 always (sx >= 16)
        };
        mk((unsigned)sx, m_sx, h_sx);
        mk((unsigned)sxy, m_sxy, h_sxy);
    }
    __host__ __device__ __forceinline__ static unsigned mulhi(unsigned a, unsigned b) {
#ifdef __CUDA_ARCH__
        return __umulhi(a, b);
#else
        return (unsigned)(((unsigned long long)a * b) >> 32);
#endif
    }
    // linear cell c -> (i+3, j+3, k+3)
    __host__ __device__ __forceinline__ void coords3(int c, unsigned &ix, unsigned &jy, unsigned &kz) const {
        const unsigned r = (unsigned)(c - (base - 4));  // (i+3) + sx*(j+3) + sxy*(k+3)
        kz = mulhi(r, m_sxy) >> h_sxy;
        const unsigned r2 = r - kz * (unsigned)sxy;
        jy = mulhi(r2, m_sx) >> h_sx;
        ix = r2 - jy * (unsigned)sx;
    }
    // tile of 8x4x4 cells containing linear cell c (ntx = sx/8 tiles per row, nty tiles per plane column)
    __host__ __device__ __forceinline__ int tile_of(int c, int ntx, int nty) const {
        unsigned ix, jy, kz;
        coords3(c, ix, jy, kz);
        return (int)((ix >> 3) + (unsigned)ntx * ((jy >> 2) + (unsigned)nty * (kz >> 2)));
    }
    // first cell of the contiguous storage of plane k (includes the row paddings): cell(-3,-3,k)
    __host__ __device__ __forceinline__ int plane_begin(int k) const { return base - 4 + sxy * (k + 3); }
};

// Kernel parameter block (by value; < 4 KB)
struct Dev {
    Grid g;
    double *f[19];
    double *gg[19];  // g0..g18 (fluid 2)
    double *phi, *phi_old;
    double *cn_x, *cn_y, *cn_z, *c_norm, *curv;
    double *u, *v, *w, *rho;
    int8_t *walls;
    double *w_in, *f_convec, *g_convec, *phi_convec;  // plane fields; *_convec hold 19 planes of sx*(ny+8)
    // solid / fluid boundary node lists (device SoA, built at upload)
    int *solid_cell;
    unsigned *solid_mask;  // bits 1..18: fluid neighbour e_n present (list order == increasing n); bit 31: inside 0..n+1 box
    double *solid_law;     // la_weight
    int num_solid;
    int *fluid_cell;
    double *fluid_nw;  // structure of arrays [5][num_fluid]: nwx, nwy, nwz, cos(theta), sin(theta)
    int num_fluid;
    // sparse storage of the populations (DESIGN.md "Sparse population storage"): the 38 PDF arrays hold only
    // ACTIVE nodes = fluid nodes of 1..n (A, processed by the collision kernel, raster order k,j,i) followed by
    // storage-only nodes (S: solid-boundary / ghost-plane nodes that are a D3Q19 neighbour of an A node).
    // Scalar fields (phi, cn_*, c_norm, curv, walls, u,v,w,rho) stay on the dense padded grid.
    int sparse;       // 0: PDFs on the dense grid, 1: PDFs on the active-node list
    int nA, nAct;     // number of A nodes / of all nodes with an index (A + zone S nodes)
    int *cellA;       // [nAct] dense cell of each node
    // Adjacency of the odd step, compressed per warp (32 consecutive A nodes): 37 uint4 per warp,
    //   [0]        header {irrmask, fullbase, 0, 0}
    //   [1+2(d-1)] r0 = {smask, jmask, lbase, B1}   for direction d = 1..18
    //   [2+2(d-1)] r1 = {B2, B3, B4, B5}
    // smask: lanes whose neighbour x+e_d has no node index -> compact link slot nAct + lbase + (rank of the lane among
    //        the warp's link lanes) = nAct + lbase + lane - rank, rank = number of non-link lanes below the lane.
    // jmask: non-link lanes that start a new index run; the other non-link lanes continue the previous non-link lane
    //        by +1.  Run r (1-based, r = popc(jmask at or below the lane)) stores B_r = (index of its first lane) - (rank
    //        of its first lane), so that every lane of the run decodes as B_r + rank: two popc, one shared-memory read.
    // A direction with more than five runs in the warp is IRREGULAR (bit d-1 of irrmask): its 32 indices are stored
    // verbatim in adjfull, row fullbase + popc(irrmask below d).  72 B/node of int32 indices become ~18.5 B/node.
    uint4 *adj;       // [ceil(nA/32)][37]
    int *adjfull;     // [rows][32]
    int pf_dist;      // L2 software-prefetch distance of the odd step in nodes (0 = off)
    int pf_mode;      // 1: populations + metadata (singlephase default), 2: adjacency records + cell list only
    int nlink[19];    // number of link slots of direction d (stored behind the node entries of population array opc(d))
    int *smap;        // [ntot] dense cell -> active index, -1 if not active
    int *gcell;       // [nG] non-solid cells of the (-1:n+2)^3 box: where the colour gradient is evaluated (grouped by tile
                      // when use_tiles, raster order otherwise)
    int *gcell_r;     // the same cells in the order of the FLAT sweeps (use_tiles only): bricks of flat_bx x flat_by x flat_bz
                      // cells, raster order (k, j, i) inside a brick and from brick to brick; one brick as wide as the lattice
                      // and one row high = plain raster order.  Consecutive threads then work on x-runs of neighbouring rows
                      // and planes, so the 18-neighbour stencils of a block share their operands in L1.
    int *gk5_r;       // [nG] parallel to gcell_r: index of the cell in the flat-order fluid boundary list, -1 if it has no
                      // entry: the flat K4 sweep then applies K5 (geometric wetting) on the spot and the K5 sweep is skipped
    int *gk5;         // the same for gcell and the fluid boundary list grouped by tile (tile-driven K4)
    int nG;
    // the node lists and the active nodes in flat order (see gcell_r), for the flat sweeps of K3 / K6 / K7
    int *solid_cell_r, *fluid_cell_r;
    int *aorder, *acell;  // K7 + packing in brick order (MFLBM_BRICK7): active index / dense cell of the e-th fluid node
    unsigned *solid_mask_r;
    double *solid_law_r, *fluid_nw_r;
    int full_curv;    // 1: curvature at all nodes like the reference (MP/Phase_gradient.F90:121); 0: fluid nodes only
    // sparse multiphase layout: what the collision kernel needs of the colour gradient, PACKED by active index n so that a
    // warp reads four coalesced 256-byte rows: G[0..2] = interface normal n (after K4 + K5), G[3] = 0.5*gamma*curv*|grad phi|
    // (the reference's `tmp`, MP/Kernel_multiphase.F90:118, evaluated in its order).  Written after every gradient chain by
    // k_gradient_pack (K7 lives there: 54 gathers per interface node with few registers and full occupancy, instead of
    // inside the 128-register collision kernel); exact zeros wherever |grad phi| = 0.
    double *G[4];
    // phi-uniformity tiles (sparse multiphase layout, DESIGN.md "Quiet tiles"): the padded grid is cut into tiles of
    // 8x4x4 cells; the collision kernel records which phi classes occur among the fluid nodes of each tile
    // (P: |phi-1|<=1e-7, M: |phi+1|<=1e-7, X: anything else).  A tile whose 27-tile neighbourhood shows one single
    // class for two consecutive steps is QUIET: every colour gradient there is below the reference's 1e-6 cut-off,
    // i.e. exactly zero, so the gradient chain K3..K6 and the c_norm read of the collision kernel are skipped.
    int use_tiles;
    int lazy_ok;               // 1: the dense n / |grad phi| arrays hold the previous evaluation, cells that stay below the cut-off
                               // need not be rewritten (0 with the hybrid chain, see mflbm_ctx::march_hybrid)
    int k4_smem;               // 1: K4 on active tiles stages phi through shared memory (k_gradient_tiles)
    int bc_lo_dyn, bc_hi_dyn;  // 1: the phi ghost planes below k=1 / above k=nz are rewritten every step by an inlet /
                               // outlet kernel or by the halo exchange, and whoever writes them records their phi
                               // classes like the collision kernel does (tile_record)
    int ntx, nty, ntz, ntiles;
    int tile_cur;            // which tcls / tU buffer the current step writes
    unsigned char *tcls[2];  // per-step class bits of the fluid nodes (atomicOr by k_collide)
    unsigned char *tstat;    // class bits of cells whose phi never changes (ghost / unlisted cells), X on z-ghost tiles
    unsigned char *tU[2];    // class bits OR-ed over the 27-tile neighbourhood, this step and the previous one
    unsigned char *tquiet;   // 1: tile is quiet
    int *tg_start, *ts_start, *tf_start;  // [ntiles+1] CSR ranges of gcell / the solid list / the fluid list (sorted by tile)
    int *tact, *tk3;         // [ntiles] active tiles (K4,K5,K6) / tiles within one tile of an active tile (K3), per step
    int *tcount;             // [8]: [0], [1] lengths of tact, tk3; [2] = 1: most tiles are active, the chain of this step ran
                             // over the whole lists (flat sweep) and every warp counts as active; [3] = 1: tile-driven
                             // pass selected; [4] active tiles found by the early half of a speculative step; [5], [6]
                             // layer range of the active tiles (ntz - min tz, max tz + 1; 0 = none); [7] scratch sink
    int *tk3stamp;           // [ntiles] step stamp guarding the tk3 append
    // per warp of 32 consecutive A nodes: stamp of the last tile update that found one of its nodes in an ACTIVE tile.
    // The collision kernel reads this one warp-uniform word (address known from n alone) instead of chaining
    // cellA -> tile index -> tquiet -> c_norm; a warp whose stamp is stale skips the c_norm read altogether.
    int *wstamp;             // [ceil(nA/32)]
    // march kernel (march.cuh): one code word per cell of the padded grid (type + neighbour mask / fluid-list index), and
    // per (column, chunk of march_lz planes) the stamp of the last tile update that found an active tile there
    unsigned *mcode;         // [ntot]
    int *mflag;              // [mcols_x * mcols_y * mchunks]
    int *mlist;              // work items (column + mcols_x * mcols_y * chunk) holding an active tile, appended by the tile
                             // update (tcount[10] of them; tcount[11] is the ticket counter of the march kernel)
    int mcols_x, mcols_y, mchunks, march_lz;
    int wq_stamp;            // stamp written by the last tile update (tile_stamp_warps, run by the K4 launch)
    int wq_all;              // 1: every warp counts as active (after a reset, until the next tile update)
    // scalars
    int jper, kper;   // periodic indicators (jper: y wrap of the populations lives in the adjacency, phi by k_wrap_y_phi)
    int yw_lo, yw_hi; // y-periodic: the z ghost planes below k=1 / above k=nz are exchanged (periodic wrap or neighbour slab),
                      // so the y wrap applies to them too (the reference's x-edge exchange, MP/Mpi.F90:184-207, :656-670)
    int multiphase, mrt;
    double la_nui1, la_nui2, gamma, beta, force_Z, phi_inlet, sa_inject, relaxation, uin_avg, rho_in, rho_out;
    double s_e, s_e2, s_q, s_nu, s_pi, s_t;
    double rk_weight2;  // 1/sqrt(2)/36 evaluated on the host like MP/Module.F90:225
};

#ifdef __CUDACC__
// phi class of one cell into the class buffer of the current step (see Dev::tcls): P |phi-1|<=1e-7, M |phi+1|<=1e-7, X else
__device__ __forceinline__ void tile_record(const Dev &P, int c, double phi) {
    const int tile = P.g.tile_of(c, P.ntx, P.nty);
    const unsigned bits = fabs(phi - 1.0) <= 1e-7 ? 1u : (fabs(phi + 1.0) <= 1e-7 ? 2u : 4u);
    const unsigned key = ((unsigned)tile << 3) | bits;
    // consecutive columns of a row mostly share (tile, class): the first lane of every run of equal keys reports
    const unsigned act = __activemask();
    const unsigned prev = __shfl_up_sync(act, key, 1);
    const int lane = threadIdx.x & 31;
    if (lane == 0 || !((act >> (lane - 1)) & 1u) || prev != key)
        atomicOr((unsigned *)(P.tcls[P.tile_cur] + (tile & ~3)), bits << (8 * (tile & 3)));
}
#endif

#define MFLBM_MARCH_TX 32  // column of cells one block of the march kernel owns (march.cuh)
#define MFLBM_MARCH_TY 16
#define MFLBM_ADJ_REC 37  // uint4 records per warp
#define MFLBM_SAT_SEG 4   // blocks per z slice of k_saturation (2 * nz * MFLBM_SAT_SEG partial sums fit red_len)

// index of the neighbour of this lane's node in the (regular) direction the records r0,r1 describe (see Dev::adj).
// Pure ALU on purpose: any load in here would be chained 18 times behind the previous direction's latency.
__host__ __device__ __forceinline__ int adj_index(const uint4 r0, const uint4 r1, int lane, int nAct) {
#ifdef __CUDA_ARCH__
#define MFLBM_POPC(x) __popc(x)
#define MFLBM_CLZ(x) __clz(x)
#else
#define MFLBM_POPC(x) __builtin_popcount(x)
#define MFLBM_CLZ(x) __builtin_clz(x)
#endif
    const unsigned le = 0xffffffffu >> (31 - lane);  // lanes <= lane
    const bool link = (r0.x >> lane) & 1u;
    const int rank = MFLBM_POPC(~r0.x & (le >> 1));  // non-link lanes below this lane
    const int r = MFLBM_POPC(r0.y & le);             // run of this lane, 1-based (0 only for link lanes before any run)
    const int base = r <= 1 ? (int)r0.w : r == 2 ? (int)r1.x : r == 3 ? (int)r1.y : r == 4 ? (int)r1.z : (int)r1.w;
    return link ? nAct + (int)r0.z + lane - rank : base + rank;
}

// The same for a REGULAR direction d, reading straight from the warp's records viewed as 32-bit words (the collision
// kernel's fast path: two popc, one indexed read).  Word layout of direction d: w0 = 4 + 8 (d - 1):
// {smask, jmask, lbase, B1, B2, B3, B4, B5}; run == 0 (a link lane before any run) reads lbase, which is then unused.
__host__ __device__ __forceinline__ int adj_index_fast(const int *reci, int d, int lane, int nAct) {
    const int w0 = 4 + 8 * (d - 1);
    const unsigned smask = (unsigned)reci[w0], jmask = (unsigned)reci[w0 + 1];
    const unsigned le = 0xffffffffu >> (31 - lane), lt = le >> 1;
    const int rank = MFLBM_POPC(~smask & lt);
    const int run = MFLBM_POPC(jmask & le);
    const int base = reci[w0 + 2 + run];
    return ((smask >> lane) & 1u) ? nAct + reci[w0 + 2] + lane - rank : base + rank;
}

// host-side / slow-path lookup including the irregular rows (the collision kernel has its own batched version)
__host__ __device__ __forceinline__ int adj_lookup(const uint4 *rec /* this warp's 37 records */, const int *adjfull, int q,
                                                   int lane, int nAct) {
    const uint4 hdr = rec[0];
    if ((hdr.x >> (q - 1)) & 1u) return adjfull[((size_t)hdr.y + MFLBM_POPC(hdr.x & ((1u << (q - 1)) - 1u))) * 32 + lane];
    return adj_index(rec[1 + 2 * (q - 1)], rec[2 + 2 * (q - 1)], lane, nAct);
}

}  // namespace mflbm

struct ncclComm;
struct NcclApi;

struct mflbm_ctx {
    mflbm_config cfg;
    mflbm::Dev d;
    int device;
    cudaStream_t s_main, s_halo;
    cudaEvent_t ev_t0, ev_t1, ev_slab, ev_halo, ev_fork;
    cudaEvent_t ev_phi;      // phi of the fluid nodes of the current step is final (recorded right after the collision kernels)
    bool ev_phi_valid;
    // asynchronous output staging (mflbm_output_begin / _end): packed device copies, pinned host mirrors, copy stream
    cudaStream_t s_copy;
    cudaEvent_t ev_out, ev_out_done;
    double *out_dev[5], *out_host[5];  // phi, u, v, w, rho in the caller's Fortran layout
    size_t out_elems[5];
    int out_pending;                   // field mask of the staged output in flight (0: none)
    // checkpoint staging (mflbm_checkpoint_begin / _fetch / _end)
    // speculative early gradient chain (step_impl): layer range of the active tiles as of a recent step, read back
    // asynchronously (heuristic only -- a wrong guess costs time, never correctness)
    cudaStream_t s_chain;
    cudaEvent_t ev_near, ev_chain, ev_sum[2];
    int *sum_host;                     // pinned, 2 x 8 ints (copies of Dev::tcount)
    long long sum_step[2];             // step counter at which each slot was requested (-1: never)
    long long step_count;
    int sum_next;                      // slot of the next request
    int spec_enabled;                  // MFLBM_NO_SPEC=1 switches it off
    long long spec_steps;              // steps that ran speculatively (mflbm_tile_stats-style diagnostics)
    int ckpt_mode;                     // 0 none, 1 staged (device snapshot in ckpt_*), 2 direct (context frozen)
    double *ckpt_f[19], *ckpt_g[19], *ckpt_phi, *ckpt_fc, *ckpt_gc, *ckpt_pc;
    cudaEvent_t ev_ckpt;
    std::vector<void *> allocs;
    long long bytes;
    long long adj_bytes;  // size of the compressed adjacency
    long long launches;
    std::string err;
    // reduction scratch
    double *red_dev;   // device
    double *red_host;  // pinned
    int red_len;
    // staging buffer for layout conversion on upload/download
    double *stage;
    size_t stage_bytes;
    // NCCL
    NcclApi *nccl;
    ncclComm *comm;
    int peer_lo, peer_hi;  // ranks of z-1 / z+1 neighbours (periodic ring)
    bool open_z;
    bool macro_alloc;
    bool pdf_alloc;
    // march kernel (kernels_march.cu): fused colour-gradient chain of the sparse multiphase layout
    // streamed steps (mflbm_step_streamed): double-buffered inlet profile and result slots
    double *win_dev[2], *win_stage[2];  // device w_in alternates (win_dev[0] is the upload's array) / packed host-layout staging
    double *res_dev[2], *res_host[2];   // partial sums of cal_saturation, device / pinned host
    cudaEvent_t ev_win[2], ev_res[2], ev_step[2];
    int res_np[2];
    long long stream_count;             // streamed steps so far
    bool pdl;                       // collision / wrap kernels launched with programmatic stream serialization (MFLBM_PDL)
    int flat_bx, flat_by, flat_bz;  // brick of the flat-sweep order (MFLBM_BRICK="bx,by,bz")
    int k7_bx, k7_by, k7_bz;        // brick order of K7 + packing alone (MFLBM_BRICK7; 0,1,1 = the node order itself)
    bool march_on;            // selected for this context (MFLBM_MARCH=1 or 2; otherwise the list kernels)
    bool march_hybrid;        // MFLBM_MARCH=2: the fused kernel only for the work items around the active tiles; when most
                              // tiles are active the flat sweeps of the list kernels run (faster there), without their lazy
                              // shortcuts because the dense arrays are not kept current by the fused kernel
    bool march_ready;         // cell codes built for the current walls / node lists
    int march_reject;         // why the node lists were not accepted (bit mask, 0 = accepted)
    int march_lz_flat;        // planes per work item of the sweeps over everything
    bool cn_dense_valid;      // the dense n / |grad phi| arrays hold the current values (the march kernel does not write them)
    int tile_stamp;
    bool tiles_static_ready;  // tstat computed for the current phi / wall / node-list upload
    bool solid_phi_stale;     // phi on solid boundary nodes of quiet tiles was not refreshed by the last gradient chain
    double *halo_buf[4];  // sparse NCCL exchange: send_lo, send_hi, recv_lo, recv_hi
    std::vector<int> kstartA;
    bool prof;                         // per-launch event timing of the collision kernel
    std::vector<cudaEvent_t> prof_ev;  // pairs (start, stop)
    size_t prof_used;
    double prof_ms;
    long long prof_launches;  // sparse: first A index of plane k (k=1..nz+1), for slab launches
};

namespace mflbm {
// launchers implemented in the .cu files
void launch_collide(mflbm_ctx *c, cudaStream_t st, bool odd, int k0, int k1);
void launch_fill_smap(mflbm_ctx *c, cudaStream_t st);
void launch_repack_sparse(mflbm_ctx *c, cudaStream_t st, double *pdf, double *packed, bool to_dev, int q);
void launch_halo_pack(mflbm_ctx *c, cudaStream_t st, double *buf_lo, double *buf_hi, bool push, bool unpack);
void launch_halo_phi_classes(mflbm_ctx *c, cudaStream_t st, bool lo, bool hi);
void launch_color_gradient(mflbm_ctx *c, cudaStream_t st, bool stepping = false);
void launch_phi_solid_refresh(mflbm_ctx *c, cudaStream_t st);
void launch_tiles_reset(mflbm_ctx *c, cudaStream_t st);
int tiles_prepare(mflbm_ctx *c, cudaStream_t st);
void launch_curvature(mflbm_ctx *c, cudaStream_t st);
void launch_gradient_pack(mflbm_ctx *c, cudaStream_t st);
void launch_build_gk5(mflbm_ctx *c, cudaStream_t st, int *map, int *dup, bool flat);
int march_prepare(mflbm_ctx *c, cudaStream_t st);
void launch_march(mflbm_ctx *c, cudaStream_t st, int mode, int stamp);
void launch_dense_gradient(mflbm_ctx *c, cudaStream_t st);
long long chain_selfcheck(mflbm_ctx *c, cudaStream_t st);
void launch_chain_early(mflbm_ctx *c, cudaStream_t st, int tz_lo, int tz_hi);
void launch_chain_late(mflbm_ctx *c, cudaStream_t st, int tz_lo, int tz_hi);
void launch_bc(mflbm_ctx *c, cudaStream_t st, bool after_odd);
void launch_wrap_z(mflbm_ctx *c, cudaStream_t st, bool push);
void launch_wrap_y_phi(mflbm_ctx *c, cudaStream_t st);
void launch_macro(mflbm_ctx *c, cudaStream_t st);
void launch_monitor(mflbm_ctx *c, cudaStream_t st, double *out /*device, (10)*nz*/);
int launch_saturation(mflbm_ctx *c, cudaStream_t st, double *out /*device, 2 rows of partial sums*/);
void launch_breakthrough(mflbm_ctx *c, cudaStream_t st, double *out);
void launch_steady_phasefield(mflbm_ctx *c, cudaStream_t st, double *out /*2*nz*/);
void launch_steady_cappres(mflbm_ctx *c, cudaStream_t st, double *out /*5*nz*/);
void launch_repack(mflbm_ctx *c, cudaStream_t st, double *grid, double *packed, int ghost, int nplanes_z, int kbase, bool to_grid);
void launch_repack_i8(mflbm_ctx *c, cudaStream_t st, int8_t *grid, int8_t *packed, int ghost, bool to_grid);
}  // namespace mflbm
