// mflbm_api.cu -- the extern "C" boundary declared in include/mflbm.h, the per-step schedule
// (main_iteration_kernel, MP/Main_multiphase.F90:341-486 ; SP/Main.F90:291-422) and the NVLink z-halo
// exchange that replaces the host-staged MPI of MP/Mpi.F90 / SP/Mpi.F90.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "mflbm_internal.cuh"

using namespace mflbm;

// ---------------------------------------------------------------------------------------------------
// NCCL is bound at run time (dlopen) so that single-GPU users and CPU-only symbol checks do not need it.
// ---------------------------------------------------------------------------------------------------
struct NcclApi {
    void *h;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
};

static NcclApi *load_nccl(std::string &err) {
    static NcclApi api;
    static bool tried = false, ok = false;
    if (tried) return ok ? &api : nullptr;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.h) break;
    }
    if (!api.h) {
        err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
        return nullptr;
    }
#define SYM(field, name)                                     \
    *(void **)(&api.field) = dlsym(api.h, name);             \
    if (!api.field) {                                        \
        err = std::string("NCCL symbol missing: ") + name;   \
        return nullptr;                                      \
    }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    ok = true;
    return &api;
}

static thread_local std::string g_err;  // errors before a context exists

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            char b_[512];                                                                               \
            snprintf(b_, sizeof b_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            if (ctx) ctx->err = b_;                                                                     \
            g_err = b_;                                                                                 \
            return MFLBM_ERR_CUDA;                                                                      \
        }                                                                                               \
    } while (0)

#define NC(call)                                                                                             \
    do {                                                                                                     \
        ncclResult_t r_ = (call);                                                                            \
        if (r_ != ncclSuccess) {                                                                             \
            char b_[512];                                                                                    \
            snprintf(b_, sizeof b_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, ctx->nccl->GetErrorString(r_)); \
            ctx->err = b_;                                                                                   \
            return MFLBM_ERR_NCCL;                                                                           \
        }                                                                                                    \
    } while (0)

static int fail(mflbm_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg;
    g_err = msg;
    return code;
}

template <typename T>
static int dev_alloc(mflbm_ctx *ctx, T **p, size_t n, bool zero = true) {
    void *q = nullptr;
    CU(cudaMalloc(&q, n * sizeof(T)));
    if (zero) CU(cudaMemset(q, 0, n * sizeof(T)));
    ctx->allocs.push_back(q);
    ctx->bytes += (long long)(n * sizeof(T));
    *p = (T *)q;
    return 0;
}

extern "C" const char *mflbm_version(void) { return "mflbm-b200 0.1 (sm_100a)"; }

extern "C" const char *mflbm_last_error(const mflbm_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

extern "C" int mflbm_nccl_unique_id(unsigned char id[128]) {
    std::string err;
    NcclApi *api = load_nccl(err);
    if (!api) return fail(nullptr, MFLBM_ERR_NCCL, err);
    ncclUniqueId u;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    if (api->GetUniqueId(&u) != ncclSuccess) return fail(nullptr, MFLBM_ERR_NCCL, "ncclGetUniqueId failed");
    memcpy(id, &u, 128);
    return MFLBM_OK;
}

extern "C" int mflbm_create(const mflbm_config *cfg, mflbm_ctx **out) {
    mflbm_ctx *ctx = nullptr;
    if (!cfg || !out) return fail(nullptr, MFLBM_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->struct_size != (int)sizeof(mflbm_config)) return fail(nullptr, MFLBM_ERR_ARG, "mflbm_config.struct_size mismatch");
    if (cfg->nx < 1 || cfg->ny < 1 || cfg->nz < 2) return fail(nullptr, MFLBM_ERR_ARG, "bad lattice dimensions");
    if (cfg->npz < 1 || cfg->idz < 0 || cfg->idz >= cfg->npz) return fail(nullptr, MFLBM_ERR_ARG, "bad idz/npz");
    if (cfg->jper != 0 && cfg->porous_plate_cmd != 0)
        return fail(nullptr, MFLBM_ERR_ARG, "y-periodic domains (jper=1) need the sparse population layout, the porous plate the dense one");
    if (cfg->npz > 1 && !cfg->use_nccl) return fail(nullptr, MFLBM_ERR_ARG, "npz>1 requires use_nccl=1");
    if (cfg->npz > 1 && cfg->solver == MFLBM_SOLVER_MULTIPHASE && cfg->iz_async < 4)
        return fail(nullptr, MFLBM_ERR_ARG, "multiphase needs iz_async >= 4 (overlap_phi)");
    if (cfg->npz > 1 && cfg->nz < 2 * (cfg->iz_async > 0 ? cfg->iz_async : 1))
        return fail(nullptr, MFLBM_ERR_ARG, "nz_local must be >= 2*iz_async");
    const long long sx = ((long long)cfg->nx + 8 + 15) / 16 * 16;
    const long long sxy = sx * (cfg->ny + 8);
    const long long ntot = 16 + sxy * (cfg->nz + 8) + 16;
    if (ntot >= (1LL << 31)) return fail(nullptr, MFLBM_ERR_ARG, "slab too large for int32 cell indices");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, MFLBM_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
    ctx = new mflbm_ctx();
    ctx->cfg = *cfg;
    ctx->bytes = 0;
    ctx->launches = 0;
    ctx->nccl = nullptr;
    ctx->comm = nullptr;
    ctx->macro_alloc = false;
    ctx->pdf_alloc = false;
    ctx->tiles_static_ready = false;
    ctx->solid_phi_stale = false;
    // fused colour-gradient chain (kernels_march.cu) on the sparse multiphase layout.  MEASURED (r02_march1 / march2, B200, C3
    // with random phi): bit-exact, DRAM traffic 4.6 GB per evaluation against 15 GB for the five list kernels, and still
    // slower -- 7.3 ms against 5.1 ms: 3.3 G warp instructions at 1.6 per cycle with the 16 warps per SM its 180 KB of
    // shared memory leave room for (stalls: fixed-latency dependencies and the two barriers per plane, no memory stall at
    // all).  The chain is bound by instruction latency, not by bytes.  Opt-in: MFLBM_MARCH=1.
    ctx->march_on = getenv("MFLBM_MARCH") && atoi(getenv("MFLBM_MARCH")) != 0;
    ctx->march_hybrid = ctx->march_on && atoi(getenv("MFLBM_MARCH")) == 2;
    ctx->march_ready = false;
    ctx->march_reject = 0;
    for (int b = 0; b < 2; b++) {
        ctx->win_dev[b] = ctx->win_stage[b] = ctx->res_dev[b] = ctx->res_host[b] = nullptr;
        ctx->ev_win[b] = ctx->ev_res[b] = ctx->ev_step[b] = nullptr;
        ctx->res_np[b] = 0;
    }
    ctx->stream_count = 0;
    // order of the flat sweeps of the gradient chain (Dev::gcell_r): bricks of bx x by x bz cells; bx = 0: as wide as the lattice
    ctx->flat_bx = 0; ctx->flat_by = 1; ctx->flat_bz = 1;
    ctx->pdl = getenv("MFLBM_PDL") && atoi(getenv("MFLBM_PDL")) != 0;
    ctx->k7_bx = 0; ctx->k7_by = 1; ctx->k7_bz = 1;
    if (const char *b = getenv("MFLBM_BRICK7")) {
        int x = 0, y = 1, z = 1;
        if (sscanf(b, "%d,%d,%d", &x, &y, &z) == 3 && x >= 0 && y >= 1 && z >= 1) { ctx->k7_bx = x; ctx->k7_by = y; ctx->k7_bz = z; }
    }
    if (const char *b = getenv("MFLBM_BRICK")) {
        int x = 0, y = 1, z = 1;
        if (sscanf(b, "%d,%d,%d", &x, &y, &z) == 3 && x >= 0 && y >= 1 && z >= 1) { ctx->flat_bx = x; ctx->flat_by = y; ctx->flat_bz = z; }
    }
    ctx->march_lz_flat = getenv("MFLBM_MARCH_LZ") ? std::max(1, atoi(getenv("MFLBM_MARCH_LZ"))) : 64;
    ctx->cn_dense_valid = true;
    ctx->tile_stamp = 0;
    ctx->prof = false;
    ctx->prof_used = 0;
    ctx->prof_ms = 0;
    ctx->prof_launches = 0;
    ctx->halo_buf[0] = ctx->halo_buf[1] = ctx->halo_buf[2] = ctx->halo_buf[3] = nullptr;
    ctx->ckpt_mode = 0;
    ctx->ev_ckpt = nullptr;
    ctx->s_chain = nullptr;
    ctx->ev_near = ctx->ev_chain = ctx->ev_sum[0] = ctx->ev_sum[1] = nullptr;
    ctx->sum_host = nullptr;
    ctx->sum_step[0] = ctx->sum_step[1] = -1;
    ctx->step_count = 0;
    ctx->sum_next = 0;
    ctx->spec_steps = 0;
    // MEASURED (r02_m8 / m11, B200): exact, but no gain -- C3 5.98 ms/step with it against 5.91 without, the 1536x1536x192
    // slab 25.1 against 24.4 (under sw_power_cap): the collision kernel holds every register of an SM, the chain kernels
    // only trickle in as its blocks retire and slow it down by what they gain.  Off unless MFLBM_SPEC=1.
    ctx->spec_enabled = (getenv("MFLBM_SPEC") && atoi(getenv("MFLBM_SPEC"))) ? 1 : 0;
    for (int q = 0; q < 19; q++) ctx->ckpt_f[q] = ctx->ckpt_g[q] = nullptr;
    ctx->ckpt_phi = ctx->ckpt_fc = ctx->ckpt_gc = ctx->ckpt_pc = nullptr;
    ctx->stage = nullptr;
    ctx->stage_bytes = 0;
    ctx->red_dev = nullptr;
    ctx->red_host = nullptr;
    memset(&ctx->d, 0, sizeof(ctx->d));
#define CREATE_FAIL(code)   \
    do {                    \
        g_err = ctx->err;   \
        mflbm_destroy(ctx); \
        return code;        \
    } while (0)
    if (cfg->device >= 0) {
        if (cudaSetDevice(cfg->device) != cudaSuccess) {
            ctx->err = "cudaSetDevice failed";
            CREATE_FAIL(MFLBM_ERR_CUDA);
        }
    }
    cudaGetDevice(&ctx->device);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, ctx->device);
    if (prop.major < 10) {
        ctx->err = "device is not sm_100-class; this library is built for sm_100a only";
        CREATE_FAIL(MFLBM_ERR_CUDA);
    }
    int rc = [&]() -> int {
        int lo, hi;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&ctx->s_main, cudaStreamNonBlocking, lo));
        CU(cudaStreamCreateWithPriority(&ctx->s_halo, cudaStreamNonBlocking, hi));
        CU(cudaEventCreate(&ctx->ev_t0));
        CU(cudaEventCreate(&ctx->ev_t1));
        CU(cudaEventCreateWithFlags(&ctx->ev_slab, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_phi, cudaEventDisableTiming));
        CU(cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ctx->ev_out, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_out_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_ckpt, cudaEventDisableTiming));
        CU(cudaStreamCreateWithPriority(&ctx->s_chain, cudaStreamNonBlocking, hi));
        CU(cudaEventCreateWithFlags(&ctx->ev_near, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_chain, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_sum[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_sum[1], cudaEventDisableTiming));
        CU(cudaMallocHost((void **)&ctx->sum_host, 16 * sizeof(int)));
        memset(ctx->sum_host, 0, 16 * sizeof(int));
        ctx->ev_phi_valid = false;
        Dev &d = ctx->d;
        d.g.nx = cfg->nx; d.g.ny = cfg->ny; d.g.nz = cfg->nz;
        d.g.sx = (int)sx; d.g.sxy = (int)sxy; d.g.base = 16; d.g.ntot = (int)ntot;
        d.g.set_magic();
        d.multiphase = cfg->solver == MFLBM_SOLVER_MULTIPHASE;
        d.jper = cfg->jper != 0;
        d.kper = cfg->kper != 0;
        d.yw_lo = d.jper && (cfg->kper != 0 || (cfg->npz > 1 && cfg->idz != 0));
        d.yw_hi = d.jper && (cfg->kper != 0 || (cfg->npz > 1 && cfg->idz != cfg->npz - 1));
        d.mrt = cfg->mrt;
        d.la_nui1 = cfg->la_nui1; d.la_nui2 = cfg->la_nui2; d.gamma = cfg->gamma; d.beta = cfg->beta;
        d.force_Z = cfg->force_Z; d.phi_inlet = cfg->phi_inlet; d.sa_inject = cfg->sa_inject;
        d.relaxation = cfg->relaxation; d.uin_avg = cfg->uin_avg; d.rho_in = cfg->rho_in; d.rho_out = cfg->rho_out;
        d.s_e = cfg->s_e; d.s_e2 = cfg->s_e2; d.s_q = cfg->s_q; d.s_nu = cfg->s_nu; d.s_pi = cfg->s_pi; d.s_t = cfg->s_t;
        d.rk_weight2 = 1.0 / sqrt(2.0) / 36.0;
        // the 38 (19) population arrays are allocated at the first upload, once the wall array (and with it the
        // layout: dense grid or active-node list) is known
        if (dev_alloc(ctx, &d.walls, ntot)) return MFLBM_ERR_CUDA;
        if (dev_alloc(ctx, &d.w_in, (size_t)sxy + 32)) return MFLBM_ERR_CUDA;
        if (dev_alloc(ctx, &d.f_convec, (size_t)19 * sxy + 32)) return MFLBM_ERR_CUDA;
        if (d.multiphase) {
            if (dev_alloc(ctx, &d.phi, ntot)) return MFLBM_ERR_CUDA;
            if (dev_alloc(ctx, &d.cn_x, ntot) || dev_alloc(ctx, &d.cn_y, ntot) || dev_alloc(ctx, &d.cn_z, ntot) ||
                dev_alloc(ctx, &d.c_norm, ntot) || dev_alloc(ctx, &d.curv, ntot))
                return MFLBM_ERR_CUDA;
            if (dev_alloc(ctx, &d.g_convec, (size_t)19 * sxy + 32)) return MFLBM_ERR_CUDA;
            if (dev_alloc(ctx, &d.phi_convec, (size_t)sxy + 32)) return MFLBM_ERR_CUDA;
            d.num_solid = cfg->num_solid_boundary;
            d.num_fluid = cfg->num_fluid_boundary;
            if (d.num_solid > 0) {
                if (dev_alloc(ctx, &d.solid_cell, d.num_solid) || dev_alloc(ctx, &d.solid_mask, d.num_solid) ||
                    dev_alloc(ctx, &d.solid_law, d.num_solid))
                    return MFLBM_ERR_CUDA;
            }
            if (d.num_fluid > 0) {
                if (dev_alloc(ctx, &d.fluid_cell, d.num_fluid) || dev_alloc(ctx, &d.fluid_nw, (size_t)5 * d.num_fluid))
                    return MFLBM_ERR_CUDA;
            }
        }
        ctx->red_len = 16 * (cfg->nz > cfg->ny ? cfg->nz : cfg->ny) + 64;
        CU(cudaMalloc((void **)&ctx->red_dev, ctx->red_len * sizeof(double)));
        CU(cudaMallocHost((void **)&ctx->red_host, ctx->red_len * sizeof(double)));
        return 0;
    }();
    if (rc) CREATE_FAIL(rc);
    ctx->open_z = (cfg->kper == 0 && cfg->domain_wall_status_z_min == 0 && cfg->domain_wall_status_z_max == 0);
    // which phi ghost faces are rewritten every step by a kernel that also records their phi classes: the inlet / outlet
    // kernels (same conditions as launch_bc, kernels_bc.cu) or the halo exchange (same conditions as halo_exchange);
    // the ghost planes of the single-GPU periodic wrap stay "unknown"
    {
        const bool dyn = getenv("MFLBM_STATIC_BC_TILES") == nullptr;
        const bool halo = cfg->use_nccl && cfg->npz > 1;
        ctx->d.bc_lo_dyn = dyn && ((ctx->open_z && cfg->idz == 0 && (cfg->inlet_BC == 1 || cfg->inlet_BC == 2)) ||
                                   (halo && (cfg->kper == 1 || cfg->idz != 0)));
        ctx->d.bc_hi_dyn = dyn && ((ctx->open_z && cfg->idz == cfg->npz - 1 && (cfg->outlet_BC == 1 || cfg->outlet_BC == 2)) ||
                                   (halo && (cfg->kper == 1 || cfg->idz != cfg->npz - 1)));
    }
    ctx->peer_lo = (cfg->idz - 1 + cfg->npz) % cfg->npz;
    ctx->peer_hi = (cfg->idz + 1) % cfg->npz;
    if (cfg->use_nccl && cfg->npz > 1) {
        ctx->nccl = load_nccl(ctx->err);
        if (!ctx->nccl) CREATE_FAIL(MFLBM_ERR_NCCL);
        ncclUniqueId u;
        memcpy(&u, cfg->nccl_unique_id, 128);
        ncclComm_t comm;
        ncclResult_t r = ctx->nccl->CommInitRank(&comm, cfg->npz, u, cfg->idz);
        if (r != ncclSuccess) {
            ctx->err = std::string("ncclCommInitRank: ") + ctx->nccl->GetErrorString(r);
            CREATE_FAIL(MFLBM_ERR_NCCL);
        }
        ctx->comm = comm;
    }
    *out = ctx;
    return MFLBM_OK;
}

extern "C" void mflbm_destroy(mflbm_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->comm && ctx->nccl) ctx->nccl->CommDestroy(ctx->comm);
    for (void *p : ctx->allocs) cudaFree(p);
    if (ctx->stage) cudaFree(ctx->stage);
    for (int b = 0; b < 4; b++)
        if (ctx->halo_buf[b]) cudaFree(ctx->halo_buf[b]);
    if (ctx->red_dev) cudaFree(ctx->red_dev);
    if (ctx->red_host) cudaFreeHost(ctx->red_host);
    for (int b = 0; b < 2; b++) {
        if (ctx->res_host[b]) cudaFreeHost(ctx->res_host[b]);
        if (ctx->ev_win[b]) cudaEventDestroy(ctx->ev_win[b]);
        if (ctx->ev_res[b]) cudaEventDestroy(ctx->ev_res[b]);
        if (ctx->ev_step[b]) cudaEventDestroy(ctx->ev_step[b]);
    }
    if (ctx->s_main) cudaStreamDestroy(ctx->s_main);
    if (ctx->s_halo) cudaStreamDestroy(ctx->s_halo);
    if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
    for (int f = 0; f < 5; f++) {
        if (ctx->out_dev[f]) cudaFree(ctx->out_dev[f]);
        if (ctx->out_host[f]) cudaFreeHost(ctx->out_host[f]);
    }
    for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
    {
        double *ck[] = {ctx->ckpt_phi, ctx->ckpt_fc, ctx->ckpt_gc, ctx->ckpt_pc};
        for (double *p : ck)
            if (p) cudaFree(p);
        for (int q = 0; q < 19; q++) {
            if (ctx->ckpt_f[q]) cudaFree(ctx->ckpt_f[q]);
            if (ctx->ckpt_g[q]) cudaFree(ctx->ckpt_g[q]);
        }
    }
    if (ctx->s_chain) cudaStreamDestroy(ctx->s_chain);
    if (ctx->sum_host) cudaFreeHost(ctx->sum_host);
    cudaEvent_t evs[] = {ctx->ev_t0, ctx->ev_t1, ctx->ev_slab, ctx->ev_halo, ctx->ev_fork, ctx->ev_phi, ctx->ev_out, ctx->ev_out_done, ctx->ev_ckpt,
                         ctx->ev_near, ctx->ev_chain, ctx->ev_sum[0], ctx->ev_sum[1]};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    delete ctx;
}


// ---------------------------------------------------------------------------------------------------
// Population layout.  Built once from the wall array (host side, integer work only):
//   A nodes  = fluid nodes (walls==0) of 1..nx,1..ny,1..nz in raster order (k outer, i inner)
//   S nodes  = every other node of the 0..n+1 box that is a D3Q19 neighbour of an A node
//   nbr[q-1][n] = active index of x_n + e_q   (always exists by construction)
// ---------------------------------------------------------------------------------------------------
// Stable counting sort of list entries by tile: order[new position] = old index, start = CSR ranges [ntiles+1].
static void sort_by_tile(const std::vector<int> &tile, int ntiles, std::vector<int> &order, std::vector<int> &start) {
    start.assign((size_t)ntiles + 1, 0);
    for (int t : tile) start[(size_t)t + 1]++;
    for (int t = 0; t < ntiles; t++) start[(size_t)t + 1] += start[t];
    std::vector<int> pos(start.begin(), start.end() - 1);
    order.resize(tile.size());
    for (size_t n = 0; n < tile.size(); n++) order[(size_t)pos[tile[n]]++] = (int)n;
}

// Order of the flat sweeps (Dev::gcell_r): stable counting sort of a list of cells by brick, so that a list that comes in
// raster order keeps it inside every brick.  order[r] = position in the input of the r-th entry in flat order.
static void flat_order(const mflbm_ctx *ctx, const int *cells, size_t n, std::vector<int> &order, bool k7 = false) {
    const Grid &g = ctx->d.g;
    const int bx = k7 ? ctx->k7_bx : ctx->flat_bx, by = k7 ? ctx->k7_by : ctx->flat_by, bz = k7 ? ctx->k7_bz : ctx->flat_bz;
    const int nbx = bx > 0 ? (g.sx + bx - 1) / bx : 1, nby = (g.ny + 8 + by - 1) / by, nbz = (g.nz + 8 + bz - 1) / bz;
    std::vector<int> brick(n), start;
#pragma omp parallel for schedule(static)
    for (long long e = 0; e < (long long)n; e++) {
        unsigned ix, jy, kz;
        g.coords3(cells[e], ix, jy, kz);
        brick[(size_t)e] = (bx > 0 ? (int)ix / bx : 0) + nbx * ((int)jy / by + nby * ((int)kz / bz));
    }
    sort_by_tile(brick, nbx * nby * nbz, order, start);
}

// Host-only part (integer work, OpenMP): node numbering and the compressed adjacency of the odd step.
struct HostActive {
    int nA = 0;
    long long nAct = 0;
    int nlink[19] = {0};
    std::vector<int> kstartA, cellA, adjfull;
    std::vector<uint4> adj;
};

static int build_active_host(const Grid &g, const int8_t *walls /* (-1:n+2)^3, i fastest */, HostActive &H, std::string &err,
                             bool check, bool jper = false, bool yw_lo = false, bool yw_hi = false) {
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    const long long bx = nx + 2, by = ny + 2, bz = nz + 2;  // 0..n+1 box
    auto W = [&](int i, int j, int k) -> int8_t {
        return walls[(size_t)(i + 1) + (size_t)(nx + 4) * ((size_t)(j + 1) + (size_t)(ny + 4) * (size_t)(k + 1))];
    };
    std::vector<int> idx((size_t)(bx * by * bz), -1);
    auto B = [&](int i, int j, int k) -> size_t { return (size_t)i + (size_t)bx * ((size_t)j + (size_t)by * (size_t)k); };
    // pass 1: A nodes per plane
    std::vector<int> cntA(nz + 2, 0), cntS(nz + 2, 0);
#pragma omp parallel for schedule(static)
    for (int k = 1; k <= nz; k++) {
        int c = 0;
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) c += (W(i, j, k) == 0);
        cntA[k] = c;
    }
    H.kstartA.assign(nz + 3, 0);
    for (int k = 1; k <= nz + 1; k++) H.kstartA[k] = H.kstartA[k - 1] + cntA[k - 1];
    H.kstartA[nz + 2] = H.kstartA[nz + 1];
    const int nA = H.kstartA[nz + 1];
#pragma omp parallel for schedule(static)
    for (int k = 1; k <= nz; k++) {
        int n = H.kstartA[k];
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++)
                if (W(i, j, k) == 0) idx[B(i, j, k)] = n++;
    }
    auto isA = [&](int i, int j, int k) -> bool { return i >= 1 && i <= nx && j >= 1 && j <= ny && k >= 1 && k <= nz && W(i, j, k) == 0; };
    // pass 2: S nodes (mark with -2), count per plane
#pragma omp parallel for schedule(static)
    for (int k = 0; k <= nz + 1; k++) {
        int c = 0;
        for (int j = 0; j <= ny + 1; j++)
            for (int i = 0; i <= nx + 1; i++) {
                if (isA(i, j, k)) continue;
                if (k > 2 && k < nz - 1) continue;  // outside the cell-addressable zone: served by compact link slots
                // Ghost-plane cells that are fluid by their own wall flag (= the flag of the periodic image / of the
                // neighbour slab's boundary plane) always get storage: the reference's exchange copies their slots into the
                // boundary plane unconditionally (MP/Mpi.F90:374-396, :497-519), and right after an upload they may hold
                // values nothing else reproduces (initial_fluid_distribution_option 6 draws phi in the ghost layers
                // independently of their images, MP/Init_multiphase.F90:276-311).
                bool s = (k == 0 || k == nz + 1) && i >= 1 && i <= nx && j >= 1 && j <= ny && W(i, j, k) == 0;
                for (int q = 1; q < 19 && !s; q++) s = isA(i + EX(q), j + EY(q), k + EZ(q));
                // y-periodic: also the cells a fluid node reaches through the seam (same wrap rule as nbr_of below), so that
                // the inlet / outlet kernels can address the image of a ghost-row cell by its cell index
                if (jper && ((k >= 1 && k <= nz) || (k < 1 && yw_lo) || (k > nz && yw_hi)))
                    for (int q = 1; q < 19 && !s; q++) {
                        const int j2 = j + EY(q);
                        if (j2 < 1 || j2 > ny) s = isA(i + EX(q), j2 < 1 ? j2 + ny : j2 - ny, k + EZ(q));
                    }
                if (s) {
                    idx[B(i, j, k)] = -2;
                    c++;
                }
            }
        cntS[k] = c;
    }
    std::vector<long long> startS(nz + 3, 0);
    startS[0] = nA;
    for (int k = 1; k <= nz + 2; k++) startS[k] = startS[k - 1] + cntS[k - 1];
    const long long nAct = startS[nz + 2];
    if (nAct >= (1LL << 31) - 64) { err = "too many active nodes"; return -1; }
#pragma omp parallel for schedule(static)
    for (int k = 0; k <= nz + 1; k++) {
        int n = (int)startS[k];
        for (int j = 0; j <= ny + 1; j++)
            for (int i = 0; i <= nx + 1; i++)
                if (idx[B(i, j, k)] == -2) idx[B(i, j, k)] = n++;
    }
    // cell list
    std::vector<int> &cellA = H.cellA;
    cellA.assign((size_t)nAct, 0);
#pragma omp parallel for schedule(static)
    for (int k = 0; k <= nz + 1; k++)
        for (int j = 0; j <= ny + 1; j++)
            for (int i = 0; i <= nx + 1; i++) {
                const int n = idx[B(i, j, k)];
                if (n >= 0) cellA[n] = g.cell(i, j, k);
            }
    // Compressed adjacency of the odd step (mflbm_internal.cuh "Adjacency").  For warp w = 32 consecutive A nodes and
    // direction d the neighbour x+e_d of lane l is either
    //   a compact LINK slot (x+e_d is neither an A node nor a zone S node: bounce-back storage private to this link),
    //       index nAct + lbase + rank of the lane among the warp's link lanes, or
    //   a node index that continues the previous non-link lane's index by +1, or starts a new run (a JUMP, value stored;
    //       the first four jumps of a record inline, further ones in the overflow array jval).
    const int nW = (nA + 31) / 32;
    auto nbr_of = [&](int n, int q) -> int {  // active index of x_n + e_q, or -1 (-> link)
        const unsigned r = (unsigned)(cellA[n] - (g.base - 4));
        const unsigned kz = r / (unsigned)g.sxy, r2 = r - kz * (unsigned)g.sxy;
        const unsigned jy = r2 / (unsigned)g.sx, ix = r2 - jy * (unsigned)g.sx;
        int j2 = (int)jy - 3 + EY(q);
        const int k2 = (int)kz - 3 + EZ(q);
        // y-periodic lattice (one process in y: the reference exchanges with itself, MP/Mpi.F90:147-207, :398-456): the
        // neighbour across the seam IS the periodic image, so the y faces -- and, with z periodic too, the x edges, through the
        // z ghost-plane cell of the image column, which k_wrap_z / the halo exchange with the neighbour slab serves -- need
        // no copies at all.  At an open end of the lattice the reference exchanges rows k = 1..nz only: the ghost-plane cells
        // behind the seam stay what the inlet / outlet routines make of them, and so they do here.
        if (jper && ((k2 >= 1 && k2 <= nz) || (k2 < 1 && yw_lo) || (k2 > nz && yw_hi))) j2 = j2 < 1 ? j2 + ny : (j2 > ny ? j2 - ny : j2);
        return idx[B((int)ix - 3 + EX(q), j2, k2)];
    };
    const bool stats = getenv("MFLBM_ADJ_STATS") != nullptr;  // developer: histogram of index runs per (warp, direction)
    long long hist[16] = {0};
    std::vector<int> lcnt((size_t)nW * 18 + 18, 0);
    std::vector<long long> irr((size_t)nW + 1, 0);  // irregular rows per warp -> exclusive prefix
    std::vector<unsigned> irrmask((size_t)nW + 1, 0);
#pragma omp parallel for schedule(static)
    for (int w = 0; w < nW; w++) {
        unsigned im = 0;
        const int l1 = std::min(32, nA - 32 * w);
        for (int q = 1; q < 19; q++) {
            int links = 0, jumps = 0, prev = -2;
            for (int l = 0; l < l1; l++) {
                const int v = nbr_of(32 * w + l, q);
                if (v < 0) { links++; continue; }
                if (v != prev + 1) jumps++;
                prev = v;
            }
            lcnt[(size_t)w * 18 + (q - 1)] = links;
            if (stats) {
#pragma omp atomic
                hist[jumps > 15 ? 15 : jumps]++;
            }
            if (jumps > 5) im |= 1u << (q - 1);
        }
        irrmask[w] = im;
        irr[w] = __builtin_popcount(im);
    }
    if (stats) {
        fprintf(stderr, "adjacency runs per (warp, direction):");
        for (int j = 0; j < 16; j++) fprintf(stderr, " %d:%.3f%%", j, 100.0 * hist[j] / (18.0 * std::max(nW, 1)));
        fprintf(stderr, "\n");
    }
    // exclusive prefix sums over the warps
    long long lsum[18] = {0};
    for (int w = 0; w < nW; w++)
        for (int q = 0; q < 18; q++) {
            const int c = lcnt[(size_t)w * 18 + q];
            lcnt[(size_t)w * 18 + q] = (int)lsum[q];
            lsum[q] += c;
        }
    long long osum = 0;
    for (int w = 0; w < nW; w++) {
        const long long c = irr[w];
        irr[w] = osum;
        osum += c;
    }
    H.nlink[0] = 0;
    for (int q = 1; q < 19; q++) {
        if (nAct + lsum[q - 1] >= (1LL << 31) - 64) { err = "too many active nodes"; return -1; }
        H.nlink[q] = (int)lsum[q - 1];
    }
    if (osum * 32 >= (1LL << 31)) { err = "adjacency: too many irregular rows"; return -1; }
    std::vector<uint4> &adj = H.adj;
    std::vector<int> &full = H.adjfull;
    adj.assign((size_t)std::max(nW, 1) * MFLBM_ADJ_REC, uint4{0, 0, 0, 0});
    full.assign((size_t)std::max<long long>(osum, 1) * 32, 0);
#pragma omp parallel for schedule(static)
    for (int w = 0; w < nW; w++) {
        long long row = irr[w];
        const int l1 = std::min(32, nA - 32 * w);
        uint4 *rec = &adj[(size_t)w * MFLBM_ADJ_REC];
        rec[0] = uint4{irrmask[w], (unsigned)irr[w], 0u, 0u};
        for (int q = 1; q < 19; q++) {
            const int lb = lcnt[(size_t)w * 18 + (q - 1)];
            unsigned smask = 0, jmask = 0;
            int inl[5] = {0, 0, 0, 0, 0};
            int jumps = 0, prev = -2, links = 0;
            const bool irregular = (irrmask[w] >> (q - 1)) & 1u;
            for (int l = 0; l < l1; l++) {
                const int v = nbr_of(32 * w + l, q);
                if (v < 0) {
                    if (irregular) full[(size_t)row * 32 + l] = (int)nAct + lb + links;
                    smask |= 1u << l;
                    links++;
                    continue;
                }
                if (irregular) full[(size_t)row * 32 + l] = v;
                if (v != prev + 1) {
                    jmask |= 1u << l;
                    if (jumps < 5) inl[jumps] = v - (l - links);  // B_r = first index - rank of the first lane
                    jumps++;
                }
                prev = v;
            }
            if (irregular) row++;
            rec[1 + 2 * (q - 1)] = uint4{smask, jmask, (unsigned)lb, (unsigned)inl[0]};
            rec[2 + 2 * (q - 1)] = uint4{(unsigned)inl[1], (unsigned)inl[2], (unsigned)inl[3], (unsigned)inl[4]};
        }
    }
    H.nA = nA;
    H.nAct = nAct;
    if (check) {  // decode every (node, direction) with the device function and compare with the direct lookup
        long long bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
        for (int n = 0; n < nA; n++) {
            const int w = n >> 5, l = n & 31;
            for (int q = 1; q < 19; q++) {
                const uint4 *rec = &adj[(size_t)w * MFLBM_ADJ_REC];
                const int got = adj_lookup(rec, full.data(), q, l, (int)nAct);
                // the collision kernel's fast path must agree wherever it is taken (warps without irregular directions)
                if (rec[0].x == 0) bad += (adj_index_fast(reinterpret_cast<const int *>(rec), q, l, (int)nAct) != got);
                const int v = nbr_of(n, q);
                if (v >= 0) bad += (got != v);
                else bad += (got < nAct || got >= nAct + H.nlink[q]);  // a link slot of direction q
            }
        }
        if (bad) { err = "compressed adjacency self-check failed"; return -1; }
    }
    return 0;
}

static int build_active_set(mflbm_ctx *ctx, const int8_t *walls /* (-1:n+2)^3, i fastest */) {
    Dev &d = ctx->d;
    const Grid &g = d.g;
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    auto W = [&](int i, int j, int k) -> int8_t {
        return walls[(size_t)(i + 1) + (size_t)(nx + 4) * ((size_t)(j + 1) + (size_t)(ny + 4) * (size_t)(k + 1))];
    };
    HostActive H;
    {
        std::string err;
        const char *e = getenv("MFLBM_CHECK_ADJ");
        const bool check = e ? atoi(e) != 0 : ((long long)nx * ny * nz <= 8000000LL);
        if (build_active_host(g, walls, H, err, check, d.jper != 0, d.yw_lo != 0, d.yw_hi != 0)) return fail(ctx, MFLBM_ERR_ARG, err);
    }
    ctx->kstartA = H.kstartA;
    for (int q = 0; q < 19; q++) d.nlink[q] = H.nlink[q];
    const int nA = H.nA;
    const long long nAct = H.nAct;
    std::vector<int> &cellA = H.cellA;
    std::vector<uint4> &adj = H.adj;
    std::vector<int> &full = H.adjfull;
    ctx->adj_bytes = (long long)(adj.size() * sizeof(uint4) + full.size() * sizeof(int));
    // G list: non-solid cells of the (-1:n+2)^3 box, raster order (where K4 evaluates the colour gradient)
    std::vector<int> gcnt(nz + 5, 0);
#pragma omp parallel for schedule(static)
    for (int k = -1; k <= nz + 2; k++) {
        int c = 0;
        for (int j = -1; j <= ny + 2; j++)
            for (int i = -1; i <= nx + 2; i++) c += (W(i, j, k) != 1);
        gcnt[k + 2] = c;
    }
    std::vector<long long> gstart(nz + 6, 0);
    for (int k = 0; k <= nz + 3; k++) gstart[k + 1] = gstart[k] + gcnt[k + 1];
    const long long nG = gstart[nz + 4];
    std::vector<int> gcell((size_t)(nG > 0 ? nG : 1));
#pragma omp parallel for schedule(static)
    for (int k = -1; k <= nz + 2; k++) {
        long long n = gstart[k + 1];
        for (int j = -1; j <= ny + 2; j++)
            for (int i = -1; i <= nx + 2; i++)
                if (W(i, j, k) != 1) gcell[n++] = g.cell(i, j, k);
    }
    d.nG = (int)nG;
    if (d.use_tiles && nG > 0) {  // flat sweeps (most tiles active): brick order (plain raster order by default)
        if (dev_alloc(ctx, &d.gcell_r, gcell.size(), false)) return MFLBM_ERR_CUDA;
        if (ctx->flat_bx == 0 && ctx->flat_by == 1 && ctx->flat_bz == 1) {
            CU(cudaMemcpy(d.gcell_r, gcell.data(), gcell.size() * sizeof(int), cudaMemcpyHostToDevice));
        } else {
            std::vector<int> order, sorted((size_t)nG);
            flat_order(ctx, gcell.data(), (size_t)nG, order);
#pragma omp parallel for schedule(static)
            for (long long n = 0; n < nG; n++) sorted[(size_t)n] = gcell[(size_t)order[(size_t)n]];
            CU(cudaMemcpy(d.gcell_r, sorted.data(), sorted.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
    }
    d.aorder = d.acell = nullptr;
    if (d.use_tiles && nA > 0 && !(ctx->k7_bx == 0 && ctx->k7_by == 1 && ctx->k7_bz == 1)) {  // K7 + packing in brick order
        std::vector<int> order, cells((size_t)nA);
        flat_order(ctx, cellA.data(), (size_t)nA, order, true);
#pragma omp parallel for schedule(static)
        for (int n = 0; n < nA; n++) cells[(size_t)n] = cellA[(size_t)order[(size_t)n]];
        if (dev_alloc(ctx, &d.aorder, (size_t)nA, false) || dev_alloc(ctx, &d.acell, (size_t)nA, false)) return MFLBM_ERR_CUDA;
        CU(cudaMemcpy(d.aorder, order.data(), (size_t)nA * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d.acell, cells.data(), (size_t)nA * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (d.use_tiles && nG > 0) {  // tile-driven gradient chain: gcell grouped by tile (raster order inside a tile)
        std::vector<int> tile((size_t)nG), order, start;
#pragma omp parallel for schedule(static)
        for (long long n = 0; n < nG; n++) tile[(size_t)n] = g.tile_of(gcell[(size_t)n], d.ntx, d.nty);
        sort_by_tile(tile, d.ntiles, order, start);
        std::vector<int> sorted((size_t)nG);
#pragma omp parallel for schedule(static)
        for (long long n = 0; n < nG; n++) sorted[(size_t)n] = gcell[(size_t)order[(size_t)n]];
        gcell.swap(sorted);
        CU(cudaMemcpy(d.tg_start, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    if (dev_alloc(ctx, &d.gcell, gcell.size(), false)) return MFLBM_ERR_CUDA;
    CU(cudaMemcpy(d.gcell, gcell.data(), gcell.size() * sizeof(int), cudaMemcpyHostToDevice));
    d.nA = nA;
    d.nAct = (int)nAct;
    {
        // L2 software prefetch of the odd step, in nodes ahead (about one wave of resident warps).  Measured on B200 with
        // the compressed adjacency: singlephase C2 +5 % MLUPS at 65536; multiphase C3 -3 % (the extra issue slots cost
        // more than the L2 hits return), so it is off there.  MFLBM_PF_DIST overrides.
        const char *e = getenv("MFLBM_PF_DIST");
        d.pf_dist = e ? atoi(e) : (d.multiphase ? 0 : 65536);
        const char *m = getenv("MFLBM_PF_MODE");
        d.pf_mode = m ? atoi(m) : 1;
    }
    if (dev_alloc(ctx, &d.cellA, (size_t)nAct + 32, false) || dev_alloc(ctx, &d.adj, adj.size(), false) ||
        dev_alloc(ctx, &d.adjfull, full.size(), false) || dev_alloc(ctx, &d.smap, (size_t)g.ntot, false))
        return MFLBM_ERR_CUDA;
    CU(cudaMemcpy(d.cellA, cellA.data(), (size_t)nAct * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d.adj, adj.data(), adj.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d.adjfull, full.data(), full.size() * sizeof(int), cudaMemcpyHostToDevice));
    launch_fill_smap(ctx, ctx->s_main);
    CU(cudaStreamSynchronize(ctx->s_main));
    return 0;
}

// choose the population layout from the wall array and allocate the populations
static int setup_populations(mflbm_ctx *ctx, const int8_t *walls) {
    Dev &d = ctx->d;
    const Grid &g = d.g;
    const mflbm_config &cfg = ctx->cfg;
    int variant = cfg.kernel_variant;  // 0 auto, 1 dense, 2 sparse
    if (cfg.porous_plate_cmd != 0) variant = 1;  // the porous plate copies from arbitrary (non-active) nodes
    if (cfg.jper != 0) variant = 2;              // the y wrap of the populations is part of the adjacency (build_active_host)
    if (variant == 0) {
        long long fluid = 0;
#pragma omp parallel for reduction(+ : fluid) schedule(static)
        for (int k = 1; k <= g.nz; k++)
            for (int j = 1; j <= g.ny; j++)
                for (int i = 1; i <= g.nx; i++)
                    fluid += walls[(size_t)(i + 1) + (size_t)(g.nx + 4) * ((size_t)(j + 1) + (size_t)(g.ny + 4) * (size_t)(k + 1))] == 0;
        const double porosity = (double)fluid / ((double)g.nx * g.ny * g.nz);
        variant = porosity > 0.8 ? 1 : 2;
    }
    d.sparse = variant == 2;
    d.full_curv = variant == 1;
    d.use_tiles = (d.sparse && d.multiphase && !getenv("MFLBM_NO_TILES")) ? 1 : 0;
    d.k4_smem = getenv("MFLBM_K4_SMEM") ? 1 : 0;  // measured slower than the list gathers (kernels_gradient.cu), off by default
    if (d.use_tiles) {
        d.ntx = g.sx / 8;
        d.nty = (g.ny + 8 + 3) / 4;
        d.ntz = (g.nz + 8 + 3) / 4;
        d.ntiles = d.ntx * d.nty * d.ntz;
        d.tile_cur = 0;
        const size_t nt = (size_t)d.ntiles + 4;
        if (dev_alloc(ctx, &d.tcls[0], nt) || dev_alloc(ctx, &d.tcls[1], nt) || dev_alloc(ctx, &d.tstat, nt) ||
            dev_alloc(ctx, &d.tU[0], nt) || dev_alloc(ctx, &d.tU[1], nt) || dev_alloc(ctx, &d.tquiet, nt) ||
            dev_alloc(ctx, &d.tg_start, nt) || dev_alloc(ctx, &d.ts_start, nt) || dev_alloc(ctx, &d.tf_start, nt) ||
            dev_alloc(ctx, &d.tact, nt) || dev_alloc(ctx, &d.tk3, nt) || dev_alloc(ctx, &d.tk3stamp, nt) ||
            dev_alloc(ctx, &d.tcount, 12))
            return MFLBM_ERR_CUDA;
    }
    if (!d.tcount && dev_alloc(ctx, &d.tcount, 12)) return MFLBM_ERR_CUDA;  // [7] is also the sink of Stencil19::load
    size_t n = g.ntot;
    if (d.sparse) {
        if (build_active_set(ctx, walls)) return MFLBM_ERR_CUDA;
        n = (size_t)d.nAct + 64;
        if (d.use_tiles && dev_alloc(ctx, &d.wstamp, (size_t)(d.nA + 31) / 32 + 1)) return MFLBM_ERR_CUDA;
        d.wq_all = 1;
    }
    for (int q = 0; q < 19; q++) {
        // sparse: array q also holds the compact link slots of direction opc(q) behind the node entries
        const size_t nq = d.sparse ? n + (size_t)d.nlink[OPC(q)] : n;
        // (staggering the 2 MiB-aligned array bases against L2 set conflicts was measured: no effect, r01_v4)
        if (dev_alloc(ctx, &d.f[q], nq)) return MFLBM_ERR_CUDA;
        if (d.multiphase && dev_alloc(ctx, &d.gg[q], nq)) return MFLBM_ERR_CUDA;
    }
    if (d.sparse && d.multiphase)
        for (int m = 0; m < 4; m++)
            if (dev_alloc(ctx, &d.G[m], (size_t)d.nA + 64)) return MFLBM_ERR_CUDA;
    if (!(d.sparse && d.multiphase)) ctx->march_on = false;
    d.lazy_ok = (ctx->march_on && ctx->march_hybrid) ? 0 : 1;
    if (ctx->march_on) {
        d.mcols_x = (g.nx + MFLBM_MARCH_TX - 1) / MFLBM_MARCH_TX;
        d.mcols_y = (g.ny + MFLBM_MARCH_TY - 1) / MFLBM_MARCH_TY;
        d.march_lz = getenv("MFLBM_MARCH_LZR") ? std::max(1, atoi(getenv("MFLBM_MARCH_LZR"))) : 16;
        d.mchunks = (g.nz + d.march_lz - 1) / d.march_lz;
        if (dev_alloc(ctx, &d.mcode, (size_t)g.ntot, false)) return MFLBM_ERR_CUDA;
        if (d.use_tiles) {
            const size_t ni = (size_t)d.mcols_x * d.mcols_y * d.mchunks;
            if (dev_alloc(ctx, &d.mflag, ni) || dev_alloc(ctx, &d.mlist, ni)) return MFLBM_ERR_CUDA;
        }
        ctx->march_ready = false;
    }
    ctx->pdf_alloc = true;
    return 0;
}

static int ensure_stage(mflbm_ctx *ctx, size_t bytes) {
    if (ctx->stage_bytes >= bytes) return 0;
    if (ctx->stage) cudaFree(ctx->stage);
    ctx->stage = nullptr;
    ctx->stage_bytes = 0;
    CU(cudaMalloc((void **)&ctx->stage, bytes));
    ctx->stage_bytes = bytes;
    return 0;
}

static int ensure_macro(mflbm_ctx *ctx) {
    if (ctx->macro_alloc) return 0;
    Dev &d = ctx->d;
    if (dev_alloc(ctx, &d.u, d.g.ntot) || dev_alloc(ctx, &d.v, d.g.ntot) || dev_alloc(ctx, &d.w, d.g.ntot) ||
        dev_alloc(ctx, &d.rho, d.g.ntot))
        return MFLBM_ERR_CUDA;
    ctx->macro_alloc = true;
    return 0;
}

static int ensure_phi_old(mflbm_ctx *ctx) {
    Dev &d = ctx->d;
    if (d.phi_old) return 0;
    if (dev_alloc(ctx, &d.phi_old, d.g.ntot, false)) return MFLBM_ERR_CUDA;
    // never uploaded: start from the current phase field, like the reference seeds phi_old = phi when
    // steady_state_option == 2 (MP/Init_multiphase.F90:341-347), so that the first d_phi_max is a real change
    CU(cudaMemcpyAsync(d.phi_old, d.phi, (size_t)d.g.ntot * sizeof(double), cudaMemcpyDeviceToDevice, ctx->s_main));
    return 0;
}

// one field: host (ghost width o, nplanes planes starting at grid plane kbase) <-> device grid, on stream st
static int xfer(mflbm_ctx *ctx, double *dev, double *host, int o, int nplanes, int kbase, bool up, cudaStream_t st = nullptr) {
    if (!host || !dev) return 0;
    if (!st) st = ctx->s_main;
    const Grid &g = ctx->d.g;
    const size_t n = (size_t)(g.nx + 2 * o) * (g.ny + 2 * o) * nplanes;
    if (ensure_stage(ctx, n * sizeof(double))) return MFLBM_ERR_CUDA;
    if (up) {
        CU(cudaMemcpyAsync(ctx->stage, host, n * sizeof(double), cudaMemcpyHostToDevice, st));
        launch_repack(ctx, st, dev, ctx->stage, o, nplanes, kbase, true);
        CU(cudaStreamSynchronize(st));
    } else {
        launch_repack(ctx, st, dev, ctx->stage, o, nplanes, kbase, false);
        CU(cudaMemcpyAsync(host, ctx->stage, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    return 0;
}

// one population array: caller's (0:nx+1,0:ny+1,0:nz+1) array <-> device (dense grid or active-node list).
// Sparse download overwrites only the active entries of the caller's array: all other entries are never
// touched by the reference either (they keep the values the caller's array already has).
static int xfer_pdf(mflbm_ctx *ctx, double *dev, double *host, bool up, int q, cudaStream_t st = nullptr) {
    if (!host) return 0;
    if (!ctx->pdf_alloc) return fail(ctx, MFLBM_ERR_STATE, "populations uploaded before the wall array");
    if (!st) st = ctx->s_main;
    const Grid &g = ctx->d.g;
    if (!ctx->d.sparse) return xfer(ctx, dev, host, 1, g.nz + 2, 0, up, st);
    const size_t n = (size_t)(g.nx + 2) * (g.ny + 2) * (g.nz + 2);
    if (ensure_stage(ctx, n * sizeof(double))) return MFLBM_ERR_CUDA;
    CU(cudaMemcpyAsync(ctx->stage, host, n * sizeof(double), cudaMemcpyHostToDevice, st));
    launch_repack_sparse(ctx, st, dev, ctx->stage, up, q);
    if (!up) CU(cudaMemcpyAsync(host, ctx->stage, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

// Dev::gk5_r from the flat-order fluid list: returns the number of cells listed more than once, -1 on a CUDA error
static int build_gk5(mflbm_ctx *ctx, bool flat) {
    Dev &d = ctx->d;
    int *map = nullptr, *dup = nullptr;
    if (cudaMalloc((void **)&map, (size_t)d.g.ntot * sizeof(int)) != cudaSuccess) return -1;
    if (cudaMalloc((void **)&dup, sizeof(int)) != cudaSuccess) { cudaFree(map); return -1; }
    cudaMemsetAsync(map, 0xff, (size_t)d.g.ntot * sizeof(int), ctx->s_main);
    cudaMemsetAsync(dup, 0, sizeof(int), ctx->s_main);
    launch_build_gk5(ctx, ctx->s_main, map, dup, flat);
    int h = -1;
    const bool ok = cudaMemcpyAsync(&h, dup, sizeof(int), cudaMemcpyDeviceToHost, ctx->s_main) == cudaSuccess &&
                    cudaStreamSynchronize(ctx->s_main) == cudaSuccess;
    cudaFree(map);
    cudaFree(dup);
    return ok ? h : -1;
}

extern "C" int mflbm_upload(mflbm_ctx *ctx, const mflbm_arrays *h) {
    if (!ctx || !h) return fail(ctx, MFLBM_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    Dev &d = ctx->d;
    const Grid &g = d.g;
    const int nz = g.nz;
    if (h->walls) {
        if (ctx->pdf_alloc) return fail(ctx, MFLBM_ERR_STATE, "the wall array can only be uploaded once per context");
        const size_t n = (size_t)(g.nx + 4) * (g.ny + 4) * (nz + 4);
        if (ensure_stage(ctx, n)) return MFLBM_ERR_CUDA;
        CU(cudaMemcpyAsync(ctx->stage, h->walls, n, cudaMemcpyHostToDevice, ctx->s_main));
        // cells outside the (-1:n+2) box keep walls=0 like an untouched allocation would; they are never read
        launch_repack_i8(ctx, ctx->s_main, d.walls, (int8_t *)ctx->stage, 2, true);
        CU(cudaStreamSynchronize(ctx->s_main));
        d.gk5 = nullptr;
        d.gk5_r = nullptr;  // belongs to the cell list of the previous wall array; rebuilt with the next fluid boundary list
        if (setup_populations(ctx, h->walls)) return MFLBM_ERR_CUDA;
    }
    for (int q = 0; q < 19; q++) {
        if (xfer_pdf(ctx, d.f[q], h->f[q], true, q)) return MFLBM_ERR_CUDA;
        if (d.multiphase && xfer_pdf(ctx, d.gg[q], h->g[q], true, q)) return MFLBM_ERR_CUDA;
    }
    if (h->walls || h->phi || h->solid_boundary_nodes) ctx->tiles_static_ready = false;  // quiet-tile state restarts
    if (h->walls || h->solid_boundary_nodes || h->fluid_boundary_nodes) ctx->march_ready = false;  // cell codes of the march kernel
    if (d.multiphase) {
        if (xfer(ctx, d.phi, h->phi, 4, nz + 8, -3, true)) return MFLBM_ERR_CUDA;
        if (h->phi_old) {
            if (ensure_phi_old(ctx)) return MFLBM_ERR_CUDA;
            if (xfer(ctx, d.phi_old, h->phi_old, 4, nz + 8, -3, true)) return MFLBM_ERR_CUDA;
        }
        if (xfer(ctx, d.g_convec, h->g_convec_bc, 1, 19, -3, true)) return MFLBM_ERR_CUDA;
        if (xfer(ctx, d.phi_convec, h->phi_convec_bc, 1, 1, -3, true)) return MFLBM_ERR_CUDA;
    }
    if (xfer(ctx, d.f_convec, h->f_convec_bc, 1, 19, -3, true)) return MFLBM_ERR_CUDA;
    if (xfer(ctx, d.w_in, h->w_in, 1, 1, -3, true)) return MFLBM_ERR_CUDA;
    if (d.multiphase && (h->solid_boundary_nodes || h->fluid_boundary_nodes) && !ctx->pdf_alloc)
        return fail(ctx, MFLBM_ERR_STATE, "boundary-node lists uploaded before the wall array");
    if (d.multiphase && h->solid_boundary_nodes && d.num_solid > 0) {
        std::vector<int> cell(d.num_solid);
        std::vector<unsigned> mask(d.num_solid);
        std::vector<double> law(d.num_solid);
        for (int n = 0; n < d.num_solid; n++) {
            const mflbm_solid_node &s = h->solid_boundary_nodes[n];
            if (s.ix < -2 || s.ix > g.nx + 3 || s.iy < -2 || s.iy > g.ny + 3 || s.iz < -2 || s.iz > nz + 3)
                return fail(ctx, MFLBM_ERR_ARG, "solid boundary node outside the 3-ghost-layer box");
            if (s.i_fluid_num < 0 || s.i_fluid_num > 18) return fail(ctx, MFLBM_ERR_ARG, "solid node: bad i_fluid_num");
            cell[n] = g.cell(s.ix, s.iy, s.iz);
            unsigned m = 0;
            int prev = 0;
            for (int t = 0; t < s.i_fluid_num; t++) {
                const int e = s.neighbor_list[t];
                if (e <= prev || e > 18) return fail(ctx, MFLBM_ERR_ARG, "solid node: neighbor_list must be increasing in 1..18");
                m |= 1u << e;
                prev = e;
            }
            if (s.ix >= 0 && s.ix <= g.nx + 1 && s.iy >= 0 && s.iy <= g.ny + 1 && s.iz >= 0 && s.iz <= nz + 1) m |= 0x80000000u;
            mask[n] = m;
            law[n] = s.la_weight;
        }
        if (d.tcls[0]) {  // the flat sweeps read the list in flat order (Dev::gcell_r)
            std::vector<int> order;
            flat_order(ctx, cell.data(), cell.size(), order);
            std::vector<int> cell2(d.num_solid);
            std::vector<unsigned> mask2(d.num_solid);
            std::vector<double> law2(d.num_solid);
            for (int n = 0; n < d.num_solid; n++) {
                cell2[n] = cell[order[n]]; mask2[n] = mask[order[n]]; law2[n] = law[order[n]];
            }
            if (!d.solid_cell_r && (dev_alloc(ctx, &d.solid_cell_r, d.num_solid, false) || dev_alloc(ctx, &d.solid_mask_r, d.num_solid, false) ||
                                    dev_alloc(ctx, &d.solid_law_r, d.num_solid, false)))
                return MFLBM_ERR_CUDA;
            CU(cudaMemcpy(d.solid_cell_r, cell2.data(), cell2.size() * sizeof(int), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(d.solid_mask_r, mask2.data(), mask2.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(d.solid_law_r, law2.data(), law2.size() * sizeof(double), cudaMemcpyHostToDevice));
        } else {
            d.solid_cell_r = d.solid_cell; d.solid_mask_r = d.solid_mask; d.solid_law_r = d.solid_law;
        }
        if (d.tcls[0]) {  // group by tile for the tile-driven chain (entries are independent of each other)
            std::vector<int> tile(d.num_solid), order, start;
            for (int n = 0; n < d.num_solid; n++) tile[n] = g.tile_of(cell[n], d.ntx, d.nty);
            sort_by_tile(tile, d.ntiles, order, start);
            std::vector<int> cell2(d.num_solid);
            std::vector<unsigned> mask2(d.num_solid);
            std::vector<double> law2(d.num_solid);
            for (int n = 0; n < d.num_solid; n++) {
                cell2[n] = cell[order[n]]; mask2[n] = mask[order[n]]; law2[n] = law[order[n]];
            }
            cell.swap(cell2); mask.swap(mask2); law.swap(law2);
            CU(cudaMemcpy(d.ts_start, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
        CU(cudaMemcpy(d.solid_cell, cell.data(), cell.size() * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d.solid_mask, mask.data(), mask.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d.solid_law, law.data(), law.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (d.multiphase && h->fluid_boundary_nodes && d.num_fluid > 0) {
        std::vector<int> cell(d.num_fluid);
        std::vector<double> nw((size_t)5 * d.num_fluid);
        for (int n = 0; n < d.num_fluid; n++) {
            const mflbm_fluid_node &s = h->fluid_boundary_nodes[n];
            if (s.ix < -1 || s.ix > g.nx + 2 || s.iy < -1 || s.iy > g.ny + 2 || s.iz < -1 || s.iz > nz + 2)
                return fail(ctx, MFLBM_ERR_ARG, "fluid boundary node outside the 2-ghost-layer box");
            cell[n] = g.cell(s.ix, s.iy, s.iz);
            const size_t nf = (size_t)d.num_fluid;
            nw[0 * nf + n] = s.nwx; nw[1 * nf + n] = s.nwy; nw[2 * nf + n] = s.nwz;
            nw[3 * nf + n] = cos(s.theta);  // dcos/dsin of MP/Phase_gradient.F90:238-242, evaluated once by the host libm
            nw[4 * nf + n] = sin(s.theta);
        }
        if (d.tcls[0]) {
            std::vector<int> order;
            flat_order(ctx, cell.data(), cell.size(), order);
            std::vector<int> cell2(d.num_fluid);
            std::vector<double> nw2((size_t)5 * d.num_fluid);
            for (int n = 0; n < d.num_fluid; n++) {
                cell2[n] = cell[order[n]];
                for (int m = 0; m < 5; m++) nw2[(size_t)m * d.num_fluid + n] = nw[(size_t)m * d.num_fluid + order[n]];
            }
            if (!d.fluid_cell_r && (dev_alloc(ctx, &d.fluid_cell_r, d.num_fluid, false) || dev_alloc(ctx, &d.fluid_nw_r, (size_t)5 * d.num_fluid, false)))
                return MFLBM_ERR_CUDA;
            CU(cudaMemcpy(d.fluid_cell_r, cell2.data(), cell2.size() * sizeof(int), cudaMemcpyHostToDevice));
            CU(cudaMemcpy(d.fluid_nw_r, nw2.data(), nw2.size() * sizeof(double), cudaMemcpyHostToDevice));
            // K4 + K5 in one flat sweep: which entry of this list belongs to each K4 cell (none: -1).  A cell listed twice
            // would have K5 applied twice by the separate sweep, so the fusion is only taken for lists without repeats.
            if (d.nG > 0 && d.gcell_r && !getenv("MFLBM_NO_K45")) {
                if (!d.gk5_r && dev_alloc(ctx, &d.gk5_r, (size_t)d.nG, false)) return MFLBM_ERR_CUDA;
                const int dup = build_gk5(ctx, true);
                if (dup < 0) return MFLBM_ERR_CUDA;
                if (dup > 0) d.gk5_r = nullptr;  // (the allocation stays on ctx->allocs)
            }
        } else {
            d.fluid_cell_r = d.fluid_cell; d.fluid_nw_r = d.fluid_nw;
        }
        if (d.tcls[0]) {
            std::vector<int> tile(d.num_fluid), order, start;
            for (int n = 0; n < d.num_fluid; n++) tile[n] = g.tile_of(cell[n], d.ntx, d.nty);
            sort_by_tile(tile, d.ntiles, order, start);
            std::vector<int> cell2(d.num_fluid);
            std::vector<double> nw2((size_t)5 * d.num_fluid);
            for (int n = 0; n < d.num_fluid; n++) {
                cell2[n] = cell[order[n]];
                for (int m = 0; m < 5; m++) nw2[(size_t)m * d.num_fluid + n] = nw[(size_t)m * d.num_fluid + order[n]];
            }
            cell.swap(cell2); nw.swap(nw2);
            CU(cudaMemcpy(d.tf_start, start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
        CU(cudaMemcpy(d.fluid_cell, cell.data(), cell.size() * sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d.fluid_nw, nw.data(), nw.size() * sizeof(double), cudaMemcpyHostToDevice));
        if (d.tcls[0] && d.gk5_r && d.gcell && !d.k4_smem) {  // the same fusion for the tile-driven K4 (lists grouped by tile)
            if (!d.gk5 && dev_alloc(ctx, &d.gk5, (size_t)d.nG, false)) return MFLBM_ERR_CUDA;
            if (build_gk5(ctx, false) != 0) d.gk5 = nullptr;
        }
    }
    return MFLBM_OK;
}

extern "C" int mflbm_download(mflbm_ctx *ctx, const mflbm_arrays *h) {
    if (!ctx || !h) return fail(ctx, MFLBM_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    Dev &d = ctx->d;
    const int nz = d.g.nz;
    CU(cudaStreamSynchronize(ctx->s_main));
    for (int q = 0; q < 19; q++) {
        if (xfer_pdf(ctx, d.f[q], h->f[q], false, q)) return MFLBM_ERR_CUDA;
        if (d.multiphase && xfer_pdf(ctx, d.gg[q], h->g[q], false, q)) return MFLBM_ERR_CUDA;
    }
    if (d.multiphase) {
        if (h->phi && ctx->solid_phi_stale) {  // K3 was skipped on quiet tiles: give the caller the reference's values
            launch_phi_solid_refresh(ctx, ctx->s_main);
            CU(cudaGetLastError());
        }
        if (d.sparse && (h->cn_x || h->cn_y || h->cn_z || h->c_norm || h->curv)) {  // march kernel: the dense arrays on demand
            launch_dense_gradient(ctx, ctx->s_main);
            CU(cudaGetLastError());
        }
        if (d.sparse && h->curv) {  // not maintained per step on the sparse layout: evaluate K7 now (fluid nodes)
            launch_curvature(ctx, ctx->s_main);
            CU(cudaGetLastError());
        }
        if (xfer(ctx, d.phi, h->phi, 4, nz + 8, -3, false)) return MFLBM_ERR_CUDA;
        if (xfer(ctx, d.phi_old, h->phi_old, 4, nz + 8, -3, false)) return MFLBM_ERR_CUDA;
        if (xfer(ctx, d.cn_x, h->cn_x, 2, nz + 4, -1, false) || xfer(ctx, d.cn_y, h->cn_y, 2, nz + 4, -1, false) ||
            xfer(ctx, d.cn_z, h->cn_z, 2, nz + 4, -1, false) || xfer(ctx, d.c_norm, h->c_norm, 2, nz + 4, -1, false) ||
            xfer(ctx, d.curv, h->curv, 1, nz + 2, 0, false))
            return MFLBM_ERR_CUDA;
        if (xfer(ctx, d.g_convec, h->g_convec_bc, 1, 19, -3, false)) return MFLBM_ERR_CUDA;
        if (xfer(ctx, d.phi_convec, h->phi_convec_bc, 1, 1, -3, false)) return MFLBM_ERR_CUDA;
    }
    if (xfer(ctx, d.f_convec, h->f_convec_bc, 1, 19, -3, false)) return MFLBM_ERR_CUDA;
    if (h->u || h->v || h->w || h->rho) {
        if (!ctx->macro_alloc) return fail(ctx, MFLBM_ERR_STATE, "u,v,w,rho requested before mflbm_compute_macro_vars");
        if (xfer(ctx, d.u, h->u, 1, nz + 2, 0, false) || xfer(ctx, d.v, h->v, 1, nz + 2, 0, false) ||
            xfer(ctx, d.w, h->w, 1, nz + 2, 0, false) || xfer(ctx, d.rho, h->rho, 1, nz + 2, 0, false))
            return MFLBM_ERR_CUDA;
    }
    return MFLBM_OK;
}

// ---------------------------------------------------------------------------------------------------
// z-halo exchange between slabs over NVLink (NCCL p2p).  Whole padded planes are contiguous in the
// SoA grid, so they are sent in place: no pack/unpack kernels (the reference's MP/Mpi.F90:116-146,
// :239-269, :374-396, :497-519 degenerate to address arithmetic).  Plane pairs follow the reference:
//   pull (after an even step): own k=1 {6,14,13,18,17} -> lower neighbour's k=nz+1 ; own k=nz {5,11,12,15,16} -> upper k=0
//   push (after an odd step):  own k=0 {5,11,12,15,16} -> lower neighbour's k=nz   ; own k=nz+1 {6,14,13,18,17} -> upper k=1
//   phi: own 1..4 -> lower's nz+1..nz+4 ; own nz-3..nz -> upper's -3..0 (MP/Mpi.F90:624-631, :702-727)
// The wrap between the first and last slab is skipped on a non-periodic domain.
// ---------------------------------------------------------------------------------------------------
static int halo_exchange(mflbm_ctx *ctx, cudaStream_t st, bool push) {
    const Dev &d = ctx->d;
    const Grid &g = d.g;
    const mflbm_config &cfg = ctx->cfg;
    const int nz = g.nz;
    const bool has_lo = cfg.kper == 1 || cfg.idz != 0;
    const bool has_hi = cfg.kper == 1 || cfg.idz != cfg.npz - 1;
    static const int qM[5] = {6, 14, 13, 18, 17}, qP[5] = {5, 11, 12, 15, 16};
    const size_t n = (size_t)g.sxy;
    const int lo = ctx->peer_lo, hi = ctx->peer_hi;
    // Issue order per message class: send(lo), send(hi), recv(hi), recv(lo).  With two ranks on a periodic ring both
    // neighbours are the same peer and NCCL pairs sends and receives in issue order, so my "lo" message must meet
    // the peer's "hi" receive.
    if (d.sparse) {
        const size_t nb = (size_t)(d.multiphase ? 10 : 5) * g.nx * g.ny;
        for (int b = 0; b < 4; b++)
            if (!ctx->halo_buf[b]) {
                CU(cudaMalloc((void **)&ctx->halo_buf[b], nb * sizeof(double)));
                ctx->bytes += (long long)(nb * sizeof(double));
            }
        double **hb = ctx->halo_buf;
        launch_halo_pack(ctx, st, has_lo ? hb[0] : nullptr, has_hi ? hb[1] : nullptr, push, false);
        NC(ctx->nccl->GroupStart());
        if (has_lo) NC(ctx->nccl->Send(hb[0], nb, ncclDouble, lo, ctx->comm, st));
        if (has_hi) NC(ctx->nccl->Send(hb[1], nb, ncclDouble, hi, ctx->comm, st));
        if (has_hi) NC(ctx->nccl->Recv(hb[3], nb, ncclDouble, hi, ctx->comm, st));
        if (has_lo) NC(ctx->nccl->Recv(hb[2], nb, ncclDouble, lo, ctx->comm, st));
    } else {
        NC(ctx->nccl->GroupStart());
        for (int fl = 0; fl < (d.multiphase ? 2 : 1); fl++) {
            double *const *F = fl == 0 ? d.f : d.gg;
            for (int m = 0; m < 5; m++) {
                if (!push) {
                    if (has_lo) NC(ctx->nccl->Send(F[qM[m]] + g.plane_begin(1), n, ncclDouble, lo, ctx->comm, st));
                    if (has_hi) NC(ctx->nccl->Send(F[qP[m]] + g.plane_begin(nz), n, ncclDouble, hi, ctx->comm, st));
                    if (has_hi) NC(ctx->nccl->Recv(F[qM[m]] + g.plane_begin(nz + 1), n, ncclDouble, hi, ctx->comm, st));
                    if (has_lo) NC(ctx->nccl->Recv(F[qP[m]] + g.plane_begin(0), n, ncclDouble, lo, ctx->comm, st));
                } else {
                    if (has_lo) NC(ctx->nccl->Send(F[qP[m]] + g.plane_begin(0), n, ncclDouble, lo, ctx->comm, st));
                    if (has_hi) NC(ctx->nccl->Send(F[qM[m]] + g.plane_begin(nz + 1), n, ncclDouble, hi, ctx->comm, st));
                    if (has_hi) NC(ctx->nccl->Recv(F[qP[m]] + g.plane_begin(nz), n, ncclDouble, hi, ctx->comm, st));
                    if (has_lo) NC(ctx->nccl->Recv(F[qM[m]] + g.plane_begin(1), n, ncclDouble, lo, ctx->comm, st));
                }
            }
        }
    }
    if (d.multiphase) {  // phi stays on the dense grid in both layouts: four contiguous planes per direction
        if (has_lo) NC(ctx->nccl->Send(d.phi + g.plane_begin(1), 4 * n, ncclDouble, lo, ctx->comm, st));
        if (has_hi) NC(ctx->nccl->Send(d.phi + g.plane_begin(nz - 3), 4 * n, ncclDouble, hi, ctx->comm, st));
        if (has_hi) NC(ctx->nccl->Recv(d.phi + g.plane_begin(nz + 1), 4 * n, ncclDouble, hi, ctx->comm, st));
        if (has_lo) NC(ctx->nccl->Recv(d.phi + g.plane_begin(-3), 4 * n, ncclDouble, lo, ctx->comm, st));
    }
    NC(ctx->nccl->GroupEnd());
    if (d.sparse) launch_halo_pack(ctx, st, has_lo ? ctx->halo_buf[2] : nullptr, has_hi ? ctx->halo_buf[3] : nullptr, push, true);
    launch_halo_phi_classes(ctx, st, has_lo && d.bc_lo_dyn, has_hi && d.bc_hi_dyn);
    return 0;
}

static int check_launch(mflbm_ctx *ctx) {
    CU(cudaGetLastError());
    return 0;
}

// flush recorded event pairs into the accumulated collision-kernel time
static int prof_flush(mflbm_ctx *ctx) {
    for (size_t n = 0; n + 1 < ctx->prof_used; n += 2) {
        CU(cudaEventSynchronize(ctx->prof_ev[n + 1]));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ctx->prof_ev[n], ctx->prof_ev[n + 1]));
        ctx->prof_ms += ms;
        ctx->prof_launches++;
    }
    ctx->prof_used = 0;
    return 0;
}

static int collide_timed(mflbm_ctx *ctx, cudaStream_t s, bool odd, int k0, int k1) {
    if (!ctx->prof) {
        launch_collide(ctx, s, odd, k0, k1);
        return 0;
    }
    if (ctx->prof_used + 2 > 4096 && prof_flush(ctx)) return MFLBM_ERR_CUDA;
    while (ctx->prof_ev.size() < ctx->prof_used + 2) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        ctx->prof_ev.push_back(e);
    }
    const long long before = ctx->launches;
    CU(cudaEventRecord(ctx->prof_ev[ctx->prof_used], s));
    launch_collide(ctx, s, odd, k0, k1);
    CU(cudaEventRecord(ctx->prof_ev[ctx->prof_used + 1], s));
    if (ctx->launches > before) ctx->prof_used += 2;
    return 0;
}

// Speculative early gradient chain (sparse multiphase layout with quiet tiles).  The chain of step t (K3..K6 on the
// active tiles) is latency-bound and small, the collision is bandwidth-bound and large, and the chain only needs phi(t)
// around the active tiles.  So: collide the planes around the active tiles FIRST, then run the chain on its own
// high-priority stream BESIDE the collision of the remaining (far) planes.  Where the active tiles are is a guess -- the
// layer range of a recent step, read back asynchronously, widened by one layer.  The guess cannot break anything: the
// early chain reads only what the near-plane collision has already written (layers tz_lo-2..tz_hi+2), and after the far
// planes the tile update of the rest of the lattice decides on the device whether active tiles were missed, in which case
// the whole chain runs again on the complete lists, exactly like a plain step (kernels_gradient.cu, launch_chain_late).
struct SpecPlan {
    bool on;
    int tz_lo, tz_hi;  // tile layers of the early chain
    int k_lo, k_hi;    // planes collided first on their behalf (layers tz_lo-2..tz_hi+2)
};

static SpecPlan spec_plan(mflbm_ctx *ctx) {
    SpecPlan p{false, 0, 0, 0, 0};
    const Dev &d = ctx->d;
    const mflbm_config &cfg = ctx->cfg;
    if (!ctx->spec_enabled || !d.multiphase || !d.sparse || !d.use_tiles || d.wq_all || d.jper || ctx->march_on) return p;
    int force_lo = -1, force_hi = -1;
    if (const char *f = getenv("MFLBM_SPEC_FORCE"))  // test knob "lo:hi": pretend the active tiles are in these layers
        sscanf(f, "%d:%d", &force_lo, &force_hi);
    // most recent summary that has arrived (requested after the chain of an earlier step)
    int best = -1;
    // (mflbm_run queues steps far ahead of the device, so "recent" can be tens of steps old: the margin grows with the age)
    for (int sl = 0; sl < 2; sl++)
        if (ctx->sum_step[sl] >= 0 && cudaEventQuery(ctx->ev_sum[sl]) == cudaSuccess && (best < 0 || ctx->sum_step[sl] > ctx->sum_step[best]))
            best = sl;
    const int nz = d.g.nz;
    if (force_lo >= 0 && force_hi >= force_lo) {
        p.tz_lo = std::min(force_lo, d.ntz - 1);
        p.tz_hi = std::min(force_hi, d.ntz - 1);
    } else {
        if (best < 0) return p;
        const int *t = ctx->sum_host + 8 * best;
        if (t[2] != 0 || t[0] <= 0 || t[5] <= 0 || t[6] <= 0) return p;  // flat sweep last time / no active tile at all
        const int margin = 1 + (int)std::min<long long>(6, (ctx->step_count - ctx->sum_step[best]) / 32);
        p.tz_lo = std::max(0, (d.ntz - t[5]) - margin);
        p.tz_hi = std::min(d.ntz - 1, (t[6] - 1) + margin);
    }
    // layer L holds the planes k = 4L-3 .. 4L
    p.k_lo = std::max(1, 4 * (p.tz_lo - 2) - 3);
    p.k_hi = std::min(nz, 4 * (p.tz_hi + 2));
    if (p.k_hi < p.k_lo || 2 * (p.k_hi - p.k_lo + 1) > nz) return p;  // not worth it
    // ghost planes that only exist after the halo exchange / the periodic wrap of this step must stay out of reach
    const bool ghost_lo = (ctx->comm && (cfg.kper == 1 || cfg.idz != 0)) || (!ctx->comm && cfg.kper == 1);
    const bool ghost_hi = (ctx->comm && (cfg.kper == 1 || cfg.idz != cfg.npz - 1)) || (!ctx->comm && cfg.kper == 1);
    if (ghost_lo && p.tz_lo - 2 <= 0) return p;
    if (ghost_hi && p.tz_hi + 2 >= ((nz + 4) >> 2)) return p;
    p.on = true;
    return p;
}

// main_iteration_kernel for one ntime
static int step_impl(mflbm_ctx *ctx, int ntime) {
    const mflbm_config &cfg = ctx->cfg;
    if (!ctx->pdf_alloc) return fail(ctx, MFLBM_ERR_STATE, "mflbm_step before mflbm_upload");
    if (ctx->ckpt_mode == 2) return fail(ctx, MFLBM_ERR_STATE, "a checkpoint without device snapshot is pending: call mflbm_checkpoint_end first");
    const int nz = cfg.nz;
    const bool odd = (ntime % 2) != 0;
    cudaStream_t s = ctx->s_main;
    if (tiles_prepare(ctx, s)) return fail(ctx, MFLBM_ERR_CUDA, "quiet-tile setup failed");
    if (march_prepare(ctx, s) < 0) return fail(ctx, MFLBM_ERR_CUDA, "march kernel setup failed");
    const SpecPlan sp = spec_plan(ctx);
    ctx->step_count++;
    const int iz = cfg.iz_async > 0 ? cfg.iz_async : 1;
    if (sp.on) {
        // planes collided first: around the active tiles, the planes the inlet / outlet kernels work on, and (with
        // neighbours) the boundary slabs the halo exchange waits for
        std::vector<char> first((size_t)nz + 2, 0);
        for (int k = sp.k_lo; k <= sp.k_hi; k++) first[k] = 1;
        if (ctx->open_z && cfg.idz == 0) first[1] = first[std::min(2, nz)] = 1;
        if (ctx->open_z && cfg.idz == cfg.npz - 1) first[nz] = first[std::max(1, nz - 1)] = 1;
        if (ctx->comm)
            for (int k = 1; k <= iz; k++) first[k] = first[nz + 1 - k] = 1;
        auto collide_runs = [&](char want) -> int {
            for (int k = 1; k <= nz;) {
                if (first[k] != want) { k++; continue; }
                int e = k;
                while (e + 1 <= nz && first[e + 1] == want) e++;
                if (collide_timed(ctx, s, odd, k, e)) return MFLBM_ERR_CUDA;
                k = e + 1;
            }
            return 0;
        };
        if (collide_runs(1)) return MFLBM_ERR_CUDA;
        if (ctx->comm) {
            CU(cudaEventRecord(ctx->ev_slab, s));
            CU(cudaStreamWaitEvent(ctx->s_halo, ctx->ev_slab, 0));
            if (halo_exchange(ctx, ctx->s_halo, odd)) return MFLBM_ERR_NCCL;
            CU(cudaEventRecord(ctx->ev_halo, ctx->s_halo));
        }
        launch_bc(ctx, s, odd);  // inlet / outlet planes are done; they never meet the halo faces (open domain: end ranks only)
        CU(cudaEventRecord(ctx->ev_near, s));
        CU(cudaStreamWaitEvent(ctx->s_chain, ctx->ev_near, 0));
        launch_chain_early(ctx, ctx->s_chain, sp.tz_lo, sp.tz_hi);
        CU(cudaEventRecord(ctx->ev_chain, ctx->s_chain));
        if (collide_runs(0)) return MFLBM_ERR_CUDA;  // the far planes, beside the early chain
        CU(cudaEventRecord(ctx->ev_phi, s));
        ctx->ev_phi_valid = true;
        if (ctx->comm) CU(cudaStreamWaitEvent(s, ctx->ev_halo, 0));
        CU(cudaStreamWaitEvent(s, ctx->ev_chain, 0));
        launch_chain_late(ctx, s, sp.tz_lo, sp.tz_hi);
        ctx->spec_steps++;
    } else {
        if (ctx->comm) {
            // boundary slabs first, exchange on the high-priority halo stream while the interior runs
            // (MP/Main_multiphase.F90:358-387, :423-458)
            if (collide_timed(ctx, s, odd, 1, iz) || collide_timed(ctx, s, odd, nz - iz + 1, nz)) return MFLBM_ERR_CUDA;
            CU(cudaEventRecord(ctx->ev_slab, s));
            CU(cudaStreamWaitEvent(ctx->s_halo, ctx->ev_slab, 0));
            if (halo_exchange(ctx, ctx->s_halo, odd)) return MFLBM_ERR_NCCL;
            CU(cudaEventRecord(ctx->ev_halo, ctx->s_halo));
            if (collide_timed(ctx, s, odd, iz + 1, nz - iz)) return MFLBM_ERR_CUDA;
            CU(cudaStreamWaitEvent(s, ctx->ev_halo, 0));
            if (cfg.jper == 1) launch_wrap_y_phi(ctx, s);  // after the halo planes arrived: covers the x edges
        } else {
            if (collide_timed(ctx, s, odd, 1, nz)) return MFLBM_ERR_CUDA;
            if (cfg.kper == 1) launch_wrap_z(ctx, s, odd);
            if (cfg.jper == 1) launch_wrap_y_phi(ctx, s);  // after the z wrap: the x edges are images of images
        }
        CU(cudaEventRecord(ctx->ev_phi, s));  // nothing below writes phi at a fluid node (BC: ghost planes, K3: solid nodes)
        ctx->ev_phi_valid = true;
        launch_bc(ctx, s, odd);
        launch_color_gradient(ctx, s, true);
    }
    // Where the active tiles were: feeds the plan of later steps.  mflbm_run queues steps far ahead of the device, so a
    // request is only made every 32nd step (and early on), alternating between two slots: by the time a slot is requested
    // again (64 steps later) its previous copy has long arrived, and one of the two is always readable.
    if (ctx->d.use_tiles && ctx->spec_enabled && (ctx->step_count % 32 == 0 || ctx->step_count == 2 || ctx->step_count == 8)) {
        const int sl = ctx->sum_next;
        ctx->sum_next ^= 1;
        CU(cudaMemcpyAsync(ctx->sum_host + 8 * sl, ctx->d.tcount, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
        CU(cudaEventRecord(ctx->ev_sum[sl], s));
        ctx->sum_step[sl] = ctx->step_count;
    }
    return check_launch(ctx);
}

extern "C" int mflbm_step(mflbm_ctx *ctx, int ntime) {
    if (!ctx) return fail(nullptr, MFLBM_ERR_ARG, "null context");
    CU(cudaSetDevice(ctx->device));
    return step_impl(ctx, ntime);
}

extern "C" int mflbm_run(mflbm_ctx *ctx, int ntime0, int nsteps) {
    if (!ctx) return fail(nullptr, MFLBM_ERR_ARG, "null context");
    CU(cudaSetDevice(ctx->device));
    for (int n = 0; n < nsteps; n++) {
        const int rc = step_impl(ctx, ntime0 + n);
        if (rc) return rc;
    }
    return MFLBM_OK;
}

// ---------------------------------------------------------------------------------------------------
// streamed steps (include/mflbm.h "streamed steps"): per-step host input and per-step result without a host stall
// ---------------------------------------------------------------------------------------------------
static int stream_setup(mflbm_ctx *ctx) {
    if (ctx->win_stage[0]) return 0;
    Dev &d = ctx->d;
    const size_t nplane = (size_t)d.g.sxy + 32, npack = (size_t)(d.g.nx + 2) * (d.g.ny + 2);
    ctx->win_dev[0] = d.w_in;
    if (dev_alloc(ctx, &ctx->win_dev[1], nplane)) return MFLBM_ERR_CUDA;
    for (int b = 0; b < 2; b++) {
        if (dev_alloc(ctx, &ctx->win_stage[b], npack, false) || dev_alloc(ctx, &ctx->res_dev[b], (size_t)ctx->red_len)) return MFLBM_ERR_CUDA;
        CU(cudaMallocHost((void **)&ctx->res_host[b], ctx->red_len * sizeof(double)));
        CU(cudaEventCreateWithFlags(&ctx->ev_win[b], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_res[b], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_step[b], cudaEventDisableTiming));
    }
    return 0;
}

// result slot b -> sums (waits for its read-back only)
static int stream_result(mflbm_ctx *ctx, int b, double *v1, double *v2) {
    CU(cudaEventSynchronize(ctx->ev_res[b]));
    double a = 0, c = 0;
    const int np = ctx->res_np[b];
    for (int k = 0; k < np; k++) { a += ctx->res_host[b][k]; c += ctx->res_host[b][np + k]; }
    if (v1) *v1 = a;
    if (v2) *v2 = c;
    return 0;
}

extern "C" int mflbm_step_streamed(mflbm_ctx *ctx, int ntime, const double *w_in_host, double *v1, double *v2, int *have_prev) {
    if (!ctx) return fail(nullptr, MFLBM_ERR_ARG, "null context");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->pdf_alloc) return fail(ctx, MFLBM_ERR_STATE, "mflbm_step_streamed before mflbm_upload");
    if (stream_setup(ctx)) return MFLBM_ERR_CUDA;
    Dev &d = ctx->d;
    const int b = (int)(ctx->stream_count & 1), prev = b ^ 1;
    if (w_in_host) {
        // this step's inlet profile: host -> staging -> padded plane, on the copy stream, beside the previous step; the
        // buffers of slot b were last read by the streamed step before the previous one
        if (ctx->stream_count >= 2) {
            CU(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_step[b], 0));
        } else {  // ... or by whatever ran before the first streamed step
            CU(cudaEventRecord(ctx->ev_win[b], ctx->s_main));
            CU(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_win[b], 0));
        }
        const size_t npack = (size_t)(d.g.nx + 2) * (d.g.ny + 2);
        CU(cudaMemcpyAsync(ctx->win_stage[b], w_in_host, npack * sizeof(double), cudaMemcpyHostToDevice, ctx->s_copy));
        launch_repack(ctx, ctx->s_copy, ctx->win_dev[b], ctx->win_stage[b], 1, 1, -3, true);
        CU(cudaEventRecord(ctx->ev_win[b], ctx->s_copy));
        CU(cudaStreamWaitEvent(ctx->s_main, ctx->ev_win[b], 0));
        d.w_in = ctx->win_dev[b];
    }
    // the saturation kernel of the previous streamed step reads phi beside that step's gradient chain: the collision of
    // this step must not overwrite it
    if (ctx->stream_count >= 1 && d.multiphase) CU(cudaStreamWaitEvent(ctx->s_main, ctx->ev_res[prev], 0));
    const int rc = step_impl(ctx, ntime);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev_step[b], ctx->s_main));
    if (d.multiphase) {  // cal_saturation of this step, read back asynchronously (like mflbm_cal_saturation: after ev_phi, second stream)
        cudaStream_t st = ctx->s_main;
        if (d.sparse && ctx->ev_phi_valid) {
            st = ctx->s_halo;
            CU(cudaStreamWaitEvent(st, ctx->ev_phi, 0));
        }
        ctx->res_np[b] = launch_saturation(ctx, st, ctx->res_dev[b]);
        CU(cudaMemcpyAsync(ctx->res_host[b], ctx->res_dev[b], 2 * ctx->res_np[b] * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(ctx->ev_res[b], st));
    }
    if (have_prev) *have_prev = 0;
    if (ctx->stream_count >= 1) {  // pace the host one step behind the device and hand out the previous step's result
        if (d.multiphase) {
            if (stream_result(ctx, prev, v1, v2)) return MFLBM_ERR_CUDA;
            if (have_prev) *have_prev = 1;
        } else {
            CU(cudaEventSynchronize(ctx->ev_step[prev]));
        }
    }
    ctx->stream_count++;
    return check_launch(ctx);
}

extern "C" int mflbm_stream_flush(mflbm_ctx *ctx, double *v1, double *v2) {
    if (!ctx) return fail(nullptr, MFLBM_ERR_ARG, "null context");
    CU(cudaSetDevice(ctx->device));
    if (ctx->stream_count < 1) return fail(ctx, MFLBM_ERR_STATE, "no streamed step in flight");
    const int last = (int)((ctx->stream_count - 1) & 1);
    if (ctx->d.multiphase) {
        if (stream_result(ctx, last, v1, v2)) return MFLBM_ERR_CUDA;
        CU(cudaStreamWaitEvent(ctx->s_main, ctx->ev_res[last], 0));  // whatever the caller queues next comes after the read-back
    }
    CU(cudaStreamSynchronize(ctx->s_main));
    ctx->stream_count = 0;
    return MFLBM_OK;
}

extern "C" int mflbm_color_gradient(mflbm_ctx *ctx) {
    if (!ctx) return fail(nullptr, MFLBM_ERR_ARG, "null context");
    CU(cudaSetDevice(ctx->device));
    if (tiles_prepare(ctx, ctx->s_main)) return fail(ctx, MFLBM_ERR_CUDA, "quiet-tile setup failed");
    if (march_prepare(ctx, ctx->s_main) < 0) return fail(ctx, MFLBM_ERR_CUDA, "march kernel setup failed");
    launch_color_gradient(ctx, ctx->s_main, false);
    return check_launch(ctx);
}

extern "C" int mflbm_compute_macro_vars(mflbm_ctx *ctx) {
    if (!ctx) return fail(nullptr, MFLBM_ERR_ARG, "null context");
    CU(cudaSetDevice(ctx->device));
    if (ensure_macro(ctx)) return MFLBM_ERR_CUDA;
    // phi on solid nodes of 1..n is zeroed below (MP/Misc.F90:424); the listed solid nodes outside 1..n keep the last
    // K3 value, so refresh the ones quiet tiles skipped first, and let the next gradient chain evaluate every tile
    if (ctx->d.sparse && ctx->d.multiphase) launch_dense_gradient(ctx, ctx->s_main);  // the CSF force term reads n, |grad phi|
    if (ctx->solid_phi_stale) launch_phi_solid_refresh(ctx, ctx->s_main);
    launch_macro(ctx, ctx->s_main);
    launch_tiles_reset(ctx, ctx->s_main);
    return check_launch(ctx);
}

// ---------------------------------------------------------------------------------------------------
// asynchronous output staging (include/mflbm.h "asynchronous output staging")
// ---------------------------------------------------------------------------------------------------
extern "C" int mflbm_output_begin(mflbm_ctx *ctx, int what) {
    if (!ctx) return fail(nullptr, MFLBM_ERR_ARG, "null context");
    CU(cudaSetDevice(ctx->device));
    Dev &d = ctx->d;
    if (!ctx->pdf_alloc) return fail(ctx, MFLBM_ERR_STATE, "mflbm_output_begin before mflbm_upload");
    if (ctx->out_pending) return fail(ctx, MFLBM_ERR_STATE, "an output is already in flight: call mflbm_output_end first");
    if (!(what & (MFLBM_OUT_PHI | MFLBM_OUT_MACRO))) return fail(ctx, MFLBM_ERR_ARG, "nothing requested");
    if ((what & MFLBM_OUT_PHI) && !d.multiphase) return fail(ctx, MFLBM_ERR_STATE, "phi is a multiphase field");
    const Grid &g = d.g;
    cudaStream_t s = ctx->s_main;
    if (what & MFLBM_OUT_MACRO) {  // save_macro computes the macroscopic variables first (MP/IO_multiphase.F90:684)
        const int rc = mflbm_compute_macro_vars(ctx);
        if (rc) return rc;
    } else if (ctx->solid_phi_stale) {  // like mflbm_download: the reference's phi on every listed solid node
        launch_phi_solid_refresh(ctx, s);
    }
    // field f: 0 phi (ghost 4), 1..4 u, v, w, rho (ghost 1)
    double *src[5] = {d.phi, d.u, d.v, d.w, d.rho};
    for (int f = 0; f < 5; f++) {
        const bool want = f == 0 ? (what & MFLBM_OUT_PHI) != 0 : (what & MFLBM_OUT_MACRO) != 0;
        if (!want) continue;
        const int o = f == 0 ? 4 : 1;
        const size_t n = (size_t)(g.nx + 2 * o) * (g.ny + 2 * o) * (g.nz + 2 * o);
        if (!ctx->out_dev[f]) {
            CU(cudaMalloc((void **)&ctx->out_dev[f], n * sizeof(double)));
            CU(cudaMallocHost((void **)&ctx->out_host[f], n * sizeof(double)));
            ctx->out_elems[f] = n;
        }
        launch_repack(ctx, s, src[f], ctx->out_dev[f], o, g.nz + 2 * o, 1 - o, false);  // snapshot in the caller's layout
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev_out, s));
    CU(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_out, 0));
    for (int f = 0; f < 5; f++) {
        const bool want = f == 0 ? (what & MFLBM_OUT_PHI) != 0 : (what & MFLBM_OUT_MACRO) != 0;
        if (want)
            CU(cudaMemcpyAsync(ctx->out_host[f], ctx->out_dev[f], ctx->out_elems[f] * sizeof(double), cudaMemcpyDeviceToHost, ctx->s_copy));
    }
    CU(cudaEventRecord(ctx->ev_out_done, ctx->s_copy));
    ctx->out_pending = what & (MFLBM_OUT_PHI | MFLBM_OUT_MACRO);
    return MFLBM_OK;
}

extern "C" int mflbm_output_end(mflbm_ctx *ctx, const mflbm_arrays *h) {
    if (!ctx || !h) return fail(ctx, MFLBM_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->out_pending) return fail(ctx, MFLBM_ERR_STATE, "no output in flight");
    CU(cudaEventSynchronize(ctx->ev_out_done));
    double *dst[5] = {h->phi, h->u, h->v, h->w, h->rho};
    for (int f = 0; f < 5; f++) {
        const bool have = f == 0 ? (ctx->out_pending & MFLBM_OUT_PHI) != 0 : (ctx->out_pending & MFLBM_OUT_MACRO) != 0;
        if (have && dst[f]) memcpy(dst[f], ctx->out_host[f], ctx->out_elems[f] * sizeof(double));
    }
    ctx->out_pending = 0;
    return MFLBM_OK;
}

// ---------------------------------------------------------------------------------------------------
// checkpoint staging (include/mflbm.h "checkpoint staging")
// ---------------------------------------------------------------------------------------------------
static size_t pdf_elems(const Dev &d, int q) {
    return d.sparse ? (size_t)d.nAct + 64 + (size_t)d.nlink[OPC(q)] : (size_t)d.g.ntot;
}

static void ckpt_release(mflbm_ctx *ctx, cudaStream_t st) {
    double **one[] = {&ctx->ckpt_phi, &ctx->ckpt_fc, &ctx->ckpt_gc, &ctx->ckpt_pc};
    for (double **p : one)
        if (*p) { cudaFreeAsync(*p, st); *p = nullptr; }
    for (int q = 0; q < 19; q++) {
        if (ctx->ckpt_f[q]) { cudaFreeAsync(ctx->ckpt_f[q], st); ctx->ckpt_f[q] = nullptr; }
        if (ctx->ckpt_g[q]) { cudaFreeAsync(ctx->ckpt_g[q], st); ctx->ckpt_g[q] = nullptr; }
    }
}

extern "C" int mflbm_checkpoint_begin(mflbm_ctx *ctx) {
    if (!ctx) return fail(nullptr, MFLBM_ERR_ARG, "null context");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->pdf_alloc) return fail(ctx, MFLBM_ERR_STATE, "mflbm_checkpoint_begin before mflbm_upload");
    if (ctx->ckpt_mode) return fail(ctx, MFLBM_ERR_STATE, "a checkpoint is already pending: call mflbm_checkpoint_end first");
    Dev &d = ctx->d;
    cudaStream_t s = ctx->s_main;
    if (d.multiphase && ctx->solid_phi_stale) launch_phi_solid_refresh(ctx, s);  // the reference's phi on every listed solid node
    const size_t plane = (size_t)d.g.sxy + 32, plane19 = (size_t)19 * d.g.sxy + 32;  // sizes of the plane fields as allocated
    size_t need = 0;
    for (int q = 0; q < 19; q++) need += pdf_elems(d, q) * (d.multiphase ? 2 : 1);
    need += (d.multiphase ? (size_t)d.g.ntot + plane19 + plane : 0) + plane19;
    need *= sizeof(double);
    size_t fr = 0, tot = 0;
    CU(cudaMemGetInfo(&fr, &tot));
    const char *force = getenv("MFLBM_CKPT_DIRECT");  // developer / test switch: behave as if the snapshot did not fit
    bool staged = fr > need + ((size_t)1 << 30) && !(force && atoi(force));
    if (staged) {
        auto snap = [&](double **dst, const double *src, size_t n) -> bool {
            if (cudaMallocAsync((void **)dst, n * sizeof(double), s) != cudaSuccess) { *dst = nullptr; return false; }
            return cudaMemcpyAsync(*dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, s) == cudaSuccess;
        };
        bool ok = true;
        for (int q = 0; q < 19 && ok; q++) {
            ok = snap(&ctx->ckpt_f[q], d.f[q], pdf_elems(d, q));
            if (ok && d.multiphase) ok = snap(&ctx->ckpt_g[q], d.gg[q], pdf_elems(d, q));
        }
        if (ok) ok = snap(&ctx->ckpt_fc, d.f_convec, plane19);
        if (ok && d.multiphase)
            ok = snap(&ctx->ckpt_phi, d.phi, (size_t)d.g.ntot) && snap(&ctx->ckpt_gc, d.g_convec, plane19) && snap(&ctx->ckpt_pc, d.phi_convec, plane);
        if (!ok) {  // allocation failed after all: fall back to the frozen-context mode
            cudaGetLastError();
            ckpt_release(ctx, s);
            staged = false;
        }
    }
    if (staged) {
        CU(cudaEventRecord(ctx->ev_ckpt, s));
        ctx->ckpt_mode = 1;
        return MFLBM_CKPT_STAGED;
    }
    ctx->ckpt_mode = 2;
    return MFLBM_CKPT_DIRECT;
}

extern "C" int mflbm_checkpoint_fetch(mflbm_ctx *ctx, const mflbm_arrays *h) {
    if (!ctx || !h) return fail(ctx, MFLBM_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->ckpt_mode) return fail(ctx, MFLBM_ERR_STATE, "no checkpoint pending");
    Dev &d = ctx->d;
    const bool staged = ctx->ckpt_mode == 1;
    // staged: copy stream, behind the snapshot, overlapping whatever steps are queued on the compute stream;
    // direct: the live arrays of the (frozen) context on the compute stream
    cudaStream_t st = staged ? ctx->s_copy : ctx->s_main;
    if (staged) CU(cudaStreamWaitEvent(st, ctx->ev_ckpt, 0));
    else CU(cudaStreamSynchronize(ctx->s_main));
    for (int q = 0; q < 19; q++) {
        if (xfer_pdf(ctx, staged ? ctx->ckpt_f[q] : d.f[q], h->f[q], false, q, st)) return MFLBM_ERR_CUDA;
        if (d.multiphase && xfer_pdf(ctx, staged ? ctx->ckpt_g[q] : d.gg[q], h->g[q], false, q, st)) return MFLBM_ERR_CUDA;
    }
    if (d.multiphase) {
        if (xfer(ctx, staged ? ctx->ckpt_phi : d.phi, h->phi, 4, d.g.nz + 8, -3, false, st)) return MFLBM_ERR_CUDA;
        if (xfer(ctx, staged ? ctx->ckpt_gc : d.g_convec, h->g_convec_bc, 1, 19, -3, false, st)) return MFLBM_ERR_CUDA;
        if (xfer(ctx, staged ? ctx->ckpt_pc : d.phi_convec, h->phi_convec_bc, 1, 1, -3, false, st)) return MFLBM_ERR_CUDA;
    }
    if (xfer(ctx, staged ? ctx->ckpt_fc : d.f_convec, h->f_convec_bc, 1, 19, -3, false, st)) return MFLBM_ERR_CUDA;
    return MFLBM_OK;
}

extern "C" int mflbm_checkpoint_end(mflbm_ctx *ctx) {
    if (!ctx) return fail(nullptr, MFLBM_ERR_ARG, "null context");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->ckpt_mode) return fail(ctx, MFLBM_ERR_STATE, "no checkpoint pending");
    if (ctx->ckpt_mode == 1) {
        CU(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_ckpt, 0));  // never release ahead of the snapshot copies
        ckpt_release(ctx, ctx->s_copy);
    }
    ctx->ckpt_mode = 0;
    return MFLBM_OK;
}

static int fetch_red(mflbm_ctx *ctx, int n) {
    CU(cudaMemcpyAsync(ctx->red_host, ctx->red_dev, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->s_main));
    CU(cudaStreamSynchronize(ctx->s_main));
    return 0;
}

extern "C" int mflbm_monitor(mflbm_ctx *ctx, double *tk, int tk_len) {
    if (!ctx || !tk) return fail(ctx, MFLBM_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    const int nz = ctx->cfg.nz;
    const bool mp = ctx->d.multiphase;
    const int need = mp ? 7 * nz + 3 : 2 * nz + 1;
    if (tk_len < need) return fail(ctx, MFLBM_ERR_ARG, "tk buffer too small (MP/Init_multiphase.F90:799: 7*nz+3)");
    int rc = mflbm_compute_macro_vars(ctx);
    if (rc) return rc;
    launch_monitor(ctx, ctx->s_main, ctx->red_dev);
    if (check_launch(ctx) || fetch_red(ctx, 10 * nz)) return MFLBM_ERR_CUDA;
    const double *r = ctx->red_host;
    double umax = 0, usq1 = 0, usq2 = 0;
    for (int k = 0; k < nz; k++) {
        if (umax < r[7 * nz + k]) umax = r[7 * nz + k];
        usq1 += r[8 * nz + k];
        usq2 += r[9 * nz + k];
    }
    if (mp) {
        memcpy(tk, r, 7 * nz * sizeof(double));
        tk[7 * nz] = umax; tk[7 * nz + 1] = usq1; tk[7 * nz + 2] = usq2;
    } else {
        memcpy(tk, r, nz * sizeof(double));               // fl
        memcpy(tk + nz, r + 6 * nz, nz * sizeof(double)); // pre
        tk[2 * nz] = umax;
    }
    return MFLBM_OK;
}

extern "C" int mflbm_cal_saturation(mflbm_ctx *ctx, double *v1, double *v2) {
    if (!ctx || !v1 || !v2) return fail(ctx, MFLBM_ERR_ARG, "null argument");
    if (!ctx->d.multiphase) return fail(ctx, MFLBM_ERR_STATE, "multiphase only");
    CU(cudaSetDevice(ctx->device));
    // Sparse layout: the sum runs over the fluid nodes only, whose phi is final as soon as the collision kernel of the
    // last step is done -- it does not have to queue behind that step's gradient chain.  It runs on the second stream
    // after ev_phi and overlaps the (latency-bound, low-occupancy) chain kernels.
    cudaStream_t st = ctx->s_main;
    if (ctx->d.sparse && ctx->ev_phi_valid) {
        st = ctx->s_halo;
        CU(cudaStreamWaitEvent(st, ctx->ev_phi, 0));
    }
    const int np = launch_saturation(ctx, st, ctx->red_dev);
    if (check_launch(ctx)) return MFLBM_ERR_CUDA;
    CU(cudaMemcpyAsync(ctx->red_host, ctx->red_dev, 2 * np * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    double a = 0, b = 0;
    for (int k = 0; k < np; k++) { a += ctx->red_host[k]; b += ctx->red_host[np + k]; }
    *v1 = a; *v2 = b;
    return MFLBM_OK;
}

extern "C" int mflbm_monitor_breakthrough(mflbm_ctx *ctx, int32_t *count) {
    if (!ctx || !count) return fail(ctx, MFLBM_ERR_ARG, "null argument");
    if (!ctx->d.multiphase) return fail(ctx, MFLBM_ERR_STATE, "multiphase only");
    CU(cudaSetDevice(ctx->device));
    *count = 0;
    if (ctx->cfg.idz != ctx->cfg.npz - 1) return MFLBM_OK;
    const int ny = ctx->cfg.ny;
    launch_breakthrough(ctx, ctx->s_main, ctx->red_dev);
    if (check_launch(ctx) || fetch_red(ctx, ny)) return MFLBM_ERR_CUDA;
    long long t = 0;
    for (int j = 0; j < ny; j++) t += (long long)ctx->red_host[j];
    *count = (int32_t)t;
    return MFLBM_OK;
}

extern "C" int mflbm_monitor_steady_phasefield(mflbm_ctx *ctx, double *umax_sq, double *d_phi_max) {
    if (!ctx || !umax_sq || !d_phi_max) return fail(ctx, MFLBM_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->d.multiphase) return fail(ctx, MFLBM_ERR_STATE, "multiphase only");
    if (ensure_phi_old(ctx)) return MFLBM_ERR_CUDA;
    if (ctx->solid_phi_stale) launch_phi_solid_refresh(ctx, ctx->s_main);
    int rc = mflbm_compute_macro_vars(ctx);
    if (rc) return rc;
    const int nz = ctx->cfg.nz;
    launch_steady_phasefield(ctx, ctx->s_main, ctx->red_dev);
    if (check_launch(ctx) || fetch_red(ctx, 2 * nz)) return MFLBM_ERR_CUDA;
    double a = 0, b = 0;
    for (int k = 0; k < nz; k++) {
        if (a < ctx->red_host[k]) a = ctx->red_host[k];
        if (b < ctx->red_host[nz + k]) b = ctx->red_host[nz + k];
    }
    *umax_sq = a; *d_phi_max = b;
    return MFLBM_OK;
}

extern "C" int mflbm_monitor_steady_capillarypressure(mflbm_ctx *ctx, double *umax_sq, double *pre_w, double *pre_nw,
                                                      int32_t *i_w, int32_t *i_nw) {
    if (!ctx || !umax_sq || !pre_w || !pre_nw || !i_w || !i_nw) return fail(ctx, MFLBM_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->d.multiphase) return fail(ctx, MFLBM_ERR_STATE, "multiphase only");
    int rc = mflbm_compute_macro_vars(ctx);
    if (rc) return rc;
    const int nz = ctx->cfg.nz;
    launch_steady_cappres(ctx, ctx->s_main, ctx->red_dev);
    if (check_launch(ctx) || fetch_red(ctx, 5 * nz)) return MFLBM_ERR_CUDA;
    const double *r = ctx->red_host;
    double um = 0, pw = 0, pnw = 0;
    long long cw = 0, cnw = 0;
    for (int k = 0; k < nz; k++) {
        if (um < r[k]) um = r[k];
        pw += r[nz + k]; pnw += r[2 * nz + k];
        cw += (long long)r[3 * nz + k]; cnw += (long long)r[4 * nz + k];
    }
    *umax_sq = um; *pre_w = pw; *pre_nw = pnw; *i_w = (int32_t)cw; *i_nw = (int32_t)cnw;
    return MFLBM_OK;
}

extern "C" int mflbm_set_parameter(mflbm_ctx *ctx, const char *name, double value) {
    if (!ctx || !name) return fail(ctx, MFLBM_ERR_ARG, "null argument");
    Dev &d = ctx->d;
    if (!strcmp(name, "force_Z") || !strcmp(name, "force_z")) d.force_Z = value;
    else if (!strcmp(name, "rho_in")) d.rho_in = value;
    else if (!strcmp(name, "rho_out")) d.rho_out = value;
    else if (!strcmp(name, "uin_avg")) d.uin_avg = value;
    else if (!strcmp(name, "phi_inlet")) d.phi_inlet = value;
    else if (!strcmp(name, "sa_inject")) d.sa_inject = value;
    else if (!strcmp(name, "relaxation")) d.relaxation = value;
    else if (!strcmp(name, "quiet_tiles")) {
        // developer / measurement switch: 0 = evaluate the colour-gradient chain on every node (no quiet-tile skipping),
        // 1 = back to the default.  Results are identical either way; only contexts that were set up with tiles can switch.
        if (!d.tcls[0]) return fail(ctx, MFLBM_ERR_STATE, "this context has no quiet-tile state (dense layout, singlephase or MFLBM_NO_TILES)");
        CU(cudaSetDevice(ctx->device));
        CU(cudaStreamSynchronize(ctx->s_main));
        d.use_tiles = value != 0.0 ? 1 : 0;
        d.wq_all = 1;
        ctx->tiles_static_ready = false;
    }
    else return fail(ctx, MFLBM_ERR_ARG, std::string("unknown parameter ") + name);
    return MFLBM_OK;
}

extern "C" int mflbm_sync(mflbm_ctx *ctx) {
    if (!ctx) return fail(nullptr, MFLBM_ERR_ARG, "null context");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->s_halo));
    CU(cudaStreamSynchronize(ctx->s_main));
    return MFLBM_OK;
}

extern "C" int mflbm_timer_start(mflbm_ctx *ctx) {
    if (!ctx) return fail(nullptr, MFLBM_ERR_ARG, "null context");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev_t0, ctx->s_main));
    return MFLBM_OK;
}

extern "C" int mflbm_timer_stop(mflbm_ctx *ctx, double *elapsed_ms) {
    if (!ctx || !elapsed_ms) return fail(ctx, MFLBM_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(ctx->ev_t1, ctx->s_main));
    CU(cudaEventSynchronize(ctx->ev_t1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
    *elapsed_ms = ms;
    return MFLBM_OK;
}

extern "C" int mflbm_profile(mflbm_ctx *ctx, int enable) {
    if (!ctx) return fail(nullptr, MFLBM_ERR_ARG, "null context");
    CU(cudaSetDevice(ctx->device));
    if (prof_flush(ctx)) return MFLBM_ERR_CUDA;
    ctx->prof = enable != 0;
    return MFLBM_OK;
}

extern "C" int mflbm_profile_read(mflbm_ctx *ctx, double *collide_ms, long long *collide_launches) {
    if (!ctx || !collide_ms || !collide_launches) return fail(ctx, MFLBM_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    if (prof_flush(ctx)) return MFLBM_ERR_CUDA;
    *collide_ms = ctx->prof_ms;
    *collide_launches = ctx->prof_launches;
    ctx->prof_ms = 0;
    ctx->prof_launches = 0;
    return MFLBM_OK;
}

extern "C" int mflbm_tile_stats(mflbm_ctx *ctx, long long *ntiles, long long *nquiet) {
    if (!ctx || !ntiles || !nquiet) return fail(ctx, MFLBM_ERR_ARG, "null argument");
    *ntiles = 0;
    *nquiet = 0;
    if (!ctx->d.use_tiles) return MFLBM_OK;
    CU(cudaSetDevice(ctx->device));
    std::vector<unsigned char> h((size_t)ctx->d.ntiles);
    CU(cudaStreamSynchronize(ctx->s_main));
    CU(cudaMemcpy(h.data(), ctx->d.tquiet, h.size(), cudaMemcpyDeviceToHost));
    long long q = 0;
    for (unsigned char b : h) q += b;
    *ntiles = ctx->d.ntiles;
    *nquiet = q;
    return MFLBM_OK;
}

extern "C" int mflbm_chain_info(mflbm_ctx *ctx, int *fused, int *reject_mask) {
    if (!ctx) return fail(nullptr, MFLBM_ERR_ARG, "null context");
    CU(cudaSetDevice(ctx->device));
    if (ctx->pdf_alloc && march_prepare(ctx, ctx->s_main) < 0) return fail(ctx, MFLBM_ERR_CUDA, "march kernel setup failed");
    if (fused) *fused = (ctx->march_on && ctx->march_ready) ? 1 : 0;
    if (reject_mask) *reject_mask = ctx->march_reject;
    return MFLBM_OK;
}

extern "C" int mflbm_chain_selfcheck(mflbm_ctx *ctx, long long *mismatches) {
    if (!ctx || !mismatches) return fail(ctx, MFLBM_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->pdf_alloc) return fail(ctx, MFLBM_ERR_STATE, "mflbm_chain_selfcheck before mflbm_upload");
    const long long r = chain_selfcheck(ctx, ctx->s_main);
    if (r < 0) return fail(ctx, MFLBM_ERR_CUDA, "self-check of the colour-gradient chain failed to run");
    *mismatches = r;
    return check_launch(ctx);
}

extern "C" long long mflbm_launch_count(const mflbm_ctx *ctx) { return ctx ? ctx->launches : 0; }
// internal (tests, bench.py): how many steps of this context ran with the speculative early gradient chain
extern "C" long long mflbmx_spec_steps(const mflbm_ctx *ctx) { return ctx ? ctx->spec_steps : 0; }
// internal: the two summary slots (copies of Dev::tcount[0..7]) as last read back, for diagnostics
extern "C" void mflbmx_spec_info(const mflbm_ctx *ctx, int out[16]) {
    for (int n = 0; n < 16; n++) out[n] = (ctx && ctx->sum_host) ? ctx->sum_host[n] : 0;
}
extern "C" long long mflbm_device_bytes(const mflbm_ctx *ctx) { return ctx ? ctx->bytes : 0; }

// Internal (not part of include/mflbm.h): host-only self-test of the node numbering + compressed adjacency, callable
// without a GPU (tests/test_adjacency.py).  Returns 0 when every (node, direction) decodes to the direct lookup and the
// link slots of each direction are a permutation-free ranking; fills counts[0..4] = nA, nAct, total links, irregular rows,
// warps with at least one irregular row (counts must hold 8 values).
extern "C" int mflbmx_adjacency_selftest(int nx, int ny, int nz, const int8_t *walls, long long *counts) {
    Grid g;
    g.nx = nx; g.ny = ny; g.nz = nz;
    g.sx = (nx + 8 + 15) / 16 * 16;
    g.sxy = g.sx * (ny + 8);
    g.base = 16;
    g.ntot = 16 + g.sxy * (nz + 8) + 16;
    g.set_magic();
    {   // the multiply-high cell decomposition must agree with the plain division on every cell of this grid
        long long bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
        for (long long c = g.base - 4; c < (long long)g.ntot; c++) {
            unsigned ix, jy, kz;
            g.coords3((int)c, ix, jy, kz);
            const unsigned r = (unsigned)(c - (g.base - 4));
            bad += (kz != r / (unsigned)g.sxy) || (jy != (r % (unsigned)g.sxy) / (unsigned)g.sx) || (ix != (r % (unsigned)g.sxy) % (unsigned)g.sx);
        }
        if (bad) {
            g_err = "Grid::coords3 disagrees with integer division";
            return MFLBM_ERR_STATE;
        }
    }
    HostActive H;
    std::string err;
    if (build_active_host(g, walls, H, err, true)) {
        g_err = err;
        return MFLBM_ERR_STATE;
    }
    // link slots: every link lane of direction q must get a distinct slot in [nAct, nAct + nlink[q])
    for (int q = 1; q < 19; q++) {
        std::vector<char> seen((size_t)H.nlink[q] + 1, 0);
        for (int n = 0; n < H.nA; n++) {
            const int w = n >> 5, l = n & 31;
            const uint4 *rec = &H.adj[(size_t)w * MFLBM_ADJ_REC];
            if (!((rec[1 + 2 * (q - 1)].x >> l) & 1u)) continue;
            const long long slot = (long long)adj_lookup(rec, H.adjfull.data(), q, l, (int)H.nAct) - H.nAct;
            if (slot < 0 || slot >= H.nlink[q] || seen[(size_t)slot]) {
                g_err = "link slot collision";
                return MFLBM_ERR_STATE;
            }
            seen[(size_t)slot] = 1;
        }
    }
    if (counts) {
        long long links = 0;
        for (int q = 1; q < 19; q++) links += H.nlink[q];
        counts[0] = H.nA; counts[1] = H.nAct; counts[2] = links; counts[3] = (long long)H.adjfull.size() / 32;
        long long iw = 0;
        for (int w = 0; w < (H.nA + 31) / 32; w++) iw += H.adj[(size_t)w * MFLBM_ADJ_REC].x != 0;
        counts[4] = iw;
    }
    return MFLBM_OK;
}
