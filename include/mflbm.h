/*
 * mflbm.h -- C ABI of the B200-native MF-LBM time-step hot path.
 *
 * This is the drop-in boundary: the reference (lanl/MF-LBM, Fortran 90 + OpenACC + MPI) keeps its
 * driver, control files, geometry loading and output formats and calls these entry points through
 * ISO_C_BINDING in place of its kernel subroutines (binding: fortran/mflbm_iso_c.f90).
 * The reference has no FFI of its own; every export names the reference routine / call site it
 * replaces ("MP/" = multiphase_3D/0.src/, "SP/" = singlephase_3D/0.src/).
 *
 * Conventions
 *  - every call returns int: 0 = ok, <0 = error (text via mflbm_last_error); the Fortran wrapper maps
 *    non-zero to MPI_Barrier + mpi_abort like MP/IO_multiphase.F90:543-545.
 *  - host arrays are caller-owned, contiguous, column-major (i fastest) with the reference's ghost
 *    extents (MP/Init_multiphase.F90:594-658): PDFs/u/v/w/rho/curv (0:nx+1,0:ny+1,0:nz+1);
 *    cn_x,cn_y,cn_z,c_norm and walls(int8) (-1:nx+2,...); phi (-3:nx+4,...); w_in, phi_convec_bc
 *    (0:nx+1,0:ny+1); f_convec_bc,g_convec_bc (0:nx+1,0:ny+1,0:18).  The library copies, never
 *    retains host pointers, and owns all device memory.
 *  - one context per process per GPU, not thread safe (the reference issues all device work from one
 *    host thread per rank, multiphase_3D/makefile:48).
 *  - there is NO CPU fallback: every entry point fails with MFLBM_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef MFLBM_H
#define MFLBM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFLBM_OK 0
#define MFLBM_ERR_ARG (-1)
#define MFLBM_ERR_CUDA (-2)
#define MFLBM_ERR_NCCL (-3)
#define MFLBM_ERR_STATE (-4)

#define MFLBM_SOLVER_SINGLEPHASE 0
#define MFLBM_SOLVER_MULTIPHASE 1

/* type(indirect_solid_boundary_nodes), MP/Module.F90:90-94 (gfortran/x86-64 layout, 96 bytes) */
typedef struct {
    int32_t ix, iy, iz, i_fluid_num;
    int32_t neighbor_list[18];
    double la_weight;
} mflbm_solid_node;

/* type(indirect_fluid_boundary_nodes), MP/Module.F90:98-101 (48 bytes) */
typedef struct {
    int32_t ix, iy, iz, pad_;
    double nwx, nwy, nwz, theta;
} mflbm_fluid_node;

/* Everything the kernels read from the reference's module globals (MP/Module.F90). */
typedef struct {
    int32_t struct_size;            /* = sizeof(mflbm_config), ABI guard */
    int32_t solver;                 /* MFLBM_SOLVER_* */
    int32_t nx, ny, nz;             /* local slab (MP/Module.F90:80) */
    int32_t nxGlobal, nyGlobal, nzGlobal;
    int32_t idz, npz;               /* slab position; x,y undivided (MP/IO_multiphase.F90:495-498) */
    int32_t jper, kper;             /* periodic indicators (x periodic is rejected by the reference).  jper = 1 (y is never
                                       decomposed here: the lattice exchanges with itself in y, MP/Mpi.F90:22-40, :147-207;
                                       z slabs are fine) always runs the sparse population layout */
    int32_t domain_wall_status_z_min, domain_wall_status_z_max;
    int32_t inlet_BC, outlet_BC;    /* 1 velocity / convective, 2 Zou-He pressure */
    int32_t porous_plate_cmd, Z_porous_plate;
    int32_t mrt;                    /* MP/preprocessor.h: 1..4, shipped value 2 */
    int32_t iz_async;               /* MPI_async_layers_num z (>=4 for multiphase) */
    int32_t num_solid_boundary, num_fluid_boundary;
    int32_t device;                 /* CUDA ordinal; <0 = keep current (replaces setDevice, MP/Misc.F90:437) */
    int32_t use_nccl;               /* 1: z-halo exchange between ranks with ncclSend/ncclRecv */
    int32_t kernel_variant;         /* population layout: 0 = auto (sparse active-node list when porosity <= 0.8, else dense),
                                       1 = dense (the reference's direct addressing), 2 = sparse.  porous_plate_cmd != 0
                                       always runs dense (the plate copies from arbitrary nodes), jper = 1 always sparse
                                       (the y wrap is part of its adjacency).  Results are identical. */
    int32_t reserved_i[7];
    double la_nui1, la_nui2;        /* 1/nu1, 1/nu2 (MP/Init_multiphase.F90:120-121) */
    double gamma, beta, force_Z, phi_inlet, sa_inject, relaxation, uin_avg, rho_in, rho_out;
    double s_e, s_e2, s_q, s_nu, s_pi, s_t; /* singlephase constant rates (SP/Initialization.F90:87-112) */
    double reserved_d[8];
    unsigned char nccl_unique_id[128];      /* from mflbm_nccl_unique_id on rank 0, broadcast by the caller */
} mflbm_config;

typedef struct mflbm_ctx mflbm_ctx;

/* Host-array bundle for upload/download; NULL members are skipped. */
typedef struct {
    double *f[19];           /* f0..f18 */
    double *g[19];           /* g0..g18 (multiphase) */
    double *phi;             /* (-3:n+4)^3 */
    double *phi_old;
    double *cn_x, *cn_y, *cn_z, *c_norm; /* (-1:n+2)^3 ; download only */
    double *curv;            /* (0:n+1)^3 ; download only */
    double *u, *v, *w, *rho; /* (0:n+1)^3 ; download only (valid after mflbm_compute_macro_vars) */
    int8_t *walls;           /* (-1:n+2)^3 ; upload only */
    double *w_in;            /* (0:nx+1,0:ny+1) ; upload only */
    double *f_convec_bc, *g_convec_bc, *phi_convec_bc;
    const mflbm_solid_node *solid_boundary_nodes; /* upload only, cfg.num_solid_boundary entries */
    const mflbm_fluid_node *fluid_boundary_nodes; /* upload only, cfg.num_fluid_boundary entries */
} mflbm_arrays;

/* "!$acc data" entry + setDevice: MP/Main_multiphase.F90:70,104-115 ; SP/Main.F90 */
int mflbm_create(const mflbm_config *cfg, mflbm_ctx **out);
/* "!$acc end data": MP/Main_multiphase.F90:325 */
void mflbm_destroy(mflbm_ctx *ctx);
const char *mflbm_last_error(const mflbm_ctx *ctx);
const char *mflbm_version(void);

/* copy/copyin clauses of the data region (MP/Main_multiphase.F90:104-108); also used after
 * initialization_old_multi (checkpoint restart, MP/Init_multiphase.F90:477-557) */
int mflbm_upload(mflbm_ctx *ctx, const mflbm_arrays *host);
/* "!$acc update host(...)" in save_checkpoint / save_phi / save_macro / VTK_* (MP/IO_multiphase.F90:572,654,686,857) */
int mflbm_download(mflbm_ctx *ctx, const mflbm_arrays *host);

/* call main_iteration_kernel (MP/Main_multiphase.F90:167,341-486 ; SP/Main.F90:291-422):
 * collision+streaming for this ntime parity, halo exchange, inlet/outlet/porous-plate BCs and
 * (multiphase) color_gradient.  Asynchronous on the library's streams. */
int mflbm_step(mflbm_ctx *ctx, int ntime);
/* nsteps consecutive calls of mflbm_step starting at ntime0 (benchmark loop, MP/Main_multiphase.F90:515-531) */
int mflbm_run(mflbm_ctx *ctx, int ntime0, int nsteps);
/* Streamed steps: main_iteration_kernel for a driver that hands over a HOST input and wants a result back every time step
 * (a time-dependent inlet profile, MP/Init_multiphase.F90 inlet_vel_profile_rectangular; cal_saturation as a per-step
 * monitor, MP/Monitor.F90:472-507) without stalling the device twice per step the way mflbm_upload + mflbm_step +
 * mflbm_cal_saturation do.  w_in_host (pinned, the reference's w_in(0:nx+1,0:ny+1); NULL = keep the current profile) is
 * copied on the copy stream while the previous step still runs; the saturation sums of step ntime are read back
 * asynchronously and handed out by the NEXT call: *v1, *v2 = sums after the previous streamed step, *have_prev = 0 on the
 * first call (singlephase: no result, *have_prev stays 0; the call then only paces the host one step behind the device).
 * mflbm_stream_flush waits for the last streamed step and returns its sums.  Results and populations are those of
 * mflbm_step / mflbm_cal_saturation. */
int mflbm_step_streamed(mflbm_ctx *ctx, int ntime, const double *w_in_host, double *v1, double *v2, int *have_prev);
int mflbm_stream_flush(mflbm_ctx *ctx, double *v1, double *v2);
/* call color_gradient (MP/Main_multiphase.F90:120 ; MP/Phase_gradient.F90:5-204) */
int mflbm_color_gradient(mflbm_ctx *ctx);
/* call compute_macro_vars (MP/Misc.F90:372-430 ; SP/Misc.F90:368-423) */
int mflbm_compute_macro_vars(mflbm_ctx *ctx);

/* device part of monitor (MP/Monitor.F90:27-107 ; SP/Monitor.F90:18-64): fills the reference's tk buffer.
 * multiphase: tk[0:nz)=fl1, [nz:2nz)=fl2, vol1, vol2, mass1, mass2, pre, then umax, usq1, usq2 (7*nz+3)
 * singlephase: tk[0:nz)=fl, [nz:2nz)=pre, then umax (2*nz+1).  Calls compute_macro_vars first. */
int mflbm_monitor(mflbm_ctx *ctx, double *tk, int tk_len);
/* cal_saturation partial sums v1,v2 of this slab (MP/Monitor.F90:527-538) */
int mflbm_cal_saturation(mflbm_ctx *ctx, double *v1, double *v2);
/* monitor_breakthrough count on plane nz-1 of the last slab (MP/Monitor.F90:483-495); 0 on other slabs */
int mflbm_monitor_breakthrough(mflbm_ctx *ctx, int32_t *outlet_phase1_count);
/* monitor_multiphase_steady_phasefield device part (MP/Monitor.F90:303-334): max u^2, max |phi-phi_old|, phi_old<-phi */
int mflbm_monitor_steady_phasefield(mflbm_ctx *ctx, double *umax_sq, double *d_phi_max);
/* monitor_multiphase_steady_capillarypressure device part (MP/Monitor.F90:383-423) */
int mflbm_monitor_steady_capillarypressure(mflbm_ctx *ctx, double *umax_sq, double *pre_w, double *pre_nw,
                                           int32_t *i_w, int32_t *i_nw);

/* run-time changes of module scalars the driver makes between steps (force_z increments
 * MP/Main_multiphase.F90, rho_in); name is the reference variable name */
int mflbm_set_parameter(mflbm_ctx *ctx, const char *name, double value);

/* "!$acc wait" / MPI_Barrier + system_clock of benchmark (MP/Main_multiphase.F90:524-538) */
int mflbm_sync(mflbm_ctx *ctx);
int mflbm_timer_start(mflbm_ctx *ctx);
int mflbm_timer_stop(mflbm_ctx *ctx, double *elapsed_ms); /* CUDA-event time on the compute stream */

/* Per-kernel device timing of the dominant kernel (collision + streaming) for the roofline report: while enabled,
 * every launch of it is bracketed by CUDA events on its own stream.  mflbm_profile_read synchronises, returns the
 * accumulated kernel time and launch count since the last read, and resets the counters.  The analogue of the
 * reference's -Dgpu_profiling cudaProfilerStart/Stop bracket (MP/Main_multiphase.F90:525-533). */
int mflbm_profile(mflbm_ctx *ctx, int enable);
int mflbm_profile_read(mflbm_ctx *ctx, double *collide_ms, long long *collide_launches);

/* Quiet-tile statistics of the sparse multiphase layout (DESIGN.md "Quiet tiles"): number of 8x4x4-cell tiles and how
 * many of them the last gradient chain skipped because phi is uniform around them.  Informational (bench.py reports
 * the fraction; tests use it to make sure the skipping path is exercised); 0/0 when the layout has no tiles. */
int mflbm_tile_stats(mflbm_ctx *ctx, long long *ntiles, long long *nquiet);
/* Which kernels evaluate color_gradient (MP/Phase_gradient.F90:5-265) on this context: *fused = 1 when the single fused
 * kernel of the sparse multiphase layout is in use (csrc/march.cuh), 0 for the five reference-order kernels.  The fused
 * kernel needs node lists as the reference's geometry_preprocessing_new makes them (every node listed once, on a cell of
 * its kind, la_weight = sum of the listed neighbours' weights); *reject_mask says what was found otherwise
 * (1 solid entry misplaced / duplicated, 2 foreign la_weight, 4 fluid entry misplaced / duplicated). */
int mflbm_chain_info(mflbm_ctx *ctx, int *fused, int *reject_mask);
/* Developer check: re-evaluates color_gradient on the current phase field with the five reference-order kernels and counts
 * the entries of the packed colour gradient the collision kernel reads that differ bit for bit from what the last
 * evaluation left there (0 expected; only meaningful right after mflbm_step / mflbm_color_gradient). */
int mflbm_chain_selfcheck(mflbm_ctx *ctx, long long *mismatches);

/* kernel launches issued by this context since create (bench.py "gpu_launches") */
long long mflbm_launch_count(const mflbm_ctx *ctx);
/* algorithmic device bytes held by the context */
long long mflbm_device_bytes(const mflbm_ctx *ctx);

/* NCCL bootstrap: rank 0 calls this and broadcasts the 128 bytes (MPI_Bcast in the Fortran driver,
 * torch.distributed in bench.py); replaces MPI_CART_CREATE for the z ring (MP/Mpi_misc.F90:19-38) */
int mflbm_nccl_unique_id(unsigned char id[128]);

/* ---- asynchronous output staging ----------------------------------------------------------
 * save_phi / save_macro / VTK_* of the reference (MP/IO_multiphase.F90:646-712, :857-890) do `!$acc update host(...)`
 * and write files while the device waits.  mflbm_output_begin snapshots the requested fields at the current step into
 * packed device buffers (for MFLBM_OUT_MACRO after running compute_macro_vars, like save_macro does, MP/Misc.F90:372-430)
 * and starts the device-to-host copy into pinned memory on a copy stream; it returns at once and the step loop goes
 * on.  mflbm_output_end waits for that copy and fills the caller's arrays (host->phi, host->u, v, w, rho in the
 * reference's extents); the Fortran writers then format exactly the arrays they format today.  One output may be in
 * flight per context. */
#define MFLBM_OUT_PHI 1
#define MFLBM_OUT_MACRO 2
int mflbm_output_begin(mflbm_ctx *ctx, int what);
int mflbm_output_end(mflbm_ctx *ctx, const mflbm_arrays *host);

/* ---- checkpoint staging (second half) -------------------------------------------------------
 * save_checkpoint (MP/IO_multiphase.F90:562-642) does `!$acc update host(f0..f18, g0..g18, phi [, *_convec_bc])` and
 * writes the per-rank stream file while the device waits; initialization_old_multi (MP/Init_multiphase.F90:477-557)
 * reads the same arrays back (restart = mflbm_upload of them + mflbm_color_gradient, MP/Main_multiphase.F90:120).
 *   mflbm_checkpoint_begin  freezes the state of the current step: when free device memory allows, a device-side snapshot
 *                           of the populations (in the device layout), phi and the convective-outlet state is taken with
 *                           stream-ordered copies and the step loop may go on at once (returns MFLBM_CKPT_STAGED = 0);
 *                           otherwise nothing is copied and the context is frozen -- mflbm_step / mflbm_run are refused
 *                           until mflbm_checkpoint_end (returns MFLBM_CKPT_DIRECT = 1).
 *   mflbm_checkpoint_fetch  fills whichever of host->f[q], g[q], phi, f_convec_bc, g_convec_bc, phi_convec_bc are non-null
 *                           with the frozen state, in the reference's extents (ghost layers included), on a copy stream
 *                           that overlaps steps already queued; may be called repeatedly (one array at a time keeps the
 *                           host footprint at one array, the way save_checkpoint writes them).  Entries of a population
 *                           array that are dead storage in the sparse layout keep the caller's values, like mflbm_download.
 *   mflbm_checkpoint_end    releases the snapshot / unfreezes the context. */
#define MFLBM_CKPT_STAGED 0
#define MFLBM_CKPT_DIRECT 1
int mflbm_checkpoint_begin(mflbm_ctx *ctx);
int mflbm_checkpoint_fetch(mflbm_ctx *ctx, const mflbm_arrays *host);
int mflbm_checkpoint_end(mflbm_ctx *ctx);

/* ---- geometry preprocessing on the device -------------------------------------------------
 * Replaces geometry_preprocessing_new (MP/Geometry_preprocessing.F90:9-512), called from set_walls / the main
 * program before initialization (MP/Main_multiphase.F90:98): classification of the wall array into solid / fluid
 * boundary nodes, the two node lists of the colour-gradient chain (types above) with neighbour lists, la_weight and the
 * ISO8 wall normals of the four-times smoothed wall field.  The reference does this serially on rank 0 over the whole
 * lattice and broadcasts global lists (:4-8 flags it as too slow for large domains); here every rank hands in the z
 * window of the global wall array around its slab and gets its LOCAL lists (:424-507) back, in the reference's order
 * (k outer, i inner) with local iz.  No context is needed; lists are malloc'ed by the library. */
typedef struct {
    int32_t struct_size;              /* sizeof(mflbm_geometry_config) */
    int32_t nxGlobal, nyGlobal, nzGlobal;
    int32_t wk0, wk1;                 /* global planes (1-based, inclusive) held in walls_window(1:nxG,1:nyG,wk0:wk1); the window
                                         must reach the lattice end or extend >= 10 planes beyond the slab on that side */
    int32_t idz, npz;                 /* slab of this rank: planes idz*nz+1 .. idz*nz+nz, nz = nzGlobal/npz */
    int32_t iper, jper, kper;         /* periodic_indicator (wrap instead of replicate in the ghost layers, :56-108) */
    int32_t device;                   /* CUDA device ordinal, -1 = current */
    double theta;                     /* contact angle stored in every fluid node, radians, already converted
                                         (pi - theta, MP/IO_multiphase.F90:467-468) */
} mflbm_geometry_config;

int mflbm_geometry_preprocess(const mflbm_geometry_config *cfg, const int8_t *walls_window, mflbm_solid_node **solid,
                              int32_t *num_solid, mflbm_fluid_node **fluid, int32_t *num_fluid,
                              int64_t *num_solid_scanned, int64_t *num_fluid_scanned);
void mflbm_geometry_free(void *list);
const char *mflbm_geometry_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
