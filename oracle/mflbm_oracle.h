/*
 * mflbm_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; NOT PRODUCT CODE)
 *
 * Plain-C restatement of the lanl/MF-LBM time-step hot path (multiphase_3D and
 * singlephase_3D).  Every function cites the reference file:line it follows
 * ("MP/" = multiphase_3D/0.src/, "SP/" = singlephase_3D/0.src/).
 *
 * PARITY UNPINNED: the reference ships no golden vectors / known-answer tests
 * for this path and cannot be compiled in this image (no
 * Fortran compiler).  The oracle is pinned only by (1) the integer known
 * answers of the shipped geometry fixtures, (2) analytic fixed points the
 * reference encodes itself, see tests/test_oracle.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference arm may load this library.
 *
 * Array layout is the reference's (column-major, i fastest, ghost layers as in
 * MP/Init_multiphase.F90:594-658):
 *   PDFs f0..f18,g0..g18,u,v,w,rho,curv : (0:nx+1, 0:ny+1, 0:nz+1)
 *   cn_x,cn_y,cn_z,c_norm, walls(int8)   : (-1:nx+2, -1:ny+2, -1:nz+2)
 *   phi, phi_old                         : (-3:nx+4, -3:ny+4, -3:nz+4)
 *   w_in, phi_convec_bc                  : (0:nx+1, 0:ny+1)
 *   f_convec_bc,g_convec_bc              : (0:nx+1, 0:ny+1, 0:18)
 */
#ifndef MFLBM_ORACLE_H
#define MFLBM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* MP/Module.F90:90-94 */
typedef struct {
    int32_t ix, iy, iz, i_fluid_num;
    int32_t neighbor_list[18];
    double la_weight;
} orc_solid_node; /* 96 bytes */

/* MP/Module.F90:98-101 */
typedef struct {
    int32_t ix, iy, iz, pad_;
    double nwx, nwy, nwz, theta;
} orc_fluid_node; /* 48 bytes */

typedef struct {
    /* ---- integers ---- */
    int32_t multiphase;              /* 1 = multiphase_3D, 0 = singlephase_3D */
    int32_t nxG, nyG, nzG;           /* global lattice */
    int32_t idz, npz;                /* z-slab position (x,y are never decomposed here) */
    int32_t iper, jper, kper;        /* periodic indicators */
    int32_t wsx0, wsx1, wsy0, wsy1, wsz0, wsz1; /* domain_wall_status_{x,y,z}_{min,max} */
    int32_t n_exclude_inlet, n_exclude_outlet;
    int32_t inlet_BC, outlet_BC;     /* 1 velocity/convective, 2 pressure */
    int32_t porous_plate_cmd, Z_porous_plate;
    int32_t mrt;                     /* MP/preprocessor.h: 1..4 (shipped 2) */
    int32_t initial_fluid_distribution_option;
    int32_t modify_geometry_cmd;
    int32_t mrt_para_preset;         /* singlephase */
    int32_t steady_state_option;
    int32_t reserved_i[3];
    /* ---- doubles ---- */
    double la_nu1, la_nu2;           /* viscosities (singlephase: la_nu1 = la_nu) */
    double gamma, beta;
    double theta_deg;                /* contact angle as written in the control file */
    double force_z0;
    double sa_inject;
    double ca_0;
    double interface_z0;
    double rho_drop;                 /* singlephase pressure inlet */
    double Re, char_length;          /* singlephase velocity inlet */
    double target_inject_pore_volume;
    double reserved_d[4];
} orc_params;

typedef struct orc_state orc_state;

orc_state *orc_create(const orc_params *p);
void orc_destroy(orc_state *s);

/* geometry */
int8_t *orc_walls_global(orc_state *s);                 /* (1:nxG,1:nyG,1:nzG) */
void orc_set_walls(orc_state *s);                       /* MP/Misc.F90:6-210 (after walls_global is filled) */
void orc_geometry_preprocess(orc_state *s);             /* MP/Geometry_preprocessing.F90:9-512 */
/* init */
void orc_init_basic(orc_state *s);                      /* MP/Init_multiphase.F90:68-150 */
void orc_init_phi(orc_state *s);                        /* MP/Init_multiphase.F90:276-337 */
void orc_init_pdf(orc_state *s);                        /* MP/Init_multiphase.F90:254-273,357-470 */
/* hot path */
void orc_kernel_odd(orc_state *s, int ix0, int ix1, int iy0, int iy1, int iz0, int iz1);
void orc_kernel_even(orc_state *s, int ix0, int ix1, int iy0, int iy1, int iz0, int iz1);
void orc_color_gradient(orc_state *s);                  /* MP/Phase_gradient.F90:5-265 */
void orc_step(orc_state *s, int ntime);                 /* MP/Main_multiphase.F90:341-486 (np=1) */
void orc_compute_macro_vars(orc_state *s);              /* MP/Misc.F90:372-430 */
/* monitors; out arrays sized nz (local) */
typedef struct {
    double umax, usq1, usq2;
    double saturation, saturation_full_domain, vol1_sum, vol2_sum, mass1_sum, mass2_sum;
    double fl_avg_whole, fl1_avg_whole, fl2_avg_whole, fl_avg, fl1_avg, fl2_avg;
    double ca, umax_global, pre_in, pre_out;
    double d_phi_max;
    double pre_w, pre_nw;
    int32_t i_w, i_nw, outlet_phase1_sum, pad_;
} orc_monitor_out;
void orc_monitor(orc_state *s, orc_monitor_out *o);     /* MP/Monitor.F90:5-277 ; SP/Monitor.F90:4-172 */
void orc_cal_saturation(orc_state *s, orc_monitor_out *o);          /* MP/Monitor.F90:512-550 */
void orc_monitor_breakthrough(orc_state *s, orc_monitor_out *o);    /* MP/Monitor.F90:472-507 */
void orc_monitor_steady_phasefield(orc_state *s, orc_monitor_out *o);      /* MP/Monitor.F90:287-365 */
void orc_monitor_steady_capillarypressure(orc_state *s, orc_monitor_out *o); /* MP/Monitor.F90:370-462 */

/* accessors (pointers into the oracle state; valid until orc_destroy) */
double *orc_f(orc_state *s, int q);
double *orc_g(orc_state *s, int q);
double *orc_field(orc_state *s, const char *name); /* phi phi_old cn_x cn_y cn_z c_norm curv u v w rho w_in
                                                      f_convec_bc g_convec_bc phi_convec_bc
                                                      fl1 fl2 vol1 vol2 mass1 mass2 pre */
int8_t *orc_walls(orc_state *s);
orc_solid_node *orc_solid_nodes(orc_state *s, int *n);
orc_fluid_node *orc_fluid_nodes(orc_state *s, int *n);
int orc_get_int(orc_state *s, const char *name);      /* nx ny nz pore_sum pore_sum_effective num_solid_boundary_global ... */
long long orc_get_i64(orc_state *s, const char *name);
double orc_get_double(orc_state *s, const char *name);
void orc_set_double(orc_state *s, const char *name, double v);
void orc_set_int(orc_state *s, const char *name, int v);

#ifdef __cplusplus
}
#endif
#endif
