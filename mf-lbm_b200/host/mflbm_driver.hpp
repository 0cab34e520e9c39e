// mflbm_driver.hpp -- host-side mirror of the MF-LBM Fortran driver for the time-step hot path.
//
// The reference keeps all state in Fortran modules and drives the kernels through argument-less subroutines
// (SURVEY 8b).  This image has no Fortran compiler, so the host side above the C ABI (include/mflbm.h) is
// written in C++ with the reference's routine names, argument meaning and error behaviour:
//   read_parameter_multi / read_parameter   MP/IO_multiphase.F90:5-556, SP/IO.F90
//   set_walls, modify_geometry, read_walls, pore_profile      MP/Misc.F90:6-365
//   geometry_preprocessing_new                                MP/Geometry_preprocessing.F90:9-512
//   initialization_basic(_multi), initialization_open_velocity_inlet_BC, inlet_vel_profile_rectangular,
//   initialization_new(_multi)(_pdf)                          MP/Init_multiphase.F90:5-470, SP/Initialization.F90
//   main_iteration_kernel, color_gradient, monitor*, cal_saturation, benchmark
//                                                             -> forwarded to libmflbm.so (GPU), never computed here
// Init-time routines run on the host cores exactly like the reference (they are outside the hot path and
// stay with the driver per BASELINE.json north_star); everything per-step goes through the C ABI.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/mflbm.h"

namespace mflbm_host {

// control-file parameters (names as in template-simulation_control.txt)
struct Control {
    int multiphase = 1;
    int initial_fluid_distribution_option = 1, benchmark_cmd = 0, extreme_large_sim_cmd = 0, breakthrough_check = 0;
    int steady_state_option = 0, output_fieldData_precision_cmd = 0, modify_geometry_cmd = 0, external_geometry_read_cmd = 0;
    int geometry_preprocess_cmd = 0, porous_plate_cmd = 0, Z_porous_plate = 0, change_inlet_fluid_phase_cmd = 0;
    int nxGlobal = 40, nyGlobal = 40, nzGlobal = 60;
    int n_exclude_inlet = 10, n_exclude_outlet = 10;
    int domain_wall_status_x_min = 1, domain_wall_status_x_max = 1, domain_wall_status_y_min = 1, domain_wall_status_y_max = 1;
    int domain_wall_status_z_min = 0, domain_wall_status_z_max = 0;
    int iper = 0, jper = 0, kper = 0;
    int npx = 1, npy = 1, npz = 1;
    int ix_async = 0, iy_async = 4, iz_async = 4;
    double la_nu1 = 0.004, la_nu2 = 0.4, gamma = 0.03, theta_deg = 30.0, beta = 0.95;
    int inlet_BC = 1, outlet_BC = 1;
    double sa_inject = 1.0, target_inject_pore_volume = 1.0, interface_z0 = 8.0, ca_0 = 100e-6, force_z0 = 0.0, sa_target = 0.4;
    long long ntime_max = 100000000;
    int ntime_max_benchmark = 100, ntime_visual = 10000000, ntime_animation = 2000, ntime_monitor = 1000;
    int ntime_monitor_profile_ratio = 5, ntime_clock_sum = 1000, ntime_display_steps = 1000;
    double convergence_criteria = 1e-6;
    double checkpoint_save_timer = 2.0, checkpoint_2rd_save_timer = 5.5, simulation_duration_timer = 15.7;
    double d_vol_animation = 0.05, d_vol_detail = -1.0, d_vol_monitor = 0.01;
    int mrt = 2;  // MP/preprocessor.h
    // singlephase-only keys (SP/IO.F90)
    int mrt_para_preset = 1;
    double char_length = 1.0, Re = 1.0, rho_drop = 0.0;
};

struct MonitorResult {
    double saturation = 0, saturation_full_domain = 0, vol1_sum = 0, vol2_sum = 0, mass1_sum = 0, mass2_sum = 0;
    double fl_avg_whole = 0, fl1_avg_whole = 0, fl2_avg_whole = 0, fl_avg = 0, fl1_avg = 0, fl2_avg = 0;
    double ca = 0, umax_global = 0, kinetic_energy1 = 0, kinetic_energy2 = 0, pre_in = 0, pre_out = 0, flowrate = 0;
    int simulation_end_indicator = 0;
};

// One z-slab (idz of npz) of the driver state.  Arrays use the reference's extents, i fastest.
class Driver {
public:
    Control c;
    int idz = 0;
    int nx = 0, ny = 0, nz = 0;
    // derived scalars (MP/Init_multiphase.F90:68-150)
    double la_x = 0, la_y = 0, la_z = 0, A_xy = 0, A_xy_effective = 0, la_nui1 = 0, la_nui2 = 0, theta = 0, phi_inlet = 0;
    double force_Z = 0, rho_in = 1, rho_out = 1, relaxation = 1, uin_avg = 0, uin_avg_0 = 0, flowrate = 0;
    double s_e = 0, s_e2 = 0, s_q = 0, s_nu = 0, s_pi = 0, s_t = 0;
    long long pore_sum = 0, pore_sum_effective = 0;
    std::vector<int> pore_profile_z;  // profile over the planes held in walls_global (index k - wk0)
    // geometry
    std::vector<int8_t> walls_global;  // (1:nxG,1:nyG,wk0:wk1): the whole lattice, or a z window around this slab
    int wk0 = 1, wk1 = 0;              // global plane range held in walls_global (wk1 < wk0: not set -> whole lattice)
    long long pore_sum_local = 0;      // fluid nodes of this slab
    std::vector<int8_t> walls;         // (-1:n+2)^3
    std::vector<mflbm_solid_node> solid_boundary_nodes;
    std::vector<mflbm_fluid_node> fluid_boundary_nodes;
    long long num_solid_boundary_global = 0, num_fluid_boundary_global = 0;
    // fields handed to mflbm_upload
    std::vector<double> f[19], g[19], phi, w_in, f_convec_bc, g_convec_bc, phi_convec_bc;
    // device context
    bool device_geometry = false;  // true: geometry_preprocessing_new runs on the GPU (mflbm_geometry_preprocess)
    int geometry_device = -1;
    bool lazy_pdfs = false;  // true: populations are generated one array at a time during upload (large lattices)
    mflbm_ctx *ctx = nullptr;
    std::string error;

    ~Driver();
    // ---- control file ----
    bool read_parameter(const std::string &path);  // read_parameter_multi / read_parameter + consistency checks
    bool check_parameters();                        // MP/IO_multiphase.F90:455-552 (returns false = mpi_abort in the reference)
    // ---- geometry ----
    bool read_walls(const std::string &path);
    // only the planes around this rank's slab, straight from the file (no whole-lattice broadcast)
    bool read_walls_window(const std::string &path, int margin);       // MP/Misc.F90:247-295
    void modify_geometry();                         // MP/Misc.F90:213-244
    void set_walls();                               // MP/Misc.F90:6-210 (after walls_global is filled) + pore_profile
    void geometry_preprocessing_new();
    bool geometry_preprocessing_device();           // same lists from the device kernels (csrc/kernels_geometry.cu)              // MP/Geometry_preprocessing.F90:9-512
    // ---- initialisation ----
    void initialization_basic();                    // initialization_basic_multi / initialization_basic (scalars, w_in)
    void initialization_new();                      // initialization_new_multi(_pdf) / initialization_new(_pdf)
    // ---- device ----
    bool create_context(int device, int use_nccl, const unsigned char *nccl_id, int kernel_variant);
    bool upload(bool with_geometry = true);
    void random_phase_field(unsigned long long seed);      // option 6 with a seeded, decomposition-independent draw
    bool reinitialize(int option, unsigned long long seed);  // new initial fluid distribution on the same geometry
    bool save_checkpoint(const std::string &path, int ntime);        // MP/IO_multiphase.F90:562-642 (staged, array by array)
    bool initialization_old(const std::string &path, int *ntime0);   // MP/Init_multiphase.F90:477-557
    bool main_iteration_kernel(int ntime);          // MP/Main_multiphase.F90:341-486 -> mflbm_step
    bool color_gradient();                          // MP/Phase_gradient.F90:5 -> mflbm_color_gradient
    bool monitor(int ntime, MonitorResult *out, const std::string &outdir);  // MP/Monitor.F90:5-277 (np==1 tail)
    bool cal_saturation(double *saturation_full_domain);                     // MP/Monitor.F90:512-550
    bool benchmark(int warmup, int rounds, int steps, double *best_mlups, double *ms_per_step);  // MP/Main_multiphase.F90:498-556

    double pdf_value(int i, int j, int k, int q, int fluid) const;
    void fill_pdf(std::vector<double> &a, int q, int fluid) const;

private:
    void inlet_vel_profile_rectangular(double vel_avg, int num_terms);  // MP/Misc.F90:625-665
};

}  // namespace mflbm_host
