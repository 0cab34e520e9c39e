// mflbm_driver_capi.cpp -- flat C interface of the host driver mirror for ctypes (tests, bench.py).
#include <cstring>
#include <string>

#include "mflbm_driver.hpp"

using mflbm_host::Driver;
using mflbm_host::MonitorResult;

extern "C" {

void *mfd_create(void) { return new Driver(); }
void mfd_destroy(void *h) { delete (Driver *)h; }
const char *mfd_error(void *h) { return ((Driver *)h)->error.c_str(); }

int mfd_read_parameter(void *h, const char *path) { return ((Driver *)h)->read_parameter(path) ? 0 : -1; }
int mfd_read_walls(void *h, const char *path) { return ((Driver *)h)->read_walls(path) ? 0 : -1; }
int mfd_read_walls_window(void *h, const char *path, int margin) { return ((Driver *)h)->read_walls_window(path, margin) ? 0 : -1; }
void mfd_set_idz(void *h, int idz) { ((Driver *)h)->idz = idz; }
void mfd_set_lazy_pdfs(void *h, int on) { ((Driver *)h)->lazy_pdfs = on != 0; }
void mfd_set_device_geometry(void *h, int on, int device) {
    ((Driver *)h)->device_geometry = on != 0;
    ((Driver *)h)->geometry_device = device;
}

// walls_global handed over in memory ((1:nx,1:ny,1:nz) int8, i fastest); zero-padded to the lattice like read_walls
int mfd_set_walls_global(void *h, const int8_t *w, int nxs, int nys, int nzs) {
    Driver *d = (Driver *)h;
    const int nxG = d->c.nxGlobal, nyG = d->c.nyGlobal, nzG = d->c.nzGlobal;
    if (nxs > nxG || nys > nyG || nzs > nzG) {
        d->error = "Error! Domain size is smaller than porous media sample size! Exiting program!";
        return -1;
    }
    d->walls_global.assign((size_t)nxG * nyG * nzG, 0);
    d->wk0 = 1;
    d->wk1 = nzG;
    for (int k = 0; k < nzs; k++)
        for (int j = 0; j < nys; j++)
            std::memcpy(&d->walls_global[(size_t)nxG * ((size_t)j + (size_t)nyG * k)], w + (size_t)nxs * ((size_t)j + (size_t)nys * k), nxs);
    return 0;
}

// z window of the global wall array: planes wk0..wk0+nplanes-1 ((1:nxG,1:nyG,nplanes) int8, i fastest).  For a slab
// idz of npz the window must reach >= 12 planes beyond the slab (clipped at the ends of a non-periodic lattice;
// wrapped by the caller on a periodic one).  Lets each rank hold and preprocess only its own part of the geometry.
int mfd_set_walls_window(void *h, const int8_t *w, int nplanes, int wk0) {
    Driver *d = (Driver *)h;
    {   // same extent rule as mflbm_geometry_preprocess: the window reaches the lattice end or >= 10 planes beyond the slab
        // (smoothing 4 + ISO8 2 + list ghosts 3, + the 2 ghost planes set_walls fills); a smaller one would be read out of bounds
        const int nzG = d->c.nzGlobal, nz = nzG / (d->c.npz > 0 ? d->c.npz : 1);
        const int ks0 = d->idz * nz + 1, ks1 = d->idz * nz + nz, wk1 = wk0 + nplanes - 1;
        const bool whole = wk0 <= 1 && wk1 >= nzG;  // the whole lattice: set_walls wraps it itself when z is periodic
        const bool lo_ok = whole || wk0 <= ks0 - 10 || (d->c.kper == 0 && wk0 <= 1), hi_ok = whole || wk1 >= ks1 + 10 || (d->c.kper == 0 && wk1 >= nzG);
        if (!lo_ok || !hi_ok) {
            d->error = "wall window too small: it must reach the lattice end or extend >= 10 planes beyond the slab";
            return -1;
        }
    }
    const size_t n = (size_t)d->c.nxGlobal * d->c.nyGlobal * nplanes;
    d->walls_global.assign(w, w + n);
    d->wk0 = wk0;
    d->wk1 = wk0 + nplanes - 1;
    return 0;
}
void mfd_set_pore_sum(void *h, long long pore_sum) { ((Driver *)h)->pore_sum = pore_sum; }

// initialization_basic_multi + initialization_new_multi (MP/Main_multiphase.F90:90-95)
int mfd_setup(void *h) {
    Driver *d = (Driver *)h;
    d->set_walls();
    if (d->c.multiphase) {
        if (d->device_geometry) {
            if (!d->geometry_preprocessing_device()) return -1;
        } else {
            d->geometry_preprocessing_new();
        }
    }
    d->initialization_basic();
    d->initialization_new();
    return 0;
}

int mfd_create_context(void *h, int device, int use_nccl, const unsigned char *nccl_id, int kernel_variant) {
    return ((Driver *)h)->create_context(device, use_nccl, nccl_id, kernel_variant) ? 0 : -1;
}
int mfd_upload(void *h) { return ((Driver *)h)->upload() ? 0 : -1; }
int mfd_save_checkpoint(void *h, const char *path, int ntime) { return ((Driver *)h)->save_checkpoint(path, ntime) ? 0 : -1; }
int mfd_initialization_old(void *h, const char *path, int *ntime0) { return ((Driver *)h)->initialization_old(path, ntime0) ? 0 : -1; }
int mfd_reinitialize(void *h, int option, unsigned long long seed) { return ((Driver *)h)->reinitialize(option, seed) ? 0 : -1; }
void *mfd_ctx(void *h) { return ((Driver *)h)->ctx; }
int mfd_main_iteration_kernel(void *h, int ntime) { return ((Driver *)h)->main_iteration_kernel(ntime) ? 0 : -1; }
int mfd_color_gradient(void *h) { return ((Driver *)h)->color_gradient() ? 0 : -1; }
int mfd_benchmark(void *h, int warmup, int rounds, int steps, double *mlups, double *ms_per_step) {
    return ((Driver *)h)->benchmark(warmup, rounds, steps, mlups, ms_per_step) ? 0 : -1;
}
int mfd_monitor(void *h, int ntime, MonitorResult *out, const char *outdir) {
    return ((Driver *)h)->monitor(ntime, out, outdir ? outdir : "") ? 0 : -1;
}
int mfd_cal_saturation(void *h, double *sat) { return ((Driver *)h)->cal_saturation(sat) ? 0 : -1; }
// release the host copies of the populations once they are on the device (large runs)
void mfd_free_host_fields(void *h) {
    Driver *d = (Driver *)h;
    for (int q = 0; q < 19; q++) {
        std::vector<double>().swap(d->f[q]);
        std::vector<double>().swap(d->g[q]);
    }
    std::vector<double>().swap(d->phi);
    std::vector<int8_t>().swap(d->walls_global);
}

long long mfd_get_i64(void *h, const char *name) {
    Driver *d = (Driver *)h;
    const std::string n(name);
#define G(x) if (n == #x) return (long long)d->x;
    G(nx) G(ny) G(nz) G(pore_sum) G(pore_sum_local) G(pore_sum_effective) G(num_solid_boundary_global) G(num_fluid_boundary_global) G(idz)
#undef G
#define G(x) if (n == #x) return (long long)d->c.x;
    G(multiphase) G(nxGlobal) G(nyGlobal) G(nzGlobal) G(npz) G(kper) G(jper) G(inlet_BC) G(outlet_BC) G(ntime_max) G(ntime_monitor)
    G(ntime_max_benchmark) G(benchmark_cmd) G(breakthrough_check) G(n_exclude_inlet) G(n_exclude_outlet) G(iz_async) G(mrt_para_preset)
    G(modify_geometry_cmd) G(external_geometry_read_cmd) G(initial_fluid_distribution_option)
#undef G
    if (n == "num_solid_boundary") return (long long)d->solid_boundary_nodes.size();
    if (n == "num_fluid_boundary") return (long long)d->fluid_boundary_nodes.size();
    return -1;
}
double mfd_get_double(void *h, const char *name) {
    Driver *d = (Driver *)h;
    const std::string n(name);
#define G(x) if (n == #x) return d->x;
    G(la_x) G(la_y) G(la_z) G(A_xy) G(A_xy_effective) G(la_nui1) G(la_nui2) G(theta) G(phi_inlet) G(force_Z) G(rho_in) G(rho_out)
    G(relaxation) G(uin_avg) G(uin_avg_0) G(flowrate) G(s_e) G(s_e2) G(s_q) G(s_nu) G(s_pi) G(s_t)
#undef G
#define G(x) if (n == #x) return d->c.x;
    G(la_nu1) G(la_nu2) G(gamma) G(beta) G(theta_deg) G(sa_inject) G(ca_0) G(force_z0) G(interface_z0) G(convergence_criteria)
#undef G
    return 0.0 / 0.0;
}
// raw views for the tests (valid until the next setup / destroy)
const int8_t *mfd_walls(void *h) { return ((Driver *)h)->walls.data(); }
const int8_t *mfd_walls_global(void *h) { return ((Driver *)h)->walls_global.data(); }
const mflbm_solid_node *mfd_solid_nodes(void *h) { return ((Driver *)h)->solid_boundary_nodes.data(); }
const mflbm_fluid_node *mfd_fluid_nodes(void *h) { return ((Driver *)h)->fluid_boundary_nodes.data(); }
double *mfd_field(void *h, const char *name, int q) {
    Driver *d = (Driver *)h;
    const std::string n(name);
    if (n == "f") return d->f[q].data();
    if (n == "g") return d->g[q].data();
    if (n == "phi") return d->phi.data();
    if (n == "w_in") return d->w_in.data();
    if (n == "f_convec_bc") return d->f_convec_bc.data();
    if (n == "g_convec_bc") return d->g_convec_bc.data();
    if (n == "phi_convec_bc") return d->phi_convec_bc.data();
    return nullptr;
}
// caller-provided phase field for initial_fluid_distribution_option 6 ((-3:n+4)^3)
int mfd_set_phi(void *h, const double *phi, long long n) {
    Driver *d = (Driver *)h;
    d->phi.assign(phi, phi + n);
    return 0;
}
}
