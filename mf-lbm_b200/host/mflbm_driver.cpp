// mflbm_driver.cpp -- see mflbm_driver.hpp.  Host-side (init-time) routines of the MF-LBM driver in C++17 +
// OpenMP; everything per time step is forwarded to the CUDA library through include/mflbm.h.
#include "mflbm_driver.hpp"

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace mflbm_host {

namespace {
constexpr int EX[19] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0};
constexpr int EY[19] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1};
constexpr int EZ[19] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1};
constexpr double W0 = 1.0 / 3.0, W1 = 1.0 / 18.0, W2 = 1.0 / 36.0;
constexpr double PI = 3.14159265358979323846;   // MP/Module.F90:7
constexpr double EPS_MP = 1.110223025e-16;      // MP/Module.F90:8
inline double wq(int q) { return q == 0 ? W0 : (q <= 6 ? W1 : W2); }

// views with Fortran index semantics
struct V1 {  // (0:nx+1,0:ny+1,0:nz+1)
    size_t sx, sy;
    size_t operator()(int i, int j, int k) const { return (size_t)i + sx * ((size_t)j + sy * (size_t)k); }
};
struct V2 {  // (-1:n+2)
    size_t sx, sy;
    size_t operator()(int i, int j, int k) const { return (size_t)(i + 1) + sx * ((size_t)(j + 1) + sy * (size_t)(k + 1)); }
};
struct V4 {  // (-3:n+4)
    size_t sx, sy;
    size_t operator()(int i, int j, int k) const { return (size_t)(i + 3) + sx * ((size_t)(j + 3) + sy * (size_t)(k + 3)); }
};

std::string trim(const std::string &s) {
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
double fnum(std::string t) {  // Fortran literals: 1d-6, 100d-6
    for (char &ch : t)
        if (ch == 'd' || ch == 'D') ch = 'e';
    return std::strtod(t.c_str(), nullptr);
}
std::vector<double> nums(const std::string &v) {
    std::vector<double> out;
    std::string t;
    std::stringstream ss(v);
    while (std::getline(ss, t, ',')) {
        t = trim(t);
        if (!t.empty()) out.push_back(fnum(t));
    }
    return out;
}
}  // namespace

Driver::~Driver() {
    if (ctx) mflbm_destroy(ctx);
}

// ---------------------------------------------------------------------------------------------------
// control file: "key value[,value...]" lines, '#' comments (MP/IO_multiphase.F90:66-277, SP/IO.F90)
// ---------------------------------------------------------------------------------------------------
bool Driver::read_parameter(const std::string &path) {
    std::ifstream in(path);
    if (!in) {
        error = "Error! simulation_control.txt is not found: " + path;
        return false;
    }
    std::string line;
    bool saw_fluid2 = false, saw_single = false;
    while (std::getline(in, line)) {
        const size_t h = line.find('#');
        if (h != std::string::npos) line = line.substr(0, h);
        line = trim(line);
        if (line.empty()) continue;
        const size_t sp = line.find_first_of(" \t");
        if (sp == std::string::npos) continue;
        const std::string key = line.substr(0, sp);
        const std::vector<double> v = nums(line.substr(sp + 1));
        if (v.empty()) continue;
        auto I = [&](size_t n) { return n < v.size() ? (int)std::llround(v[n]) : 0; };
#define KI(name, field) else if (key == name) c.field = I(0)
#define KD(name, field) else if (key == name) c.field = v[0]
        if (key == "lattice_dimensions") { c.nxGlobal = I(0); c.nyGlobal = I(1); c.nzGlobal = I(2); }
        else if (key == "excluded_layers") { c.n_exclude_inlet = I(0); c.n_exclude_outlet = I(1); }
        else if (key == "domain_wall_status_x") { c.domain_wall_status_x_min = I(0); c.domain_wall_status_x_max = I(1); }
        else if (key == "domain_wall_status_y") { c.domain_wall_status_y_min = I(0); c.domain_wall_status_y_max = I(1); }
        else if (key == "domain_wall_status_z") { c.domain_wall_status_z_min = I(0); c.domain_wall_status_z_max = I(1); }
        else if (key == "periodic_indicator") { c.iper = I(0); c.jper = I(1); c.kper = I(2); }
        else if (key == "MPI_process_num") { c.npx = I(0); c.npy = I(1); c.npz = I(2); }
        else if (key == "MPI_async_layers_num") { c.ix_async = I(0); c.iy_async = I(1); c.iz_async = I(2); }
        KI("initial_fluid_distribution_option", initial_fluid_distribution_option);
        KI("benchmark_cmd", benchmark_cmd);
        KI("extreme_large_sim_cmd", extreme_large_sim_cmd);
        KI("breakthrough_check", breakthrough_check);
        KI("steady_state_option", steady_state_option);
        KD("convergence_criteria", convergence_criteria);
        KI("output_fieldData_precision_cmd", output_fieldData_precision_cmd);
        KI("modify_geometry_cmd", modify_geometry_cmd);
        KI("external_geometry_read_cmd", external_geometry_read_cmd);
        KI("geometry_preprocess_cmd", geometry_preprocess_cmd);
        KI("porous_plate_cmd", porous_plate_cmd);
        KI("Z_porous_plate", Z_porous_plate);
        KI("change_inlet_fluid_phase_cmd", change_inlet_fluid_phase_cmd);
        else if (key == "fluid1_viscosity") { c.la_nu1 = v[0]; }
        else if (key == "fluid2_viscosity") { c.la_nu2 = v[0]; saw_fluid2 = true; }
        else if (key == "fluid_viscosity") { c.la_nu1 = v[0]; saw_single = true; }
        KD("surface_tension", gamma);
        KD("theta", theta_deg);
        KD("RK_beta", beta);
        KI("inlet_BC", inlet_BC);
        KI("outlet_BC", outlet_BC);
        KD("saturation_injection", sa_inject);
        KD("target_inject_pore_volume", target_inject_pore_volume);
        KD("initial_interface_position", interface_z0);
        KD("capillary_number", ca_0);
        KD("body_force_0", force_z0);
        KD("target_fluid1_saturation", sa_target);
        else if (key == "max_time_step") c.ntime_max = (long long)v[0];
        KI("max_time_step_benchmark", ntime_max_benchmark);
        KI("ntime_visual", ntime_visual);
        KI("ntime_animation", ntime_animation);
        KI("monitor_timer", ntime_monitor);
        KI("monitor_profile_timer_ratio", ntime_monitor_profile_ratio);
        KI("computation_time_timer", ntime_clock_sum);
        KI("display_steps_timer", ntime_display_steps);
        KD("checkpoint_save_timer", checkpoint_save_timer);
        KD("checkpoint_2rd_save_timer", checkpoint_2rd_save_timer);
        KD("simulation_duration_timer", simulation_duration_timer);
        KD("d_vol_animation", d_vol_animation);
        KD("d_vol_detail", d_vol_detail);
        KD("d_vol_monitor", d_vol_monitor);
        KI("MRT_collision_parameter_preset", mrt_para_preset);
        KD("char_length", char_length);
        KD("Reynolds_number", Re);
        KD("rho_drop", rho_drop);
        // unknown keys are ignored like the reference's select-case default
#undef KI
#undef KD
    }
    if (saw_single && !saw_fluid2) c.multiphase = 0;
    return check_parameters();
}

// MP/IO_multiphase.F90:455-552: every failure is MPI_Barrier + mpi_abort in the reference
bool Driver::check_parameters() {
    if (c.iper == 1 || c.domain_wall_status_x_max == 0 || c.domain_wall_status_x_min == 0) {
        error = "Error: X direction periodic BC enabled or non-slip BC not applied at x = xmin or x = xmax! Exiting program!";
        return false;
    }
    if (c.jper == 0 && (c.domain_wall_status_y_max == 0 || c.domain_wall_status_y_min == 0)) {
        error = "Error: non-slip BC not applied at y = ymin or y = ymax while y direction periodic BC not enabled! Exiting program!";
        return false;
    }
    if (c.jper == 1 && (c.domain_wall_status_y_max == 1 || c.domain_wall_status_y_min == 1)) {
        error = "Error: non-slip BC applied at y = ymin or y = ymax while y direction periodic BC enabled! Exiting program!";
        return false;
    }
    if (c.kper == 1 && (c.domain_wall_status_z_max == 1 || c.domain_wall_status_z_min == 1)) {
        error = "Error: non-slip BC applied at z = zmin or z = zmax while z direction periodic BC enabled! Exiting program!";
        return false;
    }
    if (c.npx != 1) {
        error = "MPI error: MPI_process_num_X is not equal to 1! Exiting program!";
        return false;
    }
    if (c.npy != 1) {
        error = "y decomposition is not supported by the B200 path (z slabs only): MPI_process_num must be 1,1,npz";
        return false;
    }
    if (c.npz < 1 || c.nzGlobal % c.npz != 0) {
        error = "nzGlobal must be divisible by npz (equal slabs, template-simulation_control.txt:105-106)";
        return false;
    }
    if ((c.npz > 1 || c.kper == 1) && c.iz_async == 0) {
        error = "MPI error: iz_async is zero when MPI communication along z direction is enabled! Exiting program!";
        return false;
    }
    if (c.multiphase && c.outlet_BC == 1 && c.inlet_BC == 2) {
        error = "Inlet/outlet boundary condition error: Inlet pressure + outlet convective BC is not supported! Exiting program!";
        return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------------
bool Driver::read_walls(const std::string &path) {  // MP/Misc.F90:247-295
    FILE *fp = std::fopen(path.c_str(), "rb");
    if (!fp) {
        error = "Error! No external geometry file found! Exiting program!";
        return false;
    }
    int32_t dims[3];
    if (std::fread(dims, 4, 3, fp) != 3) {
        std::fclose(fp);
        error = "wall array file: short header";
        return false;
    }
    const int nxs = dims[0], nys = dims[1], nzs = dims[2];
    if (c.nxGlobal < nxs || c.nyGlobal < nys || c.nzGlobal < nzs) {
        std::fclose(fp);
        error = "Error! Domain size is smaller than porous media sample size! Exiting program!";
        return false;
    }
    walls_global.assign((size_t)c.nxGlobal * c.nyGlobal * c.nzGlobal, 0);
    wk0 = 1;
    wk1 = c.nzGlobal;
    std::vector<int8_t> row(nxs);
    for (int k = 1; k <= nzs; k++)
        for (int j = 1; j <= nys; j++) {
            if (std::fread(row.data(), 1, nxs, fp) != (size_t)nxs) {
                std::fclose(fp);
                error = "wall array file: short read";
                return false;
            }
            std::memcpy(&walls_global[(size_t)c.nxGlobal * ((size_t)(j - 1) + (size_t)c.nyGlobal * (k - 1))], row.data(), nxs);
        }
    std::fclose(fp);
    if (c.domain_wall_status_x_max == 1 && c.domain_wall_status_y_max == 1) {  // pad with solid walls
        for (int k = 1; k <= c.nzGlobal; k++)
            for (int j = 1; j <= c.nyGlobal; j++)
                for (int i = 1; i <= c.nxGlobal; i++)
                    if (j >= nys || i >= nxs) walls_global[(size_t)(i - 1) + (size_t)c.nxGlobal * ((size_t)(j - 1) + (size_t)c.nyGlobal * (k - 1))] = 1;
    }
    return true;
}

// The reference reads the whole wall array on rank 0 and broadcasts it to every rank (MP/Misc.F90:111-140,
// MP/Mpi_misc.F90:337-502), which is what stops it at ~1e9 nodes.  Here a rank seeks to the planes
// [slab - margin, slab + margin] of the same file (clipped at the ends of an open lattice, wrapped on a z-periodic
// one) and keeps nothing else; set_walls and the geometry preprocessing work on that window (wk0..wk1).
bool Driver::read_walls_window(const std::string &path, int margin) {
    FILE *fp = std::fopen(path.c_str(), "rb");
    if (!fp) {
        error = "Error! No external geometry file found! Exiting program!";
        return false;
    }
    int32_t dims[3];
    if (std::fread(dims, 4, 3, fp) != 3) {
        std::fclose(fp);
        error = "wall array file: short header";
        return false;
    }
    const int nxs = dims[0], nys = dims[1], nzs = dims[2];
    const int nxG = c.nxGlobal, nyG = c.nyGlobal, nzG = c.nzGlobal;
    if (nxG < nxs || nyG < nys || nzG < nzs) {
        std::fclose(fp);
        error = "Error! Domain size is smaller than porous media sample size! Exiting program!";
        return false;
    }
    const int nzl = nzG / c.npz;
    int k0 = idz * nzl + 1 - margin, k1 = idz * nzl + nzl + margin;
    if (c.kper == 0) {
        k0 = std::max(k0, 1);
        k1 = std::min(k1, nzG);
    }
    const int nplanes = k1 - k0 + 1;
    if (c.kper != 0 && nplanes >= nzG) {  // the window would wrap onto itself: hold the whole lattice instead
        std::fclose(fp);
        return read_walls(path);
    }
    walls_global.assign((size_t)nxG * nyG * nplanes, 0);
    wk0 = k0;
    wk1 = k1;
    const bool pad = c.domain_wall_status_x_max == 1 && c.domain_wall_status_y_max == 1;
    std::vector<int8_t> row(nxs);
    for (int kk = 0; kk < nplanes; kk++) {
        const int k = ((k0 + kk - 1) % nzG + nzG) % nzG + 1;  // source plane, wrapped when periodic
        int8_t *plane = &walls_global[(size_t)nxG * nyG * kk];
        if (k <= nzs) {
            if (std::fseek(fp, 12L + (long)nxs * nys * (long)(k - 1), SEEK_SET) != 0) {
                std::fclose(fp);
                error = "wall array file: seek failed";
                return false;
            }
            for (int j = 1; j <= nys; j++) {
                if (std::fread(row.data(), 1, nxs, fp) != (size_t)nxs) {
                    std::fclose(fp);
                    error = "wall array file: short read";
                    return false;
                }
                std::memcpy(plane + (size_t)nxG * (j - 1), row.data(), nxs);
            }
        }
        if (pad)  // same rule as read_walls: solid beyond the sample in x / y
            for (int j = 1; j <= nyG; j++)
                for (int i = 1; i <= nxG; i++)
                    if (j >= nys || i >= nxs) plane[(size_t)(i - 1) + (size_t)nxG * (j - 1)] = 1;
    }
    std::fclose(fp);
    return true;
}

void Driver::modify_geometry() {  // MP/Misc.F90:213-244
    const int nxG = c.nxGlobal, nyG = c.nyGlobal, nzG = c.nzGlobal;
    const double xc = 0.5 * (double)(nxG + 1), yc = 0.5 * (double)(nyG + 1), zc = 0.5 * (double)(nzG + 1);
    const double r1 = 0.25 * nyG, r2 = nyG * 0.5;
    const int buffer = 10;
#pragma omp parallel for
    // every plane of the held window, the wrapped ghost planes of a z-periodic window included: they image the global plane
    // kg, which the rank holding it modifies too (the reference modifies the whole global array before distributing it)
    for (int kw = wk0; kw <= wk1; kw++) {
        const int k = ((kw - 1) % nzG + nzG) % nzG + 1;
        if (k != kw && c.kper != 1) continue;  // outside an open lattice: not a lattice plane
        for (int j = 1; j <= nyG; j++)
            for (int i = 1; i <= nxG; i++) {
                const double dx = i - xc, dy = j - yc, dz = k - zc;
                int8_t &w = walls_global[(size_t)(i - 1) + (size_t)nxG * ((size_t)(j - 1) + (size_t)nyG * (kw - wk0))];
                if (dx * dx + dy * dy + dz * dz < r1 * r1) w = 1;
                if (dx * dx + dy * dy > r2 * r2 && k > buffer && k < nzG - buffer + 1) w = 1;
            }
    }
}

void Driver::set_walls() {  // MP/Misc.F90:6-210, pore_profile :298-365, transport_walls MP/Mpi_misc.F90:337-502
    const int nxG = c.nxGlobal, nyG = c.nyGlobal, nzG = c.nzGlobal;
    nx = nxG;
    ny = nyG;
    nz = nzG / c.npz;
    if (wk1 < wk0) {  // no geometry handed over: empty duct over the whole lattice
        wk0 = 1;
        wk1 = nzG;
        walls_global.assign((size_t)nxG * nyG * nzG, 0);
    }
    // G(i,j,k): k is a global plane index inside the held window; on a periodic lattice the caller supplies the
    // wrapped planes of a window that crosses the ends, a whole-lattice array wraps here.
    const bool whole = wk0 == 1 && wk1 == nzG;
    auto kmap = [&](int k) { return (whole && c.kper == 1) ? ((k - 1) % nzG + nzG) % nzG + 1 : k; };
    auto G = [&](int i, int j, int k) -> int8_t & { return walls_global[(size_t)(i - 1) + (size_t)nxG * ((size_t)(j - 1) + (size_t)nyG * (size_t)(kmap(k) - wk0))]; };
    if (c.modify_geometry_cmd == 1) modify_geometry();
#pragma omp parallel for
    for (int k = wk0; k <= wk1; k++) {
        const int kg = c.kper == 1 ? ((k - 1) % nzG + nzG) % nzG + 1 : k;  // global plane this window plane images
        for (int j = 1; j <= nyG; j++)
            for (int i = 1; i <= nxG; i++) {
                if ((c.domain_wall_status_z_min == 1 && kg == 1) || (c.domain_wall_status_z_max == 1 && kg == nzG) ||
                    (c.domain_wall_status_x_min == 1 && i == 1) || (c.domain_wall_status_x_max == 1 && i == nxG) ||
                    (c.domain_wall_status_y_min == 1 && j == 1) || (c.domain_wall_status_y_max == 1 && j == nyG))
                    walls_global[(size_t)(i - 1) + (size_t)nxG * ((size_t)(j - 1) + (size_t)nyG * (size_t)(k - wk0))] = 1;
            }
    }
    const V2 w2{(size_t)nx + 4, (size_t)ny + 4};
    walls.assign((size_t)(nx + 4) * (ny + 4) * (nz + 4), 0);
#pragma omp parallel for
    for (int k = 1; k <= nz; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) walls[w2(i, j, k)] = G(i, j, idz * nz + k);
#pragma omp parallel for
    for (int k = -1; k <= nz + 2; k++)
        for (int j = -1; j <= ny + 2; j++)
            for (int i = -1; i <= nx + 2; i++) {
                if ((idz == 0 && k <= 1 && c.domain_wall_status_z_min == 1) || (idz == c.npz - 1 && k >= nz && c.domain_wall_status_z_max == 1) ||
                    (i <= 1 && c.domain_wall_status_x_min == 1) || (i >= nx && c.domain_wall_status_x_max == 1) ||
                    (j <= 1 && c.domain_wall_status_y_min == 1) || (j >= ny && c.domain_wall_status_y_max == 1))
                    walls[w2(i, j, k)] = 1;
            }
    if (wk0 <= 1 && wk1 >= 1) {
        int icount = 0;
        for (int j = 1; j <= nyG; j++)
            for (int i = 1; i <= nxG; i++) icount += (G(i, j, 1) <= 0);
        A_xy_effective = icount;
    }
    // pore profile over the held planes; the totals are global only when the whole lattice is held (otherwise the
    // caller sums pore_sum_local over the slabs, MP/Misc.F90:338-352)
    pore_profile_z.assign(wk1 - wk0 + 1, 0);
#pragma omp parallel for
    for (int k = wk0; k <= wk1; k++) {
        int cnt = 0;
        for (int j = 1; j <= nyG; j++)
            for (int i = 1; i <= nxG; i++) cnt += (walls_global[(size_t)(i - 1) + (size_t)nxG * ((size_t)(j - 1) + (size_t)nyG * (size_t)(k - wk0))] <= 0);
        pore_profile_z[k - wk0] = cnt;
    }
    pore_sum = 0;
    pore_sum_effective = 0;
    pore_sum_local = 0;
    for (int k = std::max(1, wk0); k <= std::min(nzG, wk1); k++) {
        pore_sum += pore_profile_z[k - wk0];
        if (k >= 1 + c.n_exclude_inlet && k <= nzG - c.n_exclude_outlet) pore_sum_effective += pore_profile_z[k - wk0];
        if (k >= idz * nz + 1 && k <= idz * nz + nz) pore_sum_local += pore_profile_z[k - wk0];
    }
    // neighbour slabs' planes into the z ghost layers (ztransport_walls(0,0,2)); periodic wrap when kper
    for (int k = 1; k <= 2; k++)
        for (int j = 1; j <= ny; j++)
            for (int i = 1; i <= nx; i++) {
                if (c.kper == 1 || idz != 0) walls[w2(i, j, k - 2)] = G(i, j, idz * nz + (k - 2));
                if (c.kper == 1 || idz != c.npz - 1) walls[w2(i, j, k + nz)] = G(i, j, idz * nz + nz + k);
            }
    // ytransport_walls(0,2,2): y is never decomposed here, so a y-periodic lattice exchanges with itself
    // (MP/Mpi_misc.F90:337-502, after the z transport: the z ghost planes take part)
    if (c.jper == 1)
        for (int k = -1; k <= nz + 2; k++)
            for (int j = 1; j <= 2; j++)
                for (int i = 1; i <= nx; i++) {
                    const int8_t lo = walls[w2(i, j, k)], hi = walls[w2(i, ny + j - 2, k)];
                    walls[w2(i, j - 2, k)] = hi;
                    walls[w2(i, j + ny, k)] = lo;
                }
}

namespace {
struct Off3 {
    signed char a, b, c;
};
// ISO8 stencil of the wall normal, grouped by weight; each entry is ws(x+o) - ws(x-o), accumulated left to right
// in the reference's order (MP/Geometry_preprocessing.F90:234-377)
constexpr int ISO8_CNT[7] = {1, 4, 4, 1, 8, 12, 4};
constexpr double ISO8_W[7] = {4.0 / 45.0, 1.0 / 21.0, 2.0 / 105.0, 5.0 / 504.0, 1.0 / 315.0, 1.0 / 630.0, 1.0 / 5040.0};
constexpr Off3 ISO8_X[34] = {{1,0,0},{1,1,0},{1,-1,0},{1,0,1},{1,0,-1},{1,1,1},{1,1,-1},{1,-1,1},{1,-1,-1},{2,0,0},
    {2,1,0},{2,-1,0},{2,0,1},{2,0,-1},{1,2,0},{1,-2,0},{1,0,2},{1,0,-2},
    {2,1,1},{2,1,-1},{2,-1,1},{2,-1,-1},{1,2,1},{1,2,-1},{1,-2,1},{1,-2,-1},{1,1,2},{1,1,-2},{1,-1,2},{1,-1,-2},
    {2,2,0},{2,-2,0},{2,0,2},{2,0,-2}};
constexpr Off3 ISO8_Y[34] = {{0,1,0},{1,1,0},{-1,1,0},{0,1,1},{0,1,-1},{1,1,1},{1,1,-1},{-1,1,-1},{-1,1,1},{0,2,0},
    {2,1,0},{-2,1,0},{0,2,1},{0,2,-1},{1,2,0},{-1,2,0},{0,1,2},{0,1,-2},
    {2,1,1},{2,1,-1},{-2,1,1},{-2,1,-1},{1,2,1},{1,2,-1},{-1,2,1},{-1,2,-1},{1,1,2},{1,1,-2},{-1,1,2},{-1,1,-2},
    {2,2,0},{-2,2,0},{0,2,2},{0,2,-2}};
constexpr Off3 ISO8_Z[34] = {{0,0,1},{0,1,1},{0,-1,1},{1,0,1},{-1,0,1},{1,1,1},{1,-1,1},{-1,1,1},{-1,-1,1},{0,0,2},
    {0,1,2},{0,-1,2},{2,0,1},{-2,0,1},{0,2,1},{0,-2,1},{1,0,2},{-1,0,2},
    {2,1,1},{2,-1,1},{-2,1,1},{-2,-1,1},{1,2,1},{1,-2,1},{-1,2,1},{-1,-2,1},{1,1,2},{1,-1,2},{-1,1,2},{-1,-1,2},
    {0,2,2},{0,-2,2},{2,0,2},{-2,0,2}};

double iso8(const double *ws, size_t c, ptrdiff_t sy, ptrdiff_t sz, const Off3 *tab) {
    double res = 0.0;
    int t = 0;
    for (int grp = 0; grp < 7; grp++) {
        double acc = 0.0;
        for (int m = 0; m < ISO8_CNT[grp]; m++, t++) {
            const ptrdiff_t o = tab[t].a + sy * tab[t].b + sz * tab[t].c;
            if (m == 0) acc = ws[c + o] - ws[c - o];
            else {
                acc = acc + ws[c + o];
                acc = acc - ws[c - o];
            }
        }
        res = grp == 0 ? ISO8_W[0] * acc : res + ISO8_W[grp] * acc;
    }
    return res;
}
}  // namespace

void Driver::geometry_preprocessing_new() {  // MP/Geometry_preprocessing.F90:9-512
    // The reference processes the whole lattice on rank 0 (two FP64 copies with 10 ghost layers) and broadcasts the
    // global lists.  Everything in it is a local stencil (18-neighbour classification, 4 x 27-point smoothing,
    // radius-2 ISO8), so the same numbers are obtained from a z window of the wall array that extends >= 12 planes
    // beyond the slab; at the lattice ends the window is extended by the reference's replicate / periodic rule.
    const int nxG = c.nxGlobal, nyG = c.nyGlobal, nzG = c.nzGlobal;
    const int gl = 6 + 4, ophi = 4;
    const bool whole = wk0 == 1 && wk1 == nzG;
    // z range of the extended array
    int ek0 = wk0, ek1 = wk1;
    if (whole || c.kper == 0) {
        if (wk0 == 1) ek0 = 1 - gl;
        if (wk1 == nzG) ek1 = nzG + gl;
    }
    const ptrdiff_t ex = nxG + 2 * gl, ey = nyG + 2 * gl, ez = ek1 - ek0 + 1;
    const size_t ntot = (size_t)ex * ey * ez;
    auto E = [&](int i, int j, int k) { return (size_t)(i + gl - 1) + (size_t)ex * ((size_t)(j + gl - 1) + (size_t)ey * (size_t)(k - ek0)); };
    std::vector<int8_t> wt(ntot, 0);
    std::vector<double> ws1(ntot), ws2(ntot);
#pragma omp parallel for
    for (int k = wk0; k <= wk1; k++)
        for (int j = 1; j <= nyG; j++)
            for (int i = 1; i <= nxG; i++) wt[E(i, j, k)] = walls_global[(size_t)(i - 1) + (size_t)nxG * ((size_t)(j - 1) + (size_t)nyG * (size_t)(k - wk0))];
    // replicate (or wrap, when periodic) into the ghost layers: z, then y, then x (:56-108)
    for (int j = 1; j <= nyG; j++)
        for (int i = 1; i <= nxG; i++)
            for (int g = 1; g <= gl; g++) {
                if (ek0 < wk0) wt[E(i, j, 1 - g)] = c.kper == 0 ? wt[E(i, j, 1)] : wt[E(i, j, nzG + 1 - g)];
                if (ek1 > wk1) wt[E(i, j, nzG + g)] = c.kper == 0 ? wt[E(i, j, nzG)] : wt[E(i, j, g)];
            }
#pragma omp parallel for
    for (int k = ek0; k <= ek1; k++)
        for (int i = 1; i <= nxG; i++)
            for (int g = 1; g <= gl; g++) {
                wt[E(i, 1 - g, k)] = c.jper == 0 ? wt[E(i, 1, k)] : wt[E(i, nyG + 1 - g, k)];
                wt[E(i, nyG + g, k)] = c.jper == 0 ? wt[E(i, nyG, k)] : wt[E(i, g, k)];
            }
#pragma omp parallel for
    for (int k = ek0; k <= ek1; k++)
        for (int j = 1 - gl; j <= nyG + gl; j++)
            for (int g = 1; g <= gl; g++) {
                wt[E(1 - g, j, k)] = c.iper == 0 ? wt[E(1, j, k)] : wt[E(nxG + 1 - g, j, k)];
                wt[E(nxG + g, j, k)] = c.iper == 0 ? wt[E(nxG, j, k)] : wt[E(g, j, k)];
            }
#pragma omp parallel for
    for (size_t n = 0; n < ntot; n++) ws1[n] = ws2[n] = (double)wt[n];
    // classification into a second array (the reference updates in place; the outcome is order independent because
    // 2 still tests >=1 and -1 still tests <=0) so that the loop can run in parallel
    std::vector<int8_t> cls(wt);
#pragma omp parallel for collapse(2)
    for (int k = ek0 + 1; k <= ek1 - 1; k++)
        for (int j = 2 - gl; j <= nyG + gl - 1; j++)
            for (int i = 2 - gl; i <= nxG + gl - 1; i++) {
                const size_t cc = E(i, j, k);
                if (wt[cc] == 1) {
                    for (int n = 1; n <= 18; n++)
                        if (wt[E(i + EX[n], j + EY[n], k + EZ[n])] <= 0) { cls[cc] = 2; break; }
                } else if (wt[cc] == 0) {
                    for (int n = 1; n <= 18; n++)
                        if (wt[E(i + EX[n], j + EY[n], k + EZ[n])] >= 1) { cls[cc] = -1; break; }
                }
            }
    // four passes of the 27-point smoothing (:145-169), sum order n = 0..26
    static const int sx_[27] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1};
    static const int sy_[27] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, -1, 1, -1, 1};
    static const int sz_[27] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1, -1, 1, 1, -1, -1, 1, 1, -1};
    const double we[4] = {8.0 / 27.0, 2.0 / 27.0, 1.0 / 54.0, 1.0 / 216.0};
    ptrdiff_t so[27];
    double sw[27];
    for (int n = 0; n < 27; n++) {
        so[n] = sx_[n] + ex * (sy_[n] + ey * (ptrdiff_t)sz_[n]);
        sw[n] = we[sx_[n] * sx_[n] + sy_[n] * sy_[n] + sz_[n] * sz_[n]];
    }
    for (int it = 0; it < 4; it++) {
#pragma omp parallel for collapse(2)
        for (int k = ek0 + 1; k <= ek1 - 1; k++)
            for (int j = 2 - gl; j <= nyG + gl - 1; j++)
                for (int i = 2 - gl; i <= nxG + gl - 1; i++) {
                    const size_t cc = E(i, j, k);
                    double acc = 0.0;
                    for (int n = 0; n < 27; n++) acc = acc + ws1[cc + so[n]] * sw[n];
                    ws2[cc] = acc;
                }
#pragma omp parallel for collapse(2)
        for (int k = ek0 + 1; k <= ek1 - 1; k++)
            for (int j = 2 - gl; j <= nyG + gl - 1; j++)
                for (int i = 2 - gl; i <= nxG + gl - 1; i++) ws1[E(i, j, k)] = ws2[E(i, j, k)];
    }
    // lists in k-outer / i-inner order (:171-225) restricted to what this slab keeps (:424-507): solid nodes within
    // 3 ghost layers, fluid nodes within 2; the global scan range is 1-4..n+4
    theta = (180.0 - c.theta_deg) * PI / 180.0;  // MP/IO_multiphase.F90:467-468
    solid_boundary_nodes.clear();
    fluid_boundary_nodes.clear();
    num_solid_boundary_global = num_fluid_boundary_global = 0;
    const int ks0 = whole ? 1 - ophi : std::max(idz * nz + 1 - 3, c.kper ? -(1 << 30) : 1 - ophi);
    const int ks1 = whole ? nzG + ophi : std::min(idz * nz + nz + 3, c.kper ? (1 << 30) : nzG + ophi);
    for (int k = ks0; k <= ks1; k++)
        for (int j = 1 - ophi; j <= nyG + ophi; j++)
            for (int i = 1 - ophi; i <= nxG + ophi; i++) {
                const int8_t t = cls[E(i, j, k)];
                const int kl = k - idz * nz;
                if (t == 2) {
                    num_solid_boundary_global++;
                    if (i >= -2 && i <= nx + 3 && j >= -2 && j <= ny + 3 && kl >= -2 && kl <= nz + 3) {
                        mflbm_solid_node s{};
                        s.ix = i; s.iy = j; s.iz = kl;
                        int cnt = 0;
                        s.la_weight = 0.0;
                        for (int n = 1; n <= 18; n++)
                            if (cls[E(i + EX[n], j + EY[n], k + EZ[n])] <= 0) {
                                s.la_weight = s.la_weight + wq(n);
                                s.neighbor_list[cnt++] = n;
                            }
                        s.i_fluid_num = cnt;
                        solid_boundary_nodes.push_back(s);
                    }
                } else if (t == -1) {
                    num_fluid_boundary_global++;
                    if (i >= -1 && i <= nx + 2 && j >= -1 && j <= ny + 2 && kl >= -1 && kl <= nz + 2) {
                        mflbm_fluid_node f{};
                        f.ix = i; f.iy = j; f.iz = kl; f.theta = theta;
                        fluid_boundary_nodes.push_back(f);
                    }
                }
            }
    const ptrdiff_t sy = ex, sz = ex * ey;
#pragma omp parallel for
    for (long long n = 0; n < (long long)fluid_boundary_nodes.size(); n++) {
        mflbm_fluid_node &f = fluid_boundary_nodes[n];
        const size_t cc = E(f.ix, f.iy, f.iz + idz * nz);
        const double nwx = iso8(ws2.data(), cc, sy, sz, ISO8_X);
        const double nwy = iso8(ws2.data(), cc, sy, sz, ISO8_Y);
        const double nwz = iso8(ws2.data(), cc, sy, sz, ISO8_Z);
        const double tmp = 1.0 / (std::sqrt(nwx * nwx + nwy * nwy + nwz * nwz) + EPS_MP);
        f.nwx = nwx * tmp; f.nwy = nwy * tmp; f.nwz = nwz * tmp;
    }
}

// the same through the C ABI: the device kernels classify, list and compute the wall normals of this slab's window
bool Driver::geometry_preprocessing_device() {
    mflbm_geometry_config gc;
    memset(&gc, 0, sizeof(gc));
    gc.struct_size = (int32_t)sizeof(gc);
    gc.nxGlobal = c.nxGlobal; gc.nyGlobal = c.nyGlobal; gc.nzGlobal = c.nzGlobal;
    gc.wk0 = wk0; gc.wk1 = wk1;
    gc.idz = idz; gc.npz = c.npz;
    gc.iper = c.iper; gc.jper = c.jper; gc.kper = c.kper;
    gc.device = geometry_device;
    theta = (180.0 - c.theta_deg) * PI / 180.0;  // MP/IO_multiphase.F90:467-468
    gc.theta = theta;
    mflbm_solid_node *ps = nullptr;
    mflbm_fluid_node *pf = nullptr;
    int32_t ns = 0, nf = 0;
    int64_t gs = 0, gf = 0;
    if (mflbm_geometry_preprocess(&gc, walls_global.data(), &ps, &ns, &pf, &nf, &gs, &gf) != MFLBM_OK) {
        error = std::string("mflbm_geometry_preprocess: ") + mflbm_geometry_last_error();
        return false;
    }
    solid_boundary_nodes.assign(ps, ps + ns);
    fluid_boundary_nodes.assign(pf, pf + nf);
    num_solid_boundary_global = gs;
    num_fluid_boundary_global = gf;
    mflbm_geometry_free(ps);
    mflbm_geometry_free(pf);
    return true;
}

// ---------------------------------------------------------------------------------------------------
// initialisation
// ---------------------------------------------------------------------------------------------------
namespace {
int32_t ipow_wrap(int32_t n, int e) {  // default-integer n**e as gfortran evaluates it (wraps for n**5, n >= 75)
    uint32_t r = 1;
    for (int i = 0; i < e; i++) r *= (uint32_t)n;
    return (int32_t)r;
}
}  // namespace

void Driver::inlet_vel_profile_rectangular(double vel_avg, int num_terms) {  // MP/Misc.F90:625-665
    const double a = 0.5 * la_x, b = 0.5 * la_y;
    double tmp1 = 0.0;
    for (int n = 1; n <= num_terms; n += 2) tmp1 = tmp1 + (std::tanh(0.5 * (double)n * PI * b / a)) / ipow_wrap(n, 5);
    const double pi2 = PI * PI, pi5 = pi2 * pi2 * PI, pim3 = 1.0 / (pi2 * PI);
    double tmp2 = 1.0 - 192.0 / pi5 * (a / b) * tmp1;
    tmp2 = -3.0 * vel_avg / (tmp2 * (a * a));
    const size_t sx = nx + 2;
#pragma omp parallel for
    for (int j = 1; j <= ny; j++)
        for (int i = 1; i <= nx; i++) {
            if (i > 1 && i < c.nxGlobal && j > 1 && j < c.nyGlobal) {
                const double xx = i - 1.5 - a, yy = j - 1.5 - b;
                double tmp3 = 0.0;
                for (int n = 1; n <= num_terms; n += 2) {
                    const double sgn = std::pow(-1.0, 0.5 * (double)(n - 1));
                    tmp3 = tmp3 + sgn * std::cos(0.5 * n * PI * xx / a) / ipow_wrap(n, 3) *
                                      (1.0 - (std::exp(0.5 * n * PI * (yy - b) / a) + std::exp(0.5 * n * PI * (-yy - b) / a)) /
                                                 (1.0 + std::exp(0.5 * n * PI * (-b - b) / a)));
                }
                w_in[(size_t)i + sx * j] = tmp3 * (-16.0 * tmp2 * (a * a) * pim3);
            }
        }
}

void Driver::initialization_basic() {  // MP/Init_multiphase.F90:68-150,194-236 ; SP/Initialization.F90:76-130,157-197
    la_z = c.nzGlobal - 1;
    la_y = c.nyGlobal - 1 - 0.5f - 0.5f;
    la_x = c.nxGlobal - 1 - 0.5f - 0.5f;
    A_xy = la_x * la_y;
    la_nui1 = 1.0 / c.la_nu1;
    la_nui2 = c.multiphase ? 1.0 / c.la_nu2 : 0.0;
    theta = (180.0 - c.theta_deg) * PI / 180.0;
    phi_inlet = 2.0 * c.sa_inject - 1.0;
    force_Z = c.force_z0;
    rho_out = 1.0;
    rho_in = 1.0;
    relaxation = 1.0;  // MP/Main_multiphase.F90:86
    if (!c.multiphase) {  // SP/Initialization.F90:87-112 incl. the duplicated preset==1 test (preset 2 == SRT)
        const double omega = 1.0 / (3.0 * c.la_nu1 + 0.5);
        s_nu = s_e = s_e2 = s_pi = s_q = s_t = omega;
        if (c.mrt_para_preset == 1) {
            s_q = 8.0 * (2.0 - omega) / (8.0 - omega);
            s_t = s_q;
        }
    }
    w_in.assign((size_t)(nx + 2) * (ny + 2), 0.0);
    const bool open_z = c.kper == 0 && c.domain_wall_status_z_min == 0 && c.domain_wall_status_z_max == 0;
    if (open_z) {
        if (c.inlet_BC == 1) {
            if (c.multiphase) {
                force_Z = 0.0;
                uin_avg_0 = c.ca_0 * c.gamma / c.la_nu1;
            } else {
                uin_avg_0 = c.Re * c.la_nu1 / c.char_length;
            }
            uin_avg = uin_avg_0;
            flowrate = uin_avg_0 * A_xy;
            for (int j = 1; j <= ny; j++)
                for (int i = 1; i <= nx; i++)
                    w_in[(size_t)i + (size_t)(nx + 2) * j] = (i > 1 && i < c.nxGlobal && j > 1 && j < c.nyGlobal) ? uin_avg : 0.0;
            inlet_vel_profile_rectangular(uin_avg_0, 1000);
            if (c.target_inject_pore_volume > 0) {
                c.ntime_max = (long long)((double)(c.target_inject_pore_volume * pore_sum) / flowrate);
                if (c.ntime_max % 2 == 1) c.ntime_max += 1;
            }
        } else if (c.inlet_BC == 2) {
            if (c.multiphase) rho_in = rho_out - (-c.force_z0 / 3.0) * c.nzGlobal;
            else rho_in = rho_out + c.rho_drop;
        }
    }
    if (c.d_vol_monitor > 0 && c.inlet_BC == 1 && open_z) {  // MP/Init_multiphase.F90:174-184
        c.ntime_monitor = (int)((double)(c.d_vol_monitor * pore_sum) / flowrate);
        if (c.ntime_monitor % 2 == 1) c.ntime_monitor += 1;
    }
}

// equilibrium population at rest, rho = 1 (MP/Init_multiphase.F90:262-265, :370-412 ; SP/Initialization.F90:226-250)
double Driver::pdf_value(int i, int j, int k, int q, int fluid) const {
    const V4 v4{(size_t)nx + 8, (size_t)ny + 8};
    const double rho = 1.0, usqrt = 0.0;
    const double ph = c.multiphase ? phi[v4(i, j, k)] : 0.0;
    const double r = !c.multiphase ? rho : (fluid == 0 ? rho * (1.0 + ph) * 0.5 : rho * (1.0 - ph) * 0.5);
    const double w = wq(q);
    return r * w + r * w * (3.0 * 0.0 + 4.5 * 0.0 * 0.0 - 1.5 * usqrt);
}

void Driver::fill_pdf(std::vector<double> &a, int q, int fluid) const {
    const V1 v1{(size_t)nx + 2, (size_t)ny + 2};
    a.resize((size_t)(nx + 2) * (ny + 2) * (nz + 2));
#pragma omp parallel for collapse(2)
    for (int k = 0; k <= nz + 1; k++)
        for (int j = 0; j <= ny + 1; j++)
            for (int i = 0; i <= nx + 1; i++) a[v1(i, j, k)] = pdf_value(i, j, k, q, fluid);
}

void Driver::initialization_new() {  // MP/Init_multiphase.F90:243-470 ; SP/Initialization.F90:201-309
    const V4 v4{(size_t)nx + 8, (size_t)ny + 8};
    const bool open_z = c.kper == 0 && c.domain_wall_status_z_min == 0 && c.domain_wall_status_z_max == 0;
    if (c.multiphase) {
        const bool keep = c.initial_fluid_distribution_option == 6 && phi.size() == (size_t)(nx + 8) * (ny + 8) * (nz + 8);
        if (!keep) phi.assign((size_t)(nx + 8) * (ny + 8) * (nz + 8), 0.0);
        const double z0 = c.interface_z0;
#pragma omp parallel for collapse(2)
        for (int k = -3; k <= nz + 4; k++)
            for (int j = -3; j <= ny + 4; j++)
                for (int i = -3; i <= nx + 4; i++) {
                    const double x = i, y = j, z = idz * nz + k;
                    double v;
                    switch (c.initial_fluid_distribution_option) {
                    case 1: v = z <= z0 ? 1.0 : -1.0; break;
                    case 2: v = z <= z0 ? -1.0 : 1.0; break;
                    case 3: case 4: {
                        const double dx = x - (c.nxGlobal + 1) * 0.0, dz = z - (c.nzGlobal + 1) * 0.5, dy = y - (c.nyGlobal + 1) * 0.5;
                        const bool in = dx * dx + dz * dz + dy * dy <= z0 * z0;
                        v = (c.initial_fluid_distribution_option == 3) == in ? 1.0 : -1.0;
                    } break;
                    case 5: {
                        const double dx = x - (c.nxGlobal + 1) * 0.5, dz = z - (c.nzGlobal + 1) * 0.5, dy = y - (c.nyGlobal + 1) * 0.5;
                        v = dx * dx + dz * dz + dy * dy <= z0 * z0 ? 1.0 : -1.0;
                    } break;
                    default: v = phi[v4(i, j, k)]; break;  // option 6 (unseeded random_number in the reference): caller-provided field
                    }
                    if (open_z && z <= 0) v = phi_inlet;
                    phi[v4(i, j, k)] = v;
                }
    }
    if (!lazy_pdfs)
        for (int q = 0; q < 19; q++) {
            fill_pdf(f[q], q, 0);
            if (c.multiphase) fill_pdf(g[q], q, 1);
        }
    const size_t np = (size_t)(nx + 2) * (ny + 2);
    f_convec_bc.assign(np * 19, 0.0);
    if (c.multiphase) {
        g_convec_bc.assign(np * 19, 0.0);
        phi_convec_bc.assign(np, 0.0);
    }
    if (c.outlet_BC == 1 && idz == c.npz - 1) {
        for (int j = 0; j <= ny + 1; j++)
            for (int i = 0; i <= nx + 1; i++) {
                for (int q = 0; q < 19; q++) {
                    f_convec_bc[(size_t)i + (size_t)(nx + 2) * ((size_t)j + (size_t)(ny + 2) * q)] = pdf_value(i, j, nz, q, 0);
                    if (c.multiphase) g_convec_bc[(size_t)i + (size_t)(nx + 2) * ((size_t)j + (size_t)(ny + 2) * q)] = pdf_value(i, j, nz, q, 1);
                }
                if (c.multiphase) phi_convec_bc[(size_t)i + (size_t)(nx + 2) * j] = phi[v4(i, j, nz)];
            }
    }
}

// ---------------------------------------------------------------------------------------------------
// device side: everything below is a thin forwarder to the C ABI
// ---------------------------------------------------------------------------------------------------
bool Driver::create_context(int device, int use_nccl, const unsigned char *nccl_id, int kernel_variant) {
    mflbm_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = (int32_t)sizeof cfg;
    cfg.solver = c.multiphase ? MFLBM_SOLVER_MULTIPHASE : MFLBM_SOLVER_SINGLEPHASE;
    cfg.nx = nx; cfg.ny = ny; cfg.nz = nz;
    cfg.nxGlobal = c.nxGlobal; cfg.nyGlobal = c.nyGlobal; cfg.nzGlobal = c.nzGlobal;
    cfg.idz = idz; cfg.npz = c.npz; cfg.jper = c.jper; cfg.kper = c.kper;
    cfg.domain_wall_status_z_min = c.domain_wall_status_z_min;
    cfg.domain_wall_status_z_max = c.domain_wall_status_z_max;
    cfg.inlet_BC = c.inlet_BC; cfg.outlet_BC = c.outlet_BC;
    cfg.porous_plate_cmd = c.porous_plate_cmd; cfg.Z_porous_plate = c.Z_porous_plate;
    cfg.mrt = c.mrt; cfg.iz_async = c.iz_async;
    cfg.num_solid_boundary = (int32_t)solid_boundary_nodes.size();
    cfg.num_fluid_boundary = (int32_t)fluid_boundary_nodes.size();
    cfg.device = device; cfg.use_nccl = use_nccl; cfg.kernel_variant = kernel_variant;
    cfg.la_nui1 = la_nui1; cfg.la_nui2 = la_nui2; cfg.gamma = c.gamma; cfg.beta = c.beta; cfg.force_Z = force_Z;
    cfg.phi_inlet = phi_inlet; cfg.sa_inject = c.sa_inject; cfg.relaxation = relaxation; cfg.uin_avg = uin_avg;
    cfg.rho_in = rho_in; cfg.rho_out = rho_out;
    cfg.s_e = s_e; cfg.s_e2 = s_e2; cfg.s_q = s_q; cfg.s_nu = s_nu; cfg.s_pi = s_pi; cfg.s_t = s_t;
    if (nccl_id) std::memcpy(cfg.nccl_unique_id, nccl_id, 128);
    if (ctx) { mflbm_destroy(ctx); ctx = nullptr; }
    if (mflbm_create(&cfg, &ctx) != MFLBM_OK) {
        error = mflbm_last_error(nullptr);
        return false;
    }
    return true;
}

// initial_fluid_distribution_option 6 (MP/Init_multiphase.F90:306-311) with a SEEDED generator: the reference draws
// random_number() per node after an unseeded random_seed() (irreproducible); here the draw is a hash of
// the node's GLOBAL position, so every z-slab decomposition sees the same field.  phi = 1 where u <= sa_target.
void Driver::random_phase_field(unsigned long long seed) {
    const V4 v4{(size_t)nx + 8, (size_t)ny + 8};
    phi.assign((size_t)(nx + 8) * (ny + 8) * (nz + 8), 0.0);
    const long long nzG = c.nzGlobal;
#pragma omp parallel for collapse(2)
    for (int k = -3; k <= nz + 4; k++)
        for (int j = -3; j <= ny + 4; j++)
            for (int i = -3; i <= nx + 4; i++) {
                long long z = (long long)idz * nz + k;
                if (c.kper) z = ((z - 1) % nzG + nzG) % nzG + 1;
                unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)((i + 3) + (long long)(nx + 8) * ((j + 3) + (long long)(ny + 8) * (z + 3)) + 1);
                x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;  // splitmix64 finaliser
                x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
                x ^= x >> 31;
                const double u = (double)(x >> 11) * (1.0 / 9007199254740992.0);
                phi[v4(i, j, k)] = u > c.sa_target ? -1.0 : 1.0;
            }
}

// start over from a new initial fluid distribution on the SAME geometry: initialization_new_multi again, then the
// fields (phi, populations, convective-outlet state) go to the existing context; walls and node lists stay
bool Driver::reinitialize(int option, unsigned long long seed) {
    if (!ctx) { error = "reinitialize before create_context"; return false; }
    c.initial_fluid_distribution_option = option;
    if (option == 6) random_phase_field(seed);
    initialization_new();
    return upload(false);
}

bool Driver::upload(bool with_geometry) {
    mflbm_arrays a;
    std::memset(&a, 0, sizeof a);
    if (with_geometry) a.walls = walls.data();
    a.w_in = w_in.data();
    a.f_convec_bc = f_convec_bc.data();
    if (c.multiphase) {
        a.phi = phi.data();
        a.g_convec_bc = g_convec_bc.data();
        a.phi_convec_bc = phi_convec_bc.data();
        if (with_geometry) {
            a.solid_boundary_nodes = solid_boundary_nodes.data();
            a.fluid_boundary_nodes = fluid_boundary_nodes.data();
        }
    }
    if (!lazy_pdfs) {
        for (int q = 0; q < 19; q++) {
            a.f[q] = f[q].data();
            if (c.multiphase) a.g[q] = g[q].data();
        }
    }
    if (mflbm_upload(ctx, &a) != MFLBM_OK) {
        error = mflbm_last_error(ctx);
        return false;
    }
    if (lazy_pdfs) {  // large lattices: generate and hand over one population array at a time (host memory = 1 array)
        std::vector<double> tmp;
        for (int fl = 0; fl < (c.multiphase ? 2 : 1); fl++)
            for (int q = 0; q < 19; q++) {
                fill_pdf(tmp, q, fl);
                std::memset(&a, 0, sizeof a);
                (fl == 0 ? a.f : a.g)[q] = tmp.data();
                if (mflbm_upload(ctx, &a) != MFLBM_OK) {
                    error = mflbm_last_error(ctx);
                    return false;
                }
            }
    }
    return true;
}

// save_checkpoint (MP/IO_multiphase.F90:562-642 ; SP/IO.F90 likewise): the per-rank stream file
//   int32 ntime+1, f64 force_z, f64 rho_in, f0..f18 [, g0..g18, phi] [, f_convec_bc [, g_convec_bc, phi_convec_bc]]
// with every array in the reference's extents (ghost layers included), written one array at a time from the staged
// device snapshot (mflbm_checkpoint_begin / _fetch / _end): the device-to-host copies overlap whatever steps the caller
// queued after mflbm_checkpoint_begin, and the host never holds more than one array.
bool Driver::save_checkpoint(const std::string &path, int ntime) {
    FILE *fp = std::fopen(path.c_str(), "wb");
    if (!fp) { error = "cannot open " + path; return false; }
    bool ok = true;
    auto fail_ = [&](const std::string &m) { error = m; ok = false; };
    int rc = mflbm_checkpoint_begin(ctx);  // no-op error if the caller already staged it earlier
    if (rc < 0 && std::string(mflbm_last_error(ctx)).find("already pending") == std::string::npos) fail_(mflbm_last_error(ctx));
    const int32_t nt = ntime + 1;
    if (ok && (std::fwrite(&nt, 4, 1, fp) != 1 || std::fwrite(&force_Z, 8, 1, fp) != 1 || std::fwrite(&rho_in, 8, 1, fp) != 1)) fail_("write error");
    std::vector<double> tmp;
    auto put = [&](size_t slot, size_t n) {
        if (!ok) return;
        tmp.assign(n, 0.0);  // dead entries of the sparse layout (never read by anyone) are written as zeros
        mflbm_arrays a;
        std::memset(&a, 0, sizeof a);
        *(double **)((char *)&a + slot) = tmp.data();
        if (mflbm_checkpoint_fetch(ctx, &a) != MFLBM_OK) return fail_(mflbm_last_error(ctx));
        if (std::fwrite(tmp.data(), 8, n, fp) != n) fail_("write error");
    };
    const size_t n1 = (size_t)(nx + 2) * (ny + 2) * (nz + 2), n4 = (size_t)(nx + 8) * (ny + 8) * (nz + 8), np = (size_t)(nx + 2) * (ny + 2);
    for (int q = 0; q < 19; q++) put(offsetof(mflbm_arrays, f) + q * sizeof(double *), n1);
    if (c.multiphase) {
        for (int q = 0; q < 19; q++) put(offsetof(mflbm_arrays, g) + q * sizeof(double *), n1);
        put(offsetof(mflbm_arrays, phi), n4);
    }
    if (c.outlet_BC == 1) {
        put(offsetof(mflbm_arrays, f_convec_bc), np * 19);
        if (c.multiphase) {
            put(offsetof(mflbm_arrays, g_convec_bc), np * 19);
            put(offsetof(mflbm_arrays, phi_convec_bc), np);
        }
    }
    if (mflbm_checkpoint_end(ctx) != MFLBM_OK && ok) fail_(mflbm_last_error(ctx));
    std::fclose(fp);
    return ok;
}

// initialization_old_multi (MP/Init_multiphase.F90:477-557): read the stream file back, one array at a time, straight
// into the existing context; returns ntime0 (the step to continue with)
bool Driver::initialization_old(const std::string &path, int *ntime0) {
    FILE *fp = std::fopen(path.c_str(), "rb");
    if (!fp) { error = "Checkpoint data not found! Exiting program!"; return false; }
    bool ok = true;
    auto fail_ = [&](const std::string &m) { error = m; ok = false; };
    int32_t nt = 0;
    if (std::fread(&nt, 4, 1, fp) != 1 || std::fread(&force_Z, 8, 1, fp) != 1 || std::fread(&rho_in, 8, 1, fp) != 1) fail_("short checkpoint file");
    if (ok && (mflbm_set_parameter(ctx, "force_Z", force_Z) != MFLBM_OK || mflbm_set_parameter(ctx, "rho_in", rho_in) != MFLBM_OK)) fail_(mflbm_last_error(ctx));
    std::vector<double> tmp;
    auto get = [&](size_t slot, size_t n) {
        if (!ok) return;
        tmp.resize(n);
        if (std::fread(tmp.data(), 8, n, fp) != n) return fail_("short checkpoint file");
        mflbm_arrays a;
        std::memset(&a, 0, sizeof a);
        *(double **)((char *)&a + slot) = tmp.data();
        if (mflbm_upload(ctx, &a) != MFLBM_OK) fail_(mflbm_last_error(ctx));
    };
    const size_t n1 = (size_t)(nx + 2) * (ny + 2) * (nz + 2), n4 = (size_t)(nx + 8) * (ny + 8) * (nz + 8), np = (size_t)(nx + 2) * (ny + 2);
    for (int q = 0; q < 19; q++) get(offsetof(mflbm_arrays, f) + q * sizeof(double *), n1);
    if (c.multiphase) {
        for (int q = 0; q < 19; q++) get(offsetof(mflbm_arrays, g) + q * sizeof(double *), n1);
        get(offsetof(mflbm_arrays, phi), n4);
    }
    if (c.outlet_BC == 1) {
        get(offsetof(mflbm_arrays, f_convec_bc), np * 19);
        if (c.multiphase) {
            get(offsetof(mflbm_arrays, g_convec_bc), np * 19);
            get(offsetof(mflbm_arrays, phi_convec_bc), np);
        }
    }
    std::fclose(fp);
    if (ok && ntime0) *ntime0 = nt;
    return ok;
}

bool Driver::main_iteration_kernel(int ntime) {
    if (mflbm_step(ctx, ntime) != MFLBM_OK) { error = mflbm_last_error(ctx); return false; }
    return true;
}

bool Driver::color_gradient() {
    if (mflbm_color_gradient(ctx) != MFLBM_OK) { error = mflbm_last_error(ctx); return false; }
    return true;
}

bool Driver::cal_saturation(double *sat) {
    double v1 = 0, v2 = 0;
    if (mflbm_cal_saturation(ctx, &v1, &v2) != MFLBM_OK) { error = mflbm_last_error(ctx); return false; }
    *sat = v1 / (v1 + v2 + EPS_MP);  // MP/Monitor.F90:544 (np == 1)
    return true;
}

// host tail of monitor for np == 1 (MP/Monitor.F90:112-274, SP/Monitor.F90:66-172): reductions over the z profiles,
// text appended to out1.output/*.dat in the reference's formats when outdir is non-empty
bool Driver::monitor(int ntime, MonitorResult *o, const std::string &outdir) {
    if (c.npz != 1) { error = "Driver::monitor gathers np == 1 only; multi-slab callers reduce the tk buffers themselves"; return false; }
    const int nzG = c.nzGlobal;
    std::vector<double> tk(c.multiphase ? 7 * nz + 3 : 2 * nz + 1);
    if (mflbm_monitor(ctx, tk.data(), (int)tk.size()) != MFLBM_OK) { error = mflbm_last_error(ctx); return false; }
    *o = MonitorResult();
    const int k0 = c.n_exclude_inlet + 1, k1 = nzG - c.n_exclude_outlet;
    const bool open_z = c.kper == 0 && c.domain_wall_status_z_min == 0 && c.domain_wall_status_z_max == 0;
    auto app = [&](const char *name, const char *fmt, auto... args) {
        if (outdir.empty()) return;
        FILE *fp = std::fopen((outdir + "/" + name).c_str(), "a");
        if (!fp) return;
        std::fprintf(fp, fmt, args...);
        std::fclose(fp);
    };
    if (c.multiphase) {
        const double *fl1 = &tk[0], *fl2 = &tk[nz], *vol1 = &tk[2 * nz], *vol2 = &tk[3 * nz], *mass1 = &tk[4 * nz], *mass2 = &tk[5 * nz], *pre = &tk[6 * nz];
        o->umax_global = std::sqrt(tk[7 * nz]);
        o->kinetic_energy1 = 0.5 * tk[7 * nz + 1];
        o->kinetic_energy2 = 0.5 * tk[7 * nz + 2];
        for (int k = k0; k <= k1; k++) {
            o->mass1_sum += mass1[k - 1]; o->mass2_sum += mass2[k - 1];
            o->vol1_sum += vol1[k - 1]; o->vol2_sum += vol2[k - 1];
        }
        o->saturation = o->vol1_sum / (o->vol1_sum + o->vol2_sum);
        double t1 = 0, t2 = 0, t3 = 0, t4 = 0;
        for (int k = 1; k <= nzG; k++) { t1 += mass1[k - 1]; t2 += mass2[k - 1]; t3 += vol1[k - 1]; t4 += vol2[k - 1]; }
        o->saturation_full_domain = t3 / (t3 + t4);
        for (int k = 1; k <= nzG; k++) { o->fl1_avg_whole += fl1[k - 1]; o->fl2_avg_whole += fl2[k - 1]; }
        o->fl1_avg_whole /= (double)nzG; o->fl2_avg_whole /= (double)nzG;
        o->fl_avg_whole = o->fl1_avg_whole + o->fl2_avg_whole;
        for (int k = k0; k <= k1; k++) { o->fl1_avg += fl1[k - 1]; o->fl2_avg += fl2[k - 1]; }
        o->fl1_avg /= (double)(nzG - c.n_exclude_outlet - c.n_exclude_inlet);
        o->fl2_avg /= (double)(nzG - c.n_exclude_outlet - c.n_exclude_inlet);
        o->fl_avg = o->fl1_avg + o->fl2_avg;
        o->ca = (o->fl_avg / A_xy) * c.la_nu1 / c.gamma;
        app("saturation.dat", "%10d %14.7e %14.7e %14.7e %14.7e %14.7e\n", ntime, o->saturation, o->vol1_sum, o->vol2_sum, o->mass1_sum, o->mass2_sum);
        app("Ca_number.dat", "%10d %14.7e %14.7e %14.7e %14.7e\n", ntime, o->ca, o->umax_global, o->kinetic_energy1, o->kinetic_energy2);
        app("flowrate_time.dat", "%10d %14.6E %14.6E %14.6E %14.6E %14.6E %14.6E\n", ntime, o->fl_avg_whole, o->fl1_avg_whole, o->fl2_avg_whole, o->fl_avg, o->fl1_avg, o->fl2_avg);
        app("saturation_full_domain.dat", "%10d %14.7e %14.7e %14.7e %14.7e %14.7e\n", ntime, o->saturation_full_domain, t3, t4, t1, t2);
        if (open_z) {
            o->pre_in = pre[k0 - 1] / pore_profile_z[k0 - wk0];
            o->pre_out = pre[k1 - 1] / pore_profile_z[k1 - wk0];
            app("pre.dat", "%10d %14.7e %14.7e %14.7e\n", ntime, o->pre_in, o->pre_out, o->pre_in - o->pre_out);
        }
        if (std::isnan(o->saturation_full_domain) || std::isnan(o->ca)) o->simulation_end_indicator = 3;  // :253-259
        else if (o->umax_global > 0.5) o->simulation_end_indicator = 3;
    } else {
        const double *fl = &tk[0], *pre = &tk[nz];
        o->umax_global = std::sqrt(tk[2 * nz]);
        for (int k = 1; k <= nzG; k++) o->fl_avg_whole += fl[k - 1];
        o->fl_avg_whole /= (double)nzG;
        o->flowrate = o->fl_avg_whole;
        app("flowrate_time.dat", "%10d %14.6E %14.6E\n", ntime, o->fl_avg_whole, o->umax_global);
        if (open_z) {
            o->pre_in = pre[k0 - 1] / pore_profile_z[k0 - wk0];
            o->pre_out = pre[k1 - 1] / pore_profile_z[k1 - wk0];
            app("pre.dat", "%10d %14.7e %14.7e %14.7e\n", ntime, o->pre_in, o->pre_out, o->pre_in - o->pre_out);
        }
        if (std::isnan(o->fl_avg_whole) || o->umax_global > 0.5) o->simulation_end_indicator = 3;
    }
    return true;
}

// MP/Main_multiphase.F90:498-556: warm-up, then rounds of ntime_max_benchmark steps; MLUPS over pore nodes.
// Timed on the device (CUDA events) instead of system_clock.
bool Driver::benchmark(int warmup, int rounds, int steps, double *best_mlups, double *ms_per_step) {
    if (mflbm_run(ctx, 1, warmup) != MFLBM_OK || mflbm_sync(ctx) != MFLBM_OK) { error = mflbm_last_error(ctx); return false; }
    double best = 1e300;
    for (int r = 0; r < rounds; r++) {
        double ms = 0;
        if (mflbm_timer_start(ctx) != MFLBM_OK || mflbm_run(ctx, 1, steps) != MFLBM_OK || mflbm_timer_stop(ctx, &ms) != MFLBM_OK) {
            error = mflbm_last_error(ctx);
            return false;
        }
        best = std::min(best, ms);
    }
    *ms_per_step = best / steps;
    *best_mlups = (double)pore_sum * steps / (best * 1e-3) / 1e6;
    return true;
}

}  // namespace mflbm_host
