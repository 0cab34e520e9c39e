// mflbm_main.cpp -- stand-in for "program main_multiphase" / "program main" (MP/Main_multiphase.F90:20-330, SP/Main.F90)
// for boxes without a Fortran toolchain: reads ./simulation_control.txt and ./path_info.txt like the reference, runs
// the benchmark (benchmark_cmd = 1) or the main loop with monitor / breakthrough cadence, and writes the same
// out1.output/*.dat text files and job_status.txt states.  One slab only (MPI_process_num 1,1,1); multi-GPU runs go
// through bench.py / torch.distributed.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <sys/stat.h>

#include "mflbm_driver.hpp"

using namespace mflbm_host;

static int die(Driver &d, const char *what) {
    std::fprintf(stderr, " %s: %s\n", what, d.error.c_str());
    std::ofstream("job_status.txt") << "simulation_failed\n";
    return 1;
}

int main(int argc, char **argv) {
    Driver d;
    const std::string control = argc > 1 ? argv[1] : "./simulation_control.txt";
    if (!d.read_parameter(control)) return die(d, "read_parameter");
    if (d.c.npz != 1) { d.error = "mflbm_run drives one slab; use bench.py for multi-GPU"; return die(d, "MPI_process_num"); }
    if (d.c.external_geometry_read_cmd == 1) {  // MP/Misc.F90:47-71: path_info.txt line 4 = geometry file
        std::ifstream pi("./path_info.txt");
        std::string line, geo;
        for (int n = 0; n < 4 && std::getline(pi, line); n++) geo = line;
        if (!pi) { d.error = "Error! path_info.txt is not found! Exiting program!"; return die(d, "set_walls"); }
        if (!d.read_walls(geo)) return die(d, "read_walls");
    }
    std::printf(" ***************************** Initialization **********************************\n");
    d.set_walls();
    if (d.c.multiphase) d.geometry_preprocessing_new();
    d.initialization_basic();
    d.initialization_new();
    std::printf(" Total number of pore nodes = %14lld\n", d.pore_sum);
    std::printf(" Inlet open cross sectional area = %14.2f\n", d.A_xy);
    if (!d.create_context(-1, 0, nullptr, 0) || !d.upload()) return die(d, "device");
    mkdir("out1.output", 0755);
    if (d.c.multiphase && !d.color_gradient()) return die(d, "color_gradient");
    if (d.c.benchmark_cmd == 1) {
        std::printf(" ********************** Performance benchmarking *******************************\n");
        double mlups = 0, ms = 0;
        if (!d.benchmark(20, 3, d.c.ntime_max_benchmark, &mlups, &ms)) return die(d, "benchmark");
        std::printf(" Code performance: %12.4f MLUPS\n", mlups);
        FILE *fp = std::fopen("out1.output/benchmark_time.dat", "a");
        if (fp) { std::fprintf(fp, "Code performance: %12.4f MLUPS\n", mlups); std::fclose(fp); }
        MonitorResult m;
        if (!d.monitor(d.c.ntime_max_benchmark, &m, "out1.output")) return die(d, "monitor");
        std::printf(" saturation %14.6E  capillary number %14.6E\n", m.saturation, m.ca);
        std::ofstream("job_status.txt") << "simulation_done\n";
        return 0;
    }
    std::printf(" ************************** Entering main loop *********************************\n");
    int end_indicator = 0;
    long long ntime = 1;
    for (; ntime <= d.c.ntime_max; ntime++) {
        if (!d.main_iteration_kernel((int)ntime)) return die(d, "main_iteration_kernel");
        if (ntime % d.c.ntime_monitor == 0) {
            MonitorResult m;
            if (!d.monitor((int)ntime, &m, "out1.output")) return die(d, "monitor");
            end_indicator = m.simulation_end_indicator;
            if (d.c.multiphase && d.c.breakthrough_check == 1) {
                int32_t cnt = 0;
                if (mflbm_monitor_breakthrough(d.ctx, &cnt) != MFLBM_OK) return die(d, "monitor_breakthrough");
                if (cnt >= 1) { std::printf(" Breakthrough point reached! Exiting program!\n"); end_indicator = 1; }
            }
            if (ntime % d.c.ntime_display_steps == 0)
                std::printf(" ntime = %lld  saturation = %.6f  Ca = %.4e  umax = %.4e\n", ntime, m.saturation_full_domain, m.ca, m.umax_global);
        }
        if (end_indicator != 0) break;
    }
    const char *st = end_indicator == 3 ? "simulation_failed" : (end_indicator == 1 ? "simulation_done" : "simulation_reached_max_step");
    std::ofstream("job_status.txt") << st << "\n";
    std::printf(" Simulation ended: %s after %lld steps\n", st, ntime);
    return end_indicator == 3 ? 2 : 0;
}
