#!/usr/bin/env python
"""bench.py -- MLUPS of the MF-LBM time-step hot path on B200 (contract: see the task statement / DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c1] [--impl reference]
  torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, z-slab decomposition, weak scaling)

One "step" = one call of main_iteration_kernel (mflbm_step): collision+streaming of every fluid node, halo exchange,
inlet/outlet kernels and colour gradient.  MLUPS counts fluid (pore) nodes only, like the reference's benchmark
(MP/Main_multiphase.F90:540).  Default workload for every N: BASELINE.json configs[4], one 1536x1536x192 multiphase
sphere-pack slab per GPU (the largest single-GPU configuration, 97 GB; N = 8 is the 1536^3 weak-scaling target); the
slabs are identical packs stacked along z, so every GPU holds the same number of fluid nodes.  ``--workload c3`` is
configs[2] (512^3 per GPU), ``c4`` configs[3] (strong scaling 512x512x1024), ``c2`` / ``c2rock`` configs[1] (singlephase
240x240x260, synthetic / the reference's Bentheimer rock).  After the headline region (drainage front next to the
inlet) the multiphase workloads are timed again in two interface-rich states (``roofline_active``); N>1 first checks
N slabs == single-domain oracle bit for bit on small cases (``parity_ngpu``).

The GPU arm never touches oracle/: geometry, node lists and initial fields come from the host driver mirror
(mf-lbm_b200/host), everything per step from libmflbm.so through the C ABI.  The CPU oracle is only executed for
the ``cpu_baseline`` figure and for ``--impl reference``.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_UPDATE = {True: 624.0, False: 304.0}  # (38+38) PDFs + phi write + phi read ; (19+19) PDFs


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_spec(name, n_gpus):
    """Control-file overrides + geometry recipe for each BASELINE.json config."""
    if name == "c3":  # configs[2] (N=1) / weak-scaling stack (N>1), SURVEY 8(d) C3/C5 physics
        n = int(os.environ.get("MFLBM_BENCH_N", "512"))
        por = float(os.environ.get("MFLBM_BENCH_POROSITY", "0.36"))  # developer knob; the BASELINE config is 0.36
        return dict(multiphase=True, nx=n, ny=n, nz=n * n_gpus, periodic=False, unit_nz=n,
                    geometry=dict(porosity=por, rmin=8.0, rmax=20.0, seed=1, buffer=10),
                    control=dict(fluid1_viscosity=0.004, fluid2_viscosity=0.04, surface_tension=0.03, theta=30,
                                 RK_beta=0.95, inlet_BC=1, outlet_BC=1, capillary_number="100d-6",
                                 initial_interface_position=8.0, initial_fluid_distribution_option=1,
                                 excluded_layers="10,10"),
                    label="multiphase_3D scCO2/brine drainage, synthetic random-sphere-pack %dx%dx%d (porosity 0.36, "
                          "10 buffer layers each end%s), velocity inlet / convective outlet" % (
                              n, n, n * n_gpus, "" if n_gpus == 1 else "; %d identical %d-plane packs stacked along z" % (n_gpus, n)))
    if name == "c4":  # configs[3]: STRONG scaling, the 512x512x1024 lattice (flow axis z) cut into n_gpus slabs
        if 1024 % n_gpus:
            raise SystemExit("workload c4 needs a GPU count that divides 1024")
        return dict(multiphase=True, nx=512, ny=512, nz=1024, periodic=False, scaling="strong",
                    geometry=dict(porosity=0.36, rmin=8.0, rmax=20.0, seed=2, buffer=10),
                    control=dict(fluid1_viscosity=0.004, fluid2_viscosity=0.04, surface_tension=0.03, theta=30,
                                 RK_beta=0.95, inlet_BC=1, outlet_BC=1, capillary_number="100d-6",
                                 initial_interface_position=8.0, initial_fluid_distribution_option=1,
                                 excluded_layers="10,10"),
                    label="multiphase_3D strong scaling, synthetic sphere pack 512x512x1024 (porosity 0.36, 10 buffer layers "
                          "each end) in %d z-slab(s) of %d planes, velocity inlet / convective outlet" % (n_gpus, 1024 // n_gpus))
    if name == "c5":  # configs[4]: one 1536x1536x192 slab per GPU (N = 8 is the 1536^3 weak-scaling target), same physics as C3
        return dict(multiphase=True, nx=1536, ny=1536, nz=192 * n_gpus, periodic=False, unit_nz=192,
                    geometry=dict(porosity=0.36, rmin=8.0, rmax=20.0, seed=3, buffer=10),
                    control=dict(fluid1_viscosity=0.004, fluid2_viscosity=0.04, surface_tension=0.03, theta=30,
                                 RK_beta=0.95, inlet_BC=1, outlet_BC=1, capillary_number="100d-6",
                                 initial_interface_position=8.0, initial_fluid_distribution_option=1,
                                 excluded_layers="10,10"),
                    label="multiphase_3D weak scaling towards 1536^3: 1536x1536x%d = %d identical 192-plane sphere packs stacked "
                          "along z (porosity 0.36 core + 10 fluid layers at each end of every pack, so each GPU holds the same "
                          "number of fluid nodes), velocity inlet / convective outlet" % (192 * n_gpus, n_gpus))
    if name == "c2":  # configs[1]
        n = int(os.environ.get("MFLBM_BENCH_N", "240"))
        return dict(multiphase=False, nx=n, ny=n, nz=(n + 20) * n_gpus, periodic=True,
                    geometry=dict(porosity=0.18, rmin=6.0, rmax=14.0, seed=20261017, buffer=10),
                    control=dict(fluid_viscosity=0.1, body_force_0="1d-5", MRT_collision_parameter_preset=1,
                                 periodic_indicator="0,0,1", excluded_layers="10,10"),
                    label="singlephase_3D absolute-permeability run, Bentheimer-like synthetic %dx%dx%d wall array "
                          "(porosity 0.18 core + 10 fluid layers each end), periodic z, body force" % (n, n, (n + 20) * n_gpus))
    if name == "c2rock":  # configs[1] on the reference's own Bentheimer wall array (its benchmark cases 6 / 8); one GPU only
        if n_gpus != 1:
            raise SystemExit("workload c2rock is a single-GPU case")
        return dict(multiphase=False, nx=240, ny=240, nz=260, periodic=True, geometry="rock",
                    walls_file=os.path.join(ROOT, "tests", "golden", "bentheimer_in10_240_out10.bits.xz"),
                    control=dict(fluid_viscosity=0.1, body_force_0="1d-5", MRT_collision_parameter_preset=1,
                                 periodic_indicator="0,0,1", excluded_layers="10,10"),
                    label="singlephase_3D absolute-permeability run on the reference's Bentheimer wall array 240x240x260 "
                          "(bentheimer_in10_240_240_240_out10.dat, 3 670 813 pore nodes), periodic z, body force")
    if name == "c1":  # configs[0]: the reference's own CPU-runnable case
        return dict(multiphase=True, nx=40, ny=40, nz=60 * n_gpus, periodic=False, geometry=None,
                    control=dict(modify_geometry_cmd=1, breakthrough_check=1),
                    label="test_suites/3D_simulation tube-with-sphere drainage 40x40x%d" % (60 * n_gpus))
    raise SystemExit("unknown workload " + name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax = float(p[2])
                power.append(float(p[3]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm)}


def cpu_reference_run(spec, steps, warmup, sample_n=None):
    """The reference's CPU implementation of the path = the C/OpenMP oracle (the Fortran cannot be built in this image:
    no gfortran / mpif90), timing build, all host threads, on a bounded sample of the same workload."""
    from oracle.oracle import Oracle, default_params
    import mflbm_b200 as M
    from importlib import import_module
    geo = import_module("mflbm_b200.geometry")
    mp = spec["multiphase"]
    cores = os.cpu_count() or 1
    # all host cores: under torchrun only rank 0 runs this leg, and torchrun's OMP_NUM_THREADS=1 default must not stick
    # (the OpenMP runtime of the oracle library reads the variable when it is loaded, below)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    if spec["geometry"] is None:
        n, nz = spec["nx"], spec["nz"]
        walls = None
    else:
        n = sample_n or (128 if mp else 160)
        nz = n
        if spec["geometry"] == "rock":  # a central crop of the reference's rock
            full = geo.load_packed_walls(spec["walls_file"], (spec["nx"], spec["ny"], spec["nz"]))
            i0, k0 = (spec["nx"] - n) // 2, (spec["nz"] - nz) // 2
            walls = np.ascontiguousarray(full[i0:i0 + n, i0:i0 + n, k0:k0 + nz])
        else:
            walls = geo.sphere_pack(n, n, nz, periodic=spec["periodic"], **spec["geometry"])
    c = spec["control"]
    fl = lambda v: float(str(v).replace("d", "e"))
    kw = dict(multiphase=1 if mp else 0, nxG=n, nyG=n, nzG=nz)
    if mp:
        kw.update(la_nu1=fl(c.get("fluid1_viscosity", 0.004)), la_nu2=fl(c.get("fluid2_viscosity", 0.4)), gamma=fl(c.get("surface_tension", 0.03)),
                  theta_deg=fl(c.get("theta", 30)), beta=fl(c.get("RK_beta", 0.95)), ca_0=fl(c.get("capillary_number", "100d-6")),
                  modify_geometry_cmd=int(c.get("modify_geometry_cmd", 0)))
    else:
        kw.update(la_nu1=fl(c["fluid_viscosity"]), force_z0=fl(c["body_force_0"]), kper=1, inlet_BC=0, outlet_BC=0,
                  n_exclude_inlet=10, n_exclude_outlet=10)
    o = Oracle(default_params(**kw), fast=True)
    o.setup(walls)
    pore = o.get_i64("pore_sum")
    # Which code is timed.  "reference": the reference's OWN subroutines (main_iteration_kernel and everything it calls),
    # translated statement by statement from its Fortran sources by oracle/f2c_lite.py and compiled with gcc -O3 -fopenmp
    # (oracle/_ref/*_fast.so, built where /root/reference is mounted; its !$omp parallel do directives carried over) --
    # available for the open-z workloads (the periodic wrap of the reference is MPI self-exchange, which has no translation).
    # "port": the hand-written C/OpenMP restatement (oracle/mflbm_oracle.c), bit-identical to the former (tests/test_ref_pin.py).
    kind, step = "port", o.step
    if not spec["periodic"]:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from oracle import ref as R
            if os.path.exists(R.lib_path("mp" if mp else "sp", fast=True)):
                from ref_helpers import copy_state, ref_from_oracle
                r = ref_from_oracle(o, fast=True)
                copy_state(r, o)

                def step(t):
                    r.set(ntime=t)
                    r.call("main_iteration_kernel")
                if mp:
                    r.call("color_gradient")
                kind = "reference"
        except Exception as e:  # the port is always there
            sys.stderr.write("bench.py: oracle/_ref unavailable (%s), timing the port\n" % e)
            kind, step = "port", o.step
    if mp and kind == "port":
        o.color_gradient()
    for t in range(1, warmup + 1):
        step(t)
    t0 = time.perf_counter()
    for t in range(warmup + 1, warmup + steps + 1):
        step(t)
    dt = time.perf_counter() - t0
    mlups = pore * steps / dt / 1e6
    what = ("the reference's own Fortran subroutines translated to C (oracle/f2c_lite.py), gcc -O3" if kind == "reference"
            else "C/OpenMP port of the reference (oracle/mflbm_oracle.c), gcc -O3")
    return dict(value=mlups, ms_per_step=dt / steps * 1e3, cores=cores, pore=int(pore), kind=kind,
                sample="%s %dx%dx%d sample of the same medium/physics, %d timed steps, OpenMP %d threads; %s" % (
                    "multiphase" if mp else "singlephase", n, n, nz, steps, cores, what))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="c5", help="c5 (default: one 1536x1536x192 slab per GPU, BASELINE configs[4] and the largest "
                    "single-GPU configuration), c3 (512^3 per GPU, configs[2]), c4 (strong scaling 512x512x1024, configs[3]), "
                    "c2 / c2rock (singlephase 240x240x260, configs[1]), c1 (tube+sphere, configs[0])")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--variant", type=int, default=0, help="kernel_variant: 0 auto, 1 dense, 2 sparse")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--host-geometry", action="store_true",
                    help="run geometry_preprocessing_new on the host cores (default: on this rank's GPU, mflbm_geometry_preprocess)")
    ap.add_argument("--no-e2e", action="store_true", help="developer sweeps: skip the end-to-end leg (the line then has e2e = null)")
    ap.add_argument("--e2e-blocking", action="store_true", help="e2e leg with mflbm_upload + mflbm_step + mflbm_cal_saturation (three "
                    "blocking calls per step) instead of mflbm_step_streamed")
    ap.add_argument("--state", default="drainage", choices=["drainage", "random"],
                    help="developer / profiling: 'random' re-initialises with the seeded option-6 phase field BEFORE the headline region")
    ap.add_argument("--no-active", action="store_true", help="skip the interface-rich second timed regions (roofline_active = null)")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the N-slab == single-domain parity check before the timed region")
    args = ap.parse_args()
    if args.steps % 2:
        args.steps += 1  # AA pattern: keep odd/even parity across rounds like the reference (MP/Main_multiphase.F90:510)
    if args.warmup % 2:
        args.warmup += 1
    args.warmup = max(args.warmup, 4)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # torchrun pins OMP_NUM_THREADS=1; the (untimed) host-side setup -- geometry, node lists, adjacency -- is OpenMP
        # code in libmflbm.so / libmflbm_host.so, loaded below: give every rank its share of the host cores
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = world if world > 1 else 1
    spec = workload_spec(args.workload, n_gpus)
    mp = spec["multiphase"]
    bpu = BYTES_PER_UPDATE[mp]
    peak, peak_src = measured_peaks()

    # ------------------------------------------------------------------ reference arm (CPU) ----------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(workload_spec(args.workload, 1), max(2, min(args.steps, 20)), max(2, min(args.warmup, 4)))
        print(json.dumps({
            "impl": "reference", "metric": "MLUPS", "value": r["value"], "unit": "MLUPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": spec["label"], "note": "no Fortran compiler in this image (no gfortran/mpif90): this arm times the "
                       "reference's CPU path on a bounded sample of the workload -- kind 'reference' = its own subroutines "
                       "translated mechanically to C (oracle/_ref), kind 'port' = the hand-written C/OpenMP restatement"},
            "cpu_baseline": {"value": r["value"], "unit": "MLUPS", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ------------------------------------------------------------------ B200 arm -------------------------------
    import torch
    import mflbm_b200 as M
    from importlib import import_module
    geo = import_module("mflbm_b200.geometry")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback; use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    rk = import_module("mflbm_b200.dist").Ranks(backend="nccl")  # rendezvous, scalar reductions, NCCL-id broadcast
    dist = rk.dist
    allreduce = rk.allreduce

    parity = None
    if world > 1 and not args.no_parity:
        parity = parity_ngpu(rk, M, local_rank)  # N slabs == single-domain oracle, checked before anything is timed

    nx, ny, nzG = spec["nx"], spec["ny"], spec["nz"]
    nz = nzG // n_gpus
    t_setup = time.time()
    tmp = tempfile.mkdtemp(prefix="mflbm_bench_")
    ctl = M.write_control_file(os.path.join(tmp, "simulation_control.txt"), multiphase=mp,
                               lattice_dimensions="%d,%d,%d" % (nx, ny, nzG), MPI_process_num="1,1,%d" % n_gpus,
                               MPI_async_layers_num="0,0,4", external_geometry_read_cmd=0 if spec["geometry"] is None else 1,
                               **spec["control"])
    if spec["geometry"] is None:
        drv = M.Driver(ctl, idz=rank, lazy_pdfs=True, device_geometry=None if args.host_geometry else local_rank)
    else:
        k0, k1 = M.Driver.window_range(rank, n_gpus, nzG, spec["periodic"]) if n_gpus > 1 else (1, nzG)
        if spec["geometry"] == "rock":
            w = geo.load_packed_walls(spec["walls_file"], (nx, ny, nzG))
        elif spec.get("unit_nz") and n_gpus > 1:  # weak scaling: every GPU holds the same pack (equal fluid-node counts)
            w = geo.stacked_window(nx, ny, spec["unit_nz"], k0, k1, **spec["geometry"])
        else:
            w = geo.sphere_pack_window(nx, ny, nzG, k0, k1, periodic=spec["periodic"], **spec["geometry"])
        drv = M.Driver(ctl, idz=rank, walls_window=(w, k0), lazy_pdfs=True, device_geometry=None if args.host_geometry else local_rank)
        del w
    drv.setup()
    pore_local = drv.i64("pore_sum_local")
    pore_global = int(round(allreduce(float(pore_local))))
    drv.set_pore_sum(pore_global)
    nccl_id = None
    if world > 1:
        nccl_id = rk.broadcast_bytes(M.nccl_unique_id() if rank == 0 else b"", 128)
    drv.create_context(device=local_rank, nccl_unique_id=nccl_id, kernel_variant=args.variant)
    drv.upload(free_host=True)
    if mp:
        drv.color_gradient()
    drv.sync()
    setup_s = time.time() - t_setup

    def barrier():
        drv.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def timed_region(steps, warmup, clocks=False):
        """warm-up, then EXACTLY `steps` steps device-timed with CUDA events on the compute stream (max over ranks), the
        collision-kernel launches event-timed on their own stream inside it"""
        drv.run(1, warmup)
        barrier()
        launches0 = drv.launch_count
        spec0 = drv.spec_steps
        sampler = ClockSampler(local_rank) if (clocks and rank == 0) else None
        if sampler:
            sampler.start()
        drv.profile(True)
        barrier()
        spec0 = drv.spec_steps
        drv.timer_start()
        drv.run(1, steps)
        ms_local = drv.timer_stop()
        barrier()
        coll_ms, coll_launches = drv.profile_read()
        drv.profile(False)
        r = dict(ms_local=ms_local, ms=allreduce(ms_local, "max"), coll_ms=coll_ms, coll_launches=coll_launches,
                 launches=drv.launch_count - launches0, clocks=sampler.stop() if sampler else None, spec=drv.spec_steps - spec0)
        nt, nq = drv.tile_stats()
        r["quiet"] = (nq / nt) if nt else None
        r["mlups"] = pore_global * steps / (r["ms"] * 1e-3) / 1e6
        r["achieved"] = bpu * pore_local * steps / (coll_ms * 1e-3) / 1e9 if coll_ms > 0 else 0.0
        r["step_frac"] = r["mlups"] * 1e6 * bpu / 1e9 / (peak * n_gpus)
        return r

    if mp and args.state == "random":
        drv.reinitialize(6, seed=20261018)
        drv.color_gradient()
        drv.sync()
    main_r = timed_region(args.steps, args.warmup, clocks=True)
    ms, mlups, achieved = main_r["ms"], main_r["mlups"], main_r["achieved"]
    per_rank_ms = rk.allgather(main_r["ms_local"] / args.steps)
    per_rank_pore = rk.allgather(float(pore_local))

    # ncu DRAM bytes per k_collide launch: only when a capture of exactly this workload / lattice / GPU count was kept
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("%s/%dx%dx%d/n%d" % (args.workload, nx, ny, nz, n_gpus), {}).get("dram_bytes_per_launch")

    # e2e through the C ABI with HOST buffers: per step H2D of the step's host input (inlet profile w_in, pinned),
    # mflbm_step, and a D2H read of the step's result (saturation partial sums via mflbm_cal_saturation / flow monitor)
    w_in = torch.zeros((ny + 2) * (nx + 2), dtype=torch.float64).pin_memory()
    w_in.copy_(torch.from_numpy(np.ascontiguousarray(drv.field("w_in").ravel(order="F"))))
    e2e_steps = 0 if args.no_e2e else args.steps
    barrier()
    t0 = time.perf_counter()
    drv.timer_start()
    acc = 0.0
    if args.e2e_blocking:  # round-1 form: three blocking calls per step
        for t in range(1, e2e_steps + 1):
            drv.upload_w_in(w_in.data_ptr())
            drv.main_iteration_kernel(t)
            if mp:
                v1, v2 = drv.cal_saturation_parts()
                acc += v1
            else:
                drv.sync()
    else:
        # streamed steps (include/mflbm.h): every step still takes its w_in from pinned host memory and has its saturation sums
        # read back to the host, but the copy runs beside the previous step and the result is handed out one call later
        for t in range(1, e2e_steps + 1):
            r = drv.step_streamed(t, w_in.data_ptr())
            if r is not None:
                acc += r[0]
        if e2e_steps:
            acc += drv.stream_flush()[0]
    ms_e2e = drv.timer_stop()
    barrier()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = allreduce(max(ms_e2e, 0.0), "max")
    e2e_mlups = pore_global * e2e_steps / (ms_e2e * 1e-3) / 1e6 if e2e_steps else None
    h2d = (nx + 2) * (ny + 2) * 8
    # partial sums of cal_saturation copied back per step (kernels_monitor.cu launch_saturation: two rows of block sums)
    red_len = 16 * max(nz, ny) + 64
    d2h = 2 * 8 * max(1, min(148 * 8, red_len // 2, (int(pore_local) + 255) // 256)) if mp else 0

    # interface-rich regimes (the timed state above is a drainage front next to the inlet: most tiles are quiet)
    active = None
    if mp and not args.no_active:
        ksteps = max(2, (args.steps // 2) & ~1)
        # (i) same state, quiet-tile skipping switched off: the colour-gradient chain runs on every node
        drv.set_parameter("quiet_tiles", 0)
        drv.color_gradient()
        r1 = timed_region(ksteps, 4)
        drv.set_parameter("quiet_tiles", 1)
        # (ii) the reference's own benchmark state (test_suites/3D_simulation/6.performance_benchmarking:
        # initial_fluid_distribution_option 6, per-node random phi at target_fluid1_saturation 0.4), seeded here
        t_re = time.time()
        drv.reinitialize(6, seed=20261018)
        drv.color_gradient()
        drv.sync()
        t_re = time.time() - t_re
        r2 = timed_region(args.steps, args.warmup)
        active = {
            "no_quiet_tiles": {"ms_per_step": r1["ms"] / ksteps, "mlups": r1["mlups"], "frac": r1["achieved"] / peak,
                               "step_frac": r1["step_frac"], "kernel_ms_per_step": r1["coll_ms"] / ksteps, "steps": ksteps,
                               "state": "the drainage state of the headline region, quiet-tile skipping off (mflbm_set_parameter quiet_tiles 0)"},
            "ms_per_step": r2["ms"] / args.steps, "mlups": r2["mlups"], "frac": r2["achieved"] / peak, "step_frac": r2["step_frac"],
            "kernel_ms_per_step": r2["coll_ms"] / args.steps, "quiet_tile_fraction": r2["quiet"], "steps": args.steps,
            "warmup": args.warmup, "reinit_s": round(t_re, 1), "gpu_launches": int(r2["launches"]),
            "state": "initial_fluid_distribution_option 6: per-node random phi = +-1 at saturation 0.4 (seeded hash of the global "
                     "node position), the state of the reference's benchmark case 6; interface everywhere, no quiet tile"}

    out = None
    if rank == 0:
        out = {
            "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": spec.get("scaling", "weak"), "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": spec["label"], "workload_key": args.workload, "fluid_nodes": pore_global,
                       "porosity": pore_global / float(nx * ny * nzG),
                       "per_gpu_lattice": "%dx%dx%d" % (nx, ny, nz), "parallelism": "z-slab x%d" % n_gpus,
                       "l2_policy": "working set per step (%.1f GB) >> 126 MB L2, no flush needed" % (drv.device_bytes / 1e9),
                       "population_layout": "auto (kernel_variant=%d)" % args.variant, "setup_s": round(setup_s, 1),
                       "device_bytes_per_gpu": drv.device_bytes,
                       "quiet_tile_fraction": main_r["quiet"], "state": args.state,
                       "speculative_chain_steps": int(main_r["spec"]), "tile_summary": drv.spec_info(),
                       "fluid_nodes_per_rank": [int(v) for v in per_rank_pore],
                       "ms_per_step_per_rank": [round(v, 4) for v in per_rank_ms]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_collide (collision + AA streaming)", "peak_source": peak_src,
                         "bytes_per_update": bpu, "kernel_ms_per_step": main_r["coll_ms"] / args.steps,
                         "kernel_launches": main_r["coll_launches"], "step_frac_of_roofline": main_r["step_frac"]},
            "roofline_active": active,
            "e2e": None if args.no_e2e else {"value": e2e_mlups, "unit": "MLUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / max(e2e_steps, 1), "host_wall_ms_per_step": wall_e2e / max(e2e_steps, 1),
                    "what": ("per step: mflbm_upload(w_in, pinned host) + mflbm_step + mflbm_cal_saturation read-back (blocking calls)" if args.e2e_blocking
                             else "per step: mflbm_step_streamed(w_in from pinned host memory, copied beside the previous step; saturation "
                                  "sums of every step read back to the host, handed out one call later)")},
            "gpu_launches": int(main_r["launches"]), "clocks": main_r["clocks"], "parity_ngpu": parity}
    drv.close()
    if rank == 0:
        if n_gpus == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(workload_spec(args.workload, 1), 10 if mp else 20, 2)
            out["cpu_baseline"] = {"value": r["value"], "unit": "MLUPS", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out))
    rk.close()
    return 0


def parity_ngpu(rk, M, local_rank):
    """N-slab == single-domain parity on this job's N GPUs (the cases of tests/test_multi_gpu.py, stretched to 16 planes
    per slab): every rank runs one z slab of a small lattice with the STRICT (-fmad=false) library and the NCCL halo
    exchange, and compares the populations, phi and the interface normal of its slab bit for bit with a single-domain
    run of the CPU oracle.  Checker leg: the oracle is test infrastructure and is never timed or shipped."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import ctx_from_oracle, make_oracle
    npz, idz = rk.world, rk.rank
    nzG = 16 * npz
    cases = {
        "mp_open": dict(nxG=20, nyG=18, nzG=nzG, la_nu2=0.04, interface_z0=6.0, n_exclude_inlet=0, n_exclude_outlet=0),
        "mp_periodic": dict(nxG=20, nyG=18, nzG=nzG, kper=1, force_z0=2e-4, la_nu2=0.04, initial_fluid_distribution_option=5,
                            interface_z0=6.0, n_exclude_inlet=0, n_exclude_outlet=0),
        "sp_periodic": dict(multiphase=0, nxG=20, nyG=18, nzG=nzG, kper=1, force_z0=1e-5, la_nu1=0.1, n_exclude_inlet=0,
                            n_exclude_outlet=0),
        "mp_yz_periodic": dict(nxG=20, nyG=18, nzG=nzG, jper=1, kper=1, wsy0=0, wsy1=0, inlet_BC=0, outlet_BC=0, force_z0=2e-4,
                               la_nu2=0.04, initial_fluid_distribution_option=3, interface_z0=7.0, n_exclude_inlet=0, n_exclude_outlet=0),
        "mp_y_periodic_open": dict(nxG=20, nyG=18, nzG=nzG, jper=1, wsy0=0, wsy1=0, la_nu2=0.04, interface_z0=6.0, n_exclude_inlet=0,
                                   n_exclude_outlet=0),
    }
    steps = 9
    rng = np.random.default_rng(3)
    wg = (rng.random((20, 18, nzG)) < 0.25).astype(np.int8)
    wg[:, :, :3] = 0
    wg[:, :, -3:] = 0
    bad_total, names = 0, []
    for name in sorted(cases):
        for layout in ((2,) if cases[name].get("jper") else (2, 1)):  # y-periodic lattices always run the sparse layout
            ref = make_oracle(walls_global=wg, **cases[name])  # single domain
            if ref.mp:
                ref.color_gradient()
            for t in range(1, steps + 1):
                ref.step(t)
            o = make_oracle(walls_global=wg, npz=npz, idz=idz, **cases[name])  # this rank's slab: initial state only
            nid = rk.broadcast_bytes(M.nccl_unique_id() if idz == 0 else b"", 128)
            ctx = ctx_from_oracle(o, strict=True, kernel_variant=layout, device=local_rank, use_nccl=1, nccl_unique_id=nid)
            if o.mp:
                ctx.color_gradient()
            ctx.run(1, steps)
            ctx.sync()
            got = ctx.download(*(["f"] + (["g", "phi", "cn_x", "cn_y", "cn_z", "c_norm"] if o.mp else [])))
            nzl = nzG // npz
            ks = slice(idz * nzl, (idz + 1) * nzl)
            fluid = (ref.walls[2:-2, 2:-2, 2:-2] == 0)[:, :, ks]
            bad = 0
            for q in range(19):
                bad += int(np.count_nonzero(got["f"][q][1:-1, 1:-1, 1:-1][fluid] != ref.f(q)[1:-1, 1:-1, 1:-1][:, :, ks][fluid]))
                if o.mp:
                    bad += int(np.count_nonzero(got["g"][q][1:-1, 1:-1, 1:-1][fluid] != ref.g(q)[1:-1, 1:-1, 1:-1][:, :, ks][fluid]))
            if o.mp:
                bad += int(np.count_nonzero(got["phi"][4:-4, 4:-4, 4:-4][fluid] != ref.field("phi")[4:-4, 4:-4, 4:-4][:, :, ks][fluid]))
                for nm in ("cn_x", "cn_y", "cn_z", "c_norm"):
                    bad += int(np.count_nonzero(got[nm][2:-2, 2:-2, 2:-2][fluid] != ref.field(nm)[2:-2, 2:-2, 2:-2][:, :, ks][fluid]))
            ctx.close()
            o.close()
            ref.close()
            bad = int(round(rk.allreduce(float(bad))))
            bad_total += bad
            names.append("%s/%s:%s" % (name, "sparse" if layout == 2 else "dense", "ok" if bad == 0 else "%d mismatches" % bad))
    return {"cases": names, "slabs": npz, "steps": steps, "lattice": "20x18x%d" % nzG, "bit_exact": bad_total == 0,
            "mismatched_values": bad_total, "library": "libmflbm_strict.so (-fmad=false)",
            "reference": "single-domain CPU oracle (oracle/, checker only)"}


if __name__ == "__main__":
    sys.exit(main())
