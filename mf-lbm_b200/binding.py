"""ctypes binding of the C ABI in include/mflbm.h (libmflbm.so, sm_100a).

This is what the reference-side caller binds: the Fortran driver does it through
fortran/mflbm_iso_c.f90, the tests and bench.py through this module.  There is no CPU
fallback: if the shared library is missing or no B200 is present, calls raise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
CSRC_DIR = os.path.join(_HERE, "csrc")

SOLVER_SINGLEPHASE = 0
SOLVER_MULTIPHASE = 1


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "struct_size", "solver", "nx", "ny", "nz", "nxGlobal", "nyGlobal", "nzGlobal", "idz", "npz", "jper", "kper",
        "domain_wall_status_z_min", "domain_wall_status_z_max", "inlet_BC", "outlet_BC", "porous_plate_cmd",
        "Z_porous_plate", "mrt", "iz_async", "num_solid_boundary", "num_fluid_boundary", "device", "use_nccl",
        "kernel_variant")] + [("reserved_i", C.c_int32 * 7)] + [(n, C.c_double) for n in (
            "la_nui1", "la_nui2", "gamma", "beta", "force_Z", "phi_inlet", "sa_inject", "relaxation", "uin_avg",
            "rho_in", "rho_out", "s_e", "s_e2", "s_q", "s_nu", "s_pi", "s_t")] + [
                ("reserved_d", C.c_double * 8), ("nccl_unique_id", C.c_ubyte * 128)]


_DP = C.POINTER(C.c_double)


class Arrays(C.Structure):
    _fields_ = [("f", _DP * 19), ("g", _DP * 19)] + [(n, _DP) for n in (
        "phi", "phi_old", "cn_x", "cn_y", "cn_z", "c_norm", "curv", "u", "v", "w", "rho")] + [
            ("walls", C.POINTER(C.c_int8)), ("w_in", _DP), ("f_convec_bc", _DP), ("g_convec_bc", _DP),
            ("phi_convec_bc", _DP), ("solid_boundary_nodes", C.c_void_p), ("fluid_boundary_nodes", C.c_void_p)]


SOLID_DTYPE = np.dtype([("ix", "<i4"), ("iy", "<i4"), ("iz", "<i4"), ("i_fluid_num", "<i4"),
                        ("neighbor_list", "<i4", (18,)), ("la_weight", "<f8")])
FLUID_DTYPE = np.dtype([("ix", "<i4"), ("iy", "<i4"), ("iz", "<i4"), ("pad_", "<i4"),
                        ("nwx", "<f8"), ("nwy", "<f8"), ("nwz", "<f8"), ("theta", "<f8")])

EXPORTS = ("mflbm_create", "mflbm_destroy", "mflbm_last_error", "mflbm_version", "mflbm_upload", "mflbm_download",
           "mflbm_step", "mflbm_run", "mflbm_color_gradient", "mflbm_compute_macro_vars", "mflbm_monitor",
           "mflbm_cal_saturation", "mflbm_monitor_breakthrough", "mflbm_monitor_steady_phasefield",
           "mflbm_monitor_steady_capillarypressure", "mflbm_set_parameter", "mflbm_sync", "mflbm_timer_start",
           "mflbm_timer_stop", "mflbm_profile", "mflbm_profile_read", "mflbm_launch_count", "mflbm_device_bytes", "mflbm_nccl_unique_id",
           "mflbm_tile_stats", "mflbm_chain_info", "mflbm_chain_selfcheck", "mflbm_step_streamed", "mflbm_stream_flush", "mflbm_geometry_preprocess", "mflbm_geometry_free", "mflbm_geometry_last_error",
           "mflbm_output_begin", "mflbm_output_end", "mflbm_checkpoint_begin", "mflbm_checkpoint_fetch", "mflbm_checkpoint_end")


class GeometryConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("struct_size", "nxGlobal", "nyGlobal", "nzGlobal", "wk0", "wk1", "idz", "npz", "iper", "jper",
                                         "kper", "device")] + [("theta", C.c_double)]


class MflbmError(RuntimeError):
    pass


def build(verbose=False):
    """Compile libmflbm.so / libmflbm_strict.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", CSRC_DIR, "-j8"], stdout=out)
    return os.path.join(LIB_DIR, "libmflbm.so")


_LIBS = {}


def load(strict=False):
    """dlopen the in-tree CUDA library; raises if it has not been built (no fallback)."""
    key = bool(strict)
    if key in _LIBS:
        return _LIBS[key]
    # MFLBM_LIB_VARIANT=name selects an experimental build of the same sources (make variant NAME=name EXTRA=...):
    # used by tools/sweep.sh to compare kernel tunings in one GPU session
    variant = os.environ.get("MFLBM_LIB_VARIANT", "")
    name = "libmflbm_strict.so" if strict else ("libmflbm_%s.so" % variant if variant else "libmflbm.so")
    path = os.path.join(LIB_DIR, name)
    if not os.path.exists(path):
        raise MflbmError("%s not built: run __graft_entry__.build() (there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.mflbm_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.mflbm_destroy.argtypes = [vp]
    lib.mflbm_destroy.restype = None
    lib.mflbm_last_error.argtypes = [vp]
    lib.mflbm_last_error.restype = C.c_char_p
    lib.mflbm_version.restype = C.c_char_p
    lib.mflbm_upload.argtypes = [vp, C.POINTER(Arrays)]
    lib.mflbm_download.argtypes = [vp, C.POINTER(Arrays)]
    lib.mflbm_step.argtypes = [vp, C.c_int]
    lib.mflbm_run.argtypes = [vp, C.c_int, C.c_int]
    lib.mflbm_color_gradient.argtypes = [vp]
    lib.mflbm_compute_macro_vars.argtypes = [vp]
    lib.mflbm_monitor.argtypes = [vp, _DP, C.c_int]
    lib.mflbm_cal_saturation.argtypes = [vp, _DP, _DP]
    lib.mflbm_monitor_breakthrough.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.mflbm_monitor_steady_phasefield.argtypes = [vp, _DP, _DP]
    lib.mflbm_monitor_steady_capillarypressure.argtypes = [vp, _DP, _DP, _DP, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.mflbm_set_parameter.argtypes = [vp, C.c_char_p, C.c_double]
    lib.mflbm_sync.argtypes = [vp]
    lib.mflbm_timer_start.argtypes = [vp]
    lib.mflbm_timer_stop.argtypes = [vp, _DP]
    lib.mflbm_profile.argtypes = [vp, C.c_int]
    lib.mflbm_profile_read.argtypes = [vp, _DP, C.POINTER(C.c_longlong)]
    lib.mflbm_tile_stats.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    lib.mflbm_chain_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.mflbm_chain_selfcheck.argtypes = [vp, C.POINTER(C.c_longlong)]
    lib.mflbm_step_streamed.argtypes = [vp, C.c_int, C.c_void_p, _DP, _DP, C.POINTER(C.c_int)]
    lib.mflbm_stream_flush.argtypes = [vp, _DP, _DP]
    lib.mflbm_launch_count.argtypes = [vp]
    lib.mflbm_launch_count.restype = C.c_longlong
    lib.mflbm_device_bytes.argtypes = [vp]
    lib.mflbm_device_bytes.restype = C.c_longlong
    lib.mflbmx_spec_steps.argtypes = [vp]
    lib.mflbmx_spec_steps.restype = C.c_longlong
    lib.mflbmx_spec_info.argtypes = [vp, C.POINTER(C.c_int * 16)]
    lib.mflbmx_spec_info.restype = None
    lib.mflbm_nccl_unique_id.argtypes = [C.POINTER(C.c_ubyte * 128)]
    lib.mflbm_geometry_preprocess.argtypes = [C.POINTER(GeometryConfig), C.c_void_p, C.POINTER(vp), C.POINTER(C.c_int32), C.POINTER(vp),
                                              C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.mflbm_output_begin.argtypes = [vp, C.c_int]
    lib.mflbm_output_end.argtypes = [vp, C.POINTER(Arrays)]
    lib.mflbm_checkpoint_begin.argtypes = [vp]
    lib.mflbm_checkpoint_fetch.argtypes = [vp, C.POINTER(Arrays)]
    lib.mflbm_checkpoint_end.argtypes = [vp]
    lib.mflbm_geometry_free.argtypes = [vp]
    lib.mflbm_geometry_free.restype = None
    lib.mflbm_geometry_last_error.restype = C.c_char_p
    _LIBS[key] = lib
    return lib


def geometry_preprocess(walls_window, nzGlobal=None, wk0=1, idz=0, npz=1, periodic=(0, 0, 0), theta=0.0, device=-1, strict=False):
    """geometry_preprocessing_new (MP/Geometry_preprocessing.F90:9-512) on the device through the C ABI.

    walls_window: int8 (nxGlobal, nyGlobal, planes) = global planes wk0 .. wk0+planes-1 of the wall array.  Returns
    (solid_nodes, fluid_nodes, num_solid_scanned, num_fluid_scanned) with the node lists as numpy record arrays of the
    reference's derived types (SOLID_DTYPE / FLUID_DTYPE), local to slab idz of npz."""
    lib = load(strict)
    w = np.asfortranarray(walls_window, dtype=np.int8)
    nx, ny, planes = w.shape
    cfg = GeometryConfig(struct_size=C.sizeof(GeometryConfig), nxGlobal=nx, nyGlobal=ny, nzGlobal=nzGlobal or planes, wk0=wk0,
                         wk1=wk0 + planes - 1, idz=idz, npz=npz, iper=periodic[0], jper=periodic[1], kper=periodic[2], device=device,
                         theta=theta)
    ps, pf = C.c_void_p(), C.c_void_p()
    ns, nf = C.c_int32(), C.c_int32()
    gs, gf = C.c_int64(), C.c_int64()
    rc = lib.mflbm_geometry_preprocess(C.byref(cfg), w.ctypes.data, C.byref(ps), C.byref(ns), C.byref(pf), C.byref(nf), C.byref(gs),
                                       C.byref(gf))
    if rc:
        raise MflbmError("mflbm_geometry_preprocess: %s" % lib.mflbm_geometry_last_error().decode())
    try:
        solid = np.frombuffer((C.c_char * (96 * ns.value)).from_address(ps.value), dtype=SOLID_DTYPE).copy() if ns.value else np.zeros(0, SOLID_DTYPE)
        fluid = np.frombuffer((C.c_char * (48 * nf.value)).from_address(pf.value), dtype=FLUID_DTYPE).copy() if nf.value else np.zeros(0, FLUID_DTYPE)
    finally:
        lib.mflbm_geometry_free(ps)
        lib.mflbm_geometry_free(pf)
    return solid, fluid, gs.value, gf.value


def nccl_unique_id():
    lib = load()
    buf = (C.c_ubyte * 128)()
    rc = lib.mflbm_nccl_unique_id(C.byref(buf))
    if rc:
        raise MflbmError("mflbm_nccl_unique_id: %s" % lib.mflbm_last_error(None).decode())
    return bytes(buf)


def field_shape(name, nx, ny, nz):
    """Host (Fortran) extents of each array of the ABI, MP/Init_multiphase.F90:594-658."""
    if name in ("phi", "phi_old"):
        return (nx + 8, ny + 8, nz + 8)
    if name in ("cn_x", "cn_y", "cn_z", "c_norm", "walls"):
        return (nx + 4, ny + 4, nz + 4)
    if name in ("w_in", "phi_convec_bc"):
        return (nx + 2, ny + 2)
    if name in ("f_convec_bc", "g_convec_bc"):
        return (nx + 2, ny + 2, 19)
    return (nx + 2, ny + 2, nz + 2)


class Context:
    """One slab on one GPU (mflbm_ctx)."""

    def __init__(self, strict=False, **cfg):
        self.lib = load(strict)
        c = Config()
        c.struct_size = C.sizeof(Config)
        c.device = -1
        c.mrt = 2
        c.relaxation = 1.0
        c.rho_in = 1.0
        c.rho_out = 1.0
        c.npz = 1
        nccl_id = cfg.pop("nccl_unique_id", None)
        for k, v in cfg.items():
            if not hasattr(c, k):
                raise KeyError(k)
            setattr(c, k, v)
        if nccl_id is not None:
            C.memmove(c.nccl_unique_id, nccl_id, 128)
        if c.nxGlobal == 0:
            c.nxGlobal, c.nyGlobal, c.nzGlobal = c.nx, c.ny, c.nz * c.npz
        self.cfg = c
        self.nx, self.ny, self.nz = c.nx, c.ny, c.nz
        self.mp = c.solver == SOLVER_MULTIPHASE
        h = C.c_void_p()
        rc = self.lib.mflbm_create(C.byref(c), C.byref(h))
        if rc:
            raise MflbmError("mflbm_create rc=%d: %s" % (rc, self.lib.mflbm_last_error(None).decode()))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.mflbm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc:
            raise MflbmError("%s rc=%d: %s" % (what, rc, self.lib.mflbm_last_error(self.h).decode()))

    def _arrays(self, d, keep):
        a = Arrays()
        for k, v in d.items():
            if v is None:
                continue
            if k in ("f", "g"):
                for q in range(19):
                    arr = v[q]
                    assert arr.dtype == np.float64 and arr.flags.f_contiguous and arr.shape == field_shape("f", self.nx, self.ny, self.nz)
                    keep.append(arr)
                    getattr(a, k)[q] = arr.ctypes.data_as(_DP)
            elif k == "walls":
                assert v.dtype == np.int8 and v.flags.f_contiguous and v.shape == field_shape("walls", self.nx, self.ny, self.nz)
                keep.append(v)
                a.walls = v.ctypes.data_as(C.POINTER(C.c_int8))
            elif k == "solid_boundary_nodes":
                v = np.ascontiguousarray(v, dtype=SOLID_DTYPE)
                assert len(v) == self.cfg.num_solid_boundary
                keep.append(v)
                a.solid_boundary_nodes = v.ctypes.data
            elif k == "fluid_boundary_nodes":
                v = np.ascontiguousarray(v, dtype=FLUID_DTYPE)
                assert len(v) == self.cfg.num_fluid_boundary
                keep.append(v)
                a.fluid_boundary_nodes = v.ctypes.data
            else:
                assert v.dtype == np.float64 and v.flags.f_contiguous, k
                assert v.shape == field_shape(k, self.nx, self.ny, self.nz), (k, v.shape)
                keep.append(v)
                setattr(a, k, v.ctypes.data_as(_DP))
        return a

    def upload(self, **arrays):
        keep = []
        a = self._arrays(arrays, keep)
        self._chk(self.lib.mflbm_upload(self.h, C.byref(a)), "mflbm_upload")

    def download(self, *names):
        """Returns {name: ndarray} in the reference's host layout; 'f'/'g' give lists of 19 arrays."""
        out = {}
        for n in names:
            if n in ("f", "g"):
                out[n] = [np.zeros(field_shape("f", self.nx, self.ny, self.nz), order="F") for _ in range(19)]
            else:
                out[n] = np.zeros(field_shape(n, self.nx, self.ny, self.nz), order="F")
        keep = []
        a = self._arrays(out, keep)
        self._chk(self.lib.mflbm_download(self.h, C.byref(a)), "mflbm_download")
        return out

    def step(self, ntime):
        self._chk(self.lib.mflbm_step(self.h, ntime), "mflbm_step")

    def run(self, ntime0, nsteps):
        self._chk(self.lib.mflbm_run(self.h, ntime0, nsteps), "mflbm_run")

    def step_streamed(self, ntime, w_in_ptr=None):
        """one streamed step; returns (v1, v2) of the PREVIOUS streamed step or None (first call / singlephase)"""
        v1, v2, have = C.c_double(), C.c_double(), C.c_int()
        self._chk(self.lib.mflbm_step_streamed(self.h, ntime, w_in_ptr, C.byref(v1), C.byref(v2), C.byref(have)), "mflbm_step_streamed")
        return (v1.value, v2.value) if have.value else None

    def stream_flush(self):
        v1, v2 = C.c_double(), C.c_double()
        self._chk(self.lib.mflbm_stream_flush(self.h, C.byref(v1), C.byref(v2)), "mflbm_stream_flush")
        return v1.value, v2.value

    def color_gradient(self):
        self._chk(self.lib.mflbm_color_gradient(self.h), "mflbm_color_gradient")

    def compute_macro_vars(self):
        self._chk(self.lib.mflbm_compute_macro_vars(self.h), "mflbm_compute_macro_vars")

    OUT_PHI, OUT_MACRO = 1, 2

    def output_begin(self, what):
        """snapshot phi and / or u,v,w,rho at the current step and start the asynchronous device-to-host copy"""
        self._chk(self.lib.mflbm_output_begin(self.h, int(what)), "mflbm_output_begin")

    def output_end(self, *names):
        """wait for the staged output and return {name: ndarray} like download()"""
        out = {n: np.zeros(field_shape(n, self.nx, self.ny, self.nz), order="F") for n in names}
        keep = []
        a = self._arrays(out, keep)
        self._chk(self.lib.mflbm_output_end(self.h, C.byref(a)), "mflbm_output_end")
        return out

    CKPT_NAMES = ("f", "g", "phi", "f_convec_bc", "g_convec_bc", "phi_convec_bc")

    def checkpoint_begin(self):
        """freeze the state of the current step; returns 0 (device snapshot, stepping may go on) or 1 (context frozen)"""
        rc = self.lib.mflbm_checkpoint_begin(self.h)
        if rc < 0:
            self._chk(rc, "mflbm_checkpoint_begin")
        return rc

    def checkpoint_fetch(self, *names, into=None):
        """the frozen state of the named arrays in the reference's host layout ({name: ndarray}, 'f'/'g' lists of 19);
        `into` = arrays to fill instead of fresh zero arrays (dead entries of the sparse layout keep their values)"""
        out = into if into is not None else {}
        for n in names:
            if n in out:
                continue
            if n in ("f", "g"):
                out[n] = [np.zeros(field_shape("f", self.nx, self.ny, self.nz), order="F") for _ in range(19)]
            else:
                out[n] = np.zeros(field_shape(n, self.nx, self.ny, self.nz), order="F")
        keep = []
        a = self._arrays({n: out[n] for n in names}, keep)
        self._chk(self.lib.mflbm_checkpoint_fetch(self.h, C.byref(a)), "mflbm_checkpoint_fetch")
        return out

    def checkpoint_end(self):
        self._chk(self.lib.mflbm_checkpoint_end(self.h), "mflbm_checkpoint_end")

    def monitor(self):
        """Device part of monitor: returns the reference's tk buffer split into named profiles."""
        nz = self.nz
        n = 7 * nz + 3 if self.mp else 2 * nz + 1
        tk = np.zeros(n)
        self._chk(self.lib.mflbm_monitor(self.h, tk.ctypes.data_as(_DP), n), "mflbm_monitor")
        if self.mp:
            names = ("fl1", "fl2", "vol1", "vol2", "mass1", "mass2", "pre")
            out = {nm: tk[m * nz:(m + 1) * nz].copy() for m, nm in enumerate(names)}
            out.update(umax=tk[7 * nz], usq1=tk[7 * nz + 1], usq2=tk[7 * nz + 2])
        else:
            out = {"fl": tk[:nz].copy(), "pre": tk[nz:2 * nz].copy(), "umax": tk[2 * nz]}
        out["tk"] = tk
        return out

    def cal_saturation(self):
        v1, v2 = C.c_double(), C.c_double()
        self._chk(self.lib.mflbm_cal_saturation(self.h, C.byref(v1), C.byref(v2)), "mflbm_cal_saturation")
        return v1.value, v2.value

    def monitor_breakthrough(self):
        n = C.c_int32()
        self._chk(self.lib.mflbm_monitor_breakthrough(self.h, C.byref(n)), "mflbm_monitor_breakthrough")
        return n.value

    def monitor_steady_phasefield(self):
        a, b = C.c_double(), C.c_double()
        self._chk(self.lib.mflbm_monitor_steady_phasefield(self.h, C.byref(a), C.byref(b)), "mflbm_monitor_steady_phasefield")
        return a.value, b.value

    def monitor_steady_capillarypressure(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        i, j = C.c_int32(), C.c_int32()
        self._chk(self.lib.mflbm_monitor_steady_capillarypressure(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(i), C.byref(j)),
                  "mflbm_monitor_steady_capillarypressure")
        return dict(umax=a.value, pre_w=b.value, pre_nw=c.value, i_w=i.value, i_nw=j.value)

    def set_parameter(self, name, value):
        self._chk(self.lib.mflbm_set_parameter(self.h, name.encode(), float(value)), "mflbm_set_parameter")

    def sync(self):
        self._chk(self.lib.mflbm_sync(self.h), "mflbm_sync")

    def timer_start(self):
        self._chk(self.lib.mflbm_timer_start(self.h), "mflbm_timer_start")

    def timer_stop(self):
        ms = C.c_double()
        self._chk(self.lib.mflbm_timer_stop(self.h, C.byref(ms)), "mflbm_timer_stop")
        return ms.value

    def profile(self, enable):
        self._chk(self.lib.mflbm_profile(self.h, int(enable)), "mflbm_profile")

    def profile_read(self):
        ms, n = C.c_double(), C.c_longlong()
        self._chk(self.lib.mflbm_profile_read(self.h, C.byref(ms), C.byref(n)), "mflbm_profile_read")
        return ms.value, n.value

    def tile_stats(self):
        """(number of tiles, number of quiet tiles) of the last gradient chain; (0, 0) without tiles"""
        a, b = C.c_longlong(), C.c_longlong()
        self._chk(self.lib.mflbm_tile_stats(self.h, C.byref(a), C.byref(b)), "mflbm_tile_stats")
        return a.value, b.value

    def chain_info(self):
        """(fused, reject_mask): whether color_gradient runs as the single fused kernel (csrc/march.cuh), and why not"""
        a, b = C.c_int(), C.c_int()
        self._chk(self.lib.mflbm_chain_info(self.h, C.byref(a), C.byref(b)), "mflbm_chain_info")
        return a.value, b.value

    def chain_selfcheck(self):
        """entries of the packed colour gradient that differ from a re-evaluation with the reference-order kernels"""
        n = C.c_longlong()
        self._chk(self.lib.mflbm_chain_selfcheck(self.h, C.byref(n)), "mflbm_chain_selfcheck")
        return n.value

    @property
    def launch_count(self):
        return self.lib.mflbm_launch_count(self.h)

    @property
    def device_bytes(self):
        return self.lib.mflbm_device_bytes(self.h)

    @property
    def spec_steps(self):
        """steps that ran with the speculative early gradient chain (DESIGN.md "Step schedule")"""
        return self.lib.mflbmx_spec_steps(self.h)
