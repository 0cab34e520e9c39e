"""ctypes binding of the CPU oracle (oracle/mflbm_oracle.c).

TEST INFRASTRUCTURE ONLY: may be imported from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference arm.  Never from the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class OrcParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "multiphase", "nxG", "nyG", "nzG", "idz", "npz", "iper", "jper", "kper",
        "wsx0", "wsx1", "wsy0", "wsy1", "wsz0", "wsz1", "n_exclude_inlet", "n_exclude_outlet",
        "inlet_BC", "outlet_BC", "porous_plate_cmd", "Z_porous_plate", "mrt",
        "initial_fluid_distribution_option", "modify_geometry_cmd", "mrt_para_preset",
        "steady_state_option")] + [("reserved_i", C.c_int32 * 3)] + [(n, C.c_double) for n in (
            "la_nu1", "la_nu2", "gamma", "beta", "theta_deg", "force_z0", "sa_inject", "ca_0",
            "interface_z0", "rho_drop", "Re", "char_length", "target_inject_pore_volume")] + [
                ("reserved_d", C.c_double * 4)]


class SolidNode(C.Structure):
    _fields_ = [("ix", C.c_int32), ("iy", C.c_int32), ("iz", C.c_int32), ("i_fluid_num", C.c_int32),
                ("neighbor_list", C.c_int32 * 18), ("la_weight", C.c_double)]


class FluidNode(C.Structure):
    _fields_ = [("ix", C.c_int32), ("iy", C.c_int32), ("iz", C.c_int32), ("pad_", C.c_int32),
                ("nwx", C.c_double), ("nwy", C.c_double), ("nwz", C.c_double), ("theta", C.c_double)]


SOLID_DTYPE = np.dtype([("ix", "<i4"), ("iy", "<i4"), ("iz", "<i4"), ("i_fluid_num", "<i4"),
                        ("neighbor_list", "<i4", (18,)), ("la_weight", "<f8")])
FLUID_DTYPE = np.dtype([("ix", "<i4"), ("iy", "<i4"), ("iz", "<i4"), ("pad_", "<i4"),
                        ("nwx", "<f8"), ("nwy", "<f8"), ("nwz", "<f8"), ("theta", "<f8")])
assert SOLID_DTYPE.itemsize == 96 and FLUID_DTYPE.itemsize == 48


class MonitorOut(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "umax", "usq1", "usq2", "saturation", "saturation_full_domain", "vol1_sum", "vol2_sum",
        "mass1_sum", "mass2_sum", "fl_avg_whole", "fl1_avg_whole", "fl2_avg_whole", "fl_avg", "fl1_avg",
        "fl2_avg", "ca", "umax_global", "pre_in", "pre_out", "d_phi_max", "pre_w", "pre_nw")] + [
            (n, C.c_int32) for n in ("i_w", "i_nw", "outlet_phase1_sum", "pad_")]


def build(fast=False):
    """Compile the oracle with the committed Makefile (gcc only) and return the .so path."""
    name = "libmflbm_oracle_fast.so" if fast else "libmflbm_oracle.so"
    path = os.path.join(_HERE, name)
    src = os.path.join(_HERE, "mflbm_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, name], stdout=subprocess.DEVNULL)
    return path


_LIBS = {}


def _lib(fast=False):
    if fast in _LIBS:
        return _LIBS[fast]
    lib = C.CDLL(build(fast))
    vp = C.c_void_p
    lib.orc_create.restype = vp
    lib.orc_create.argtypes = [C.POINTER(OrcParams)]
    for fn in ("orc_destroy", "orc_set_walls", "orc_geometry_preprocess", "orc_init_basic", "orc_init_phi",
               "orc_init_pdf", "orc_color_gradient", "orc_compute_macro_vars"):
        getattr(lib, fn).argtypes = [vp]
        getattr(lib, fn).restype = None
    for fn in ("orc_kernel_odd", "orc_kernel_even"):
        getattr(lib, fn).argtypes = [vp] + [C.c_int] * 6
        getattr(lib, fn).restype = None
    lib.orc_step.argtypes = [vp, C.c_int]
    lib.orc_step.restype = None
    for fn in ("orc_monitor", "orc_cal_saturation", "orc_monitor_breakthrough", "orc_monitor_steady_phasefield",
               "orc_monitor_steady_capillarypressure"):
        getattr(lib, fn).argtypes = [vp, C.POINTER(MonitorOut)]
        getattr(lib, fn).restype = None
    lib.orc_walls_global.restype = C.POINTER(C.c_int8)
    lib.orc_walls_global.argtypes = [vp]
    lib.orc_walls.restype = C.POINTER(C.c_int8)
    lib.orc_walls.argtypes = [vp]
    lib.orc_f.restype = C.POINTER(C.c_double)
    lib.orc_f.argtypes = [vp, C.c_int]
    lib.orc_g.restype = C.POINTER(C.c_double)
    lib.orc_g.argtypes = [vp, C.c_int]
    lib.orc_field.restype = C.POINTER(C.c_double)
    lib.orc_field.argtypes = [vp, C.c_char_p]
    lib.orc_solid_nodes.restype = C.POINTER(SolidNode)
    lib.orc_solid_nodes.argtypes = [vp, C.POINTER(C.c_int)]
    lib.orc_fluid_nodes.restype = C.POINTER(FluidNode)
    lib.orc_fluid_nodes.argtypes = [vp, C.POINTER(C.c_int)]
    lib.orc_get_int.restype = C.c_int
    lib.orc_get_int.argtypes = [vp, C.c_char_p]
    lib.orc_get_i64.restype = C.c_longlong
    lib.orc_get_i64.argtypes = [vp, C.c_char_p]
    lib.orc_get_double.restype = C.c_double
    lib.orc_get_double.argtypes = [vp, C.c_char_p]
    lib.orc_set_double.argtypes = [vp, C.c_char_p, C.c_double]
    lib.orc_set_double.restype = None
    lib.orc_set_int.argtypes = [vp, C.c_char_p, C.c_int]
    lib.orc_set_int.restype = None
    _LIBS[fast] = lib
    return lib


def default_params(**kw):
    """Template defaults of multiphase_3D/run_template/template-simulation_control.txt."""
    d = dict(multiphase=1, nxG=40, nyG=40, nzG=60, idz=0, npz=1, iper=0, jper=0, kper=0,
             wsx0=1, wsx1=1, wsy0=1, wsy1=1, wsz0=0, wsz1=0, n_exclude_inlet=10, n_exclude_outlet=10,
             inlet_BC=1, outlet_BC=1, porous_plate_cmd=0, Z_porous_plate=0, mrt=2,
             initial_fluid_distribution_option=1, modify_geometry_cmd=0, mrt_para_preset=1,
             steady_state_option=0,
             la_nu1=0.004, la_nu2=0.4, gamma=0.03, beta=0.95, theta_deg=30.0, force_z0=0.0, sa_inject=1.0,
             ca_0=100e-6, interface_z0=8.0, rho_drop=0.0, Re=1.0, char_length=1.0,
             target_inject_pore_volume=1.0)
    d.update(kw)
    p = OrcParams()
    for k, v in d.items():
        setattr(p, k, v)
    return p


class Oracle:
    """One slab (idz of npz) of the reference solver state on the CPU."""

    def __init__(self, params, fast=False):
        self.lib = _lib(fast)
        self.p = params
        self.h = self.lib.orc_create(C.byref(params))
        if not self.h:
            raise RuntimeError("orc_create failed")
        self.nx = self.lib.orc_get_int(self.h, b"nx")
        self.ny = self.lib.orc_get_int(self.h, b"ny")
        self.nz = self.lib.orc_get_int(self.h, b"nz")
        self.mp = bool(params.multiphase)

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- views (Fortran order; index [i+o-1, j+o-1, k+o-1] for ghost width o) ----
    def _view(self, ptr, shape, dtype=np.float64):
        n = int(np.prod(shape))
        arr = np.ctypeslib.as_array(ptr, shape=(n,))
        return arr.view(dtype).reshape(shape, order="F")

    @property
    def walls_global(self):
        return self._view(self.lib.orc_walls_global(self.h), (self.p.nxG, self.p.nyG, self.p.nzG), np.int8)

    @property
    def walls(self):
        return self._view(self.lib.orc_walls(self.h), (self.nx + 4, self.ny + 4, self.nz + 4), np.int8)

    def f(self, q):
        return self._view(self.lib.orc_f(self.h, q), (self.nx + 2, self.ny + 2, self.nz + 2))

    def g(self, q):
        return self._view(self.lib.orc_g(self.h, q), (self.nx + 2, self.ny + 2, self.nz + 2))

    def field(self, name):
        nx, ny, nz = self.nx, self.ny, self.nz
        shapes = {"phi": (nx + 8, ny + 8, nz + 8), "phi_old": (nx + 8, ny + 8, nz + 8),
                  "cn_x": (nx + 4, ny + 4, nz + 4), "cn_y": (nx + 4, ny + 4, nz + 4), "cn_z": (nx + 4, ny + 4, nz + 4),
                  "c_norm": (nx + 4, ny + 4, nz + 4), "curv": (nx + 2, ny + 2, nz + 2),
                  "u": (nx + 2, ny + 2, nz + 2), "v": (nx + 2, ny + 2, nz + 2), "w": (nx + 2, ny + 2, nz + 2),
                  "rho": (nx + 2, ny + 2, nz + 2), "w_in": (nx + 2, ny + 2),
                  "f_convec_bc": (nx + 2, ny + 2, 19), "g_convec_bc": (nx + 2, ny + 2, 19),
                  "phi_convec_bc": (nx + 2, ny + 2)}
        for n in ("fl1", "fl2", "vol1", "vol2", "mass1", "mass2", "pre"):
            shapes[n] = (nz,)
        ptr = self.lib.orc_field(self.h, name.encode())
        if not ptr:
            raise KeyError(name)
        return self._view(ptr, shapes[name])

    def solid_nodes(self):
        n = C.c_int(0)
        ptr = self.lib.orc_solid_nodes(self.h, C.byref(n))
        if n.value == 0:
            return np.zeros(0, SOLID_DTYPE)
        buf = (C.c_char * (96 * n.value)).from_address(C.addressof(ptr.contents))
        return np.frombuffer(buf, dtype=SOLID_DTYPE)

    def fluid_nodes(self):
        n = C.c_int(0)
        ptr = self.lib.orc_fluid_nodes(self.h, C.byref(n))
        if n.value == 0:
            return np.zeros(0, FLUID_DTYPE)
        buf = (C.c_char * (48 * n.value)).from_address(C.addressof(ptr.contents))
        return np.frombuffer(buf, dtype=FLUID_DTYPE)

    def get_int(self, name):
        return self.lib.orc_get_int(self.h, name.encode())

    def get_i64(self, name):
        return self.lib.orc_get_i64(self.h, name.encode())

    def get_double(self, name):
        return self.lib.orc_get_double(self.h, name.encode())

    def set_double(self, name, v):
        self.lib.orc_set_double(self.h, name.encode(), float(v))

    def set_int(self, name, v):
        self.lib.orc_set_int(self.h, name.encode(), int(v))

    # ---- reference routines ----
    def set_walls(self, walls_global=None):
        if walls_global is not None:
            wg = self.walls_global
            wg[...] = 0
            sx, sy, sz = walls_global.shape
            wg[:sx, :sy, :sz] = walls_global
        self.lib.orc_set_walls(self.h)

    def geometry_preprocess(self):
        self.lib.orc_geometry_preprocess(self.h)

    def init_basic(self):
        self.lib.orc_init_basic(self.h)

    def init_phi(self):
        self.lib.orc_init_phi(self.h)

    def init_pdf(self):
        self.lib.orc_init_pdf(self.h)

    def setup(self, walls_global=None):
        """initialization_basic_multi + initialization_new_multi (MP/Main_multiphase.F90:90-95)."""
        self.set_walls(walls_global)
        if self.mp:
            self.geometry_preprocess()
        self.init_basic()
        self.init_phi()
        self.init_pdf()

    def color_gradient(self):
        self.lib.orc_color_gradient(self.h)

    def step(self, ntime):
        self.lib.orc_step(self.h, ntime)

    def kernel_odd(self, *r):
        self.lib.orc_kernel_odd(self.h, *r)

    def kernel_even(self, *r):
        self.lib.orc_kernel_even(self.h, *r)

    def compute_macro_vars(self):
        self.lib.orc_compute_macro_vars(self.h)

    def _mon(self, fn):
        o = MonitorOut()
        getattr(self.lib, fn)(self.h, C.byref(o))
        return {n: getattr(o, n) for n, _ in MonitorOut._fields_ if n != "pad_"}

    def monitor(self):
        return self._mon("orc_monitor")

    def cal_saturation(self):
        return self._mon("orc_cal_saturation")

    def monitor_breakthrough(self):
        return self._mon("orc_monitor_breakthrough")

    def monitor_steady_phasefield(self):
        return self._mon("orc_monitor_steady_phasefield")

    def monitor_steady_capillarypressure(self):
        return self._mon("orc_monitor_steady_capillarypressure")


def read_wall_array(path):
    """Reference wall-array format: 3x int32 (nx,ny,nz) then int8 walls, i fastest
    (preprocessing/1.create_geometry_to_WallArray/sample_code_3d_geometry.f90:75-77)."""
    with open(path, "rb") as fh:
        nx, ny, nz = np.frombuffer(fh.read(12), dtype="<i4")
        w = np.frombuffer(fh.read(int(nx) * int(ny) * int(nz)), dtype=np.int8)
    return w.reshape((nx, ny, nz), order="F")
