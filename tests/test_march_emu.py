"""The march kernel (mf-lbm_b200/csrc/march.cuh: K3..K7 of the colour-gradient chain fused on chip) against the oracle, on the CPU.

No GPU exists where the CPU suite runs, so the kernel source is compiled by g++ as a sequential emulation (MARCH_EMU: the
phases of a block loop over the thread index, barriers vanish, cp.async becomes a copy) and every block of a launch is run
one after the other on device-layout inputs built the way mflbm_upload builds them.  What is checked is exactly what the
collision kernel consumes: G[0..2] = interface normal at the fluid nodes after K4 + K5, G[3] = 0.5*gamma*curv*|grad phi|
(MP/Kernel_multiphase.F90:118) -- bit for bit (the emulation is compiled with -ffp-contract=off like the strict CUDA build).
"""
import ctypes as C
import os
import subprocess
from importlib import import_module

import numpy as np
import pytest

import mflbm_b200 as M
from helpers import make_oracle
from oracle.oracle import Oracle, default_params

geo = import_module("mflbm_b200.geometry")
HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "march_emu", "march_emu.cpp")


@pytest.fixture(scope="module", params=["forward", "reverse"])
def emu(request):
    """reverse: the threads of every phase run in the opposite order -- a result that depended on the order of the threads
    between two barriers (a race on the device) would show"""
    lib_path = os.path.join(HERE, "march_emu", "libmarch_emu_%s.so" % request.param)
    deps = [SRC] + [os.path.join(HERE, "..", "mf-lbm_b200", "csrc", f) for f in ("march.cuh", "gradient.cuh", "mflbm_internal.cuh")]
    if not os.path.exists(lib_path) or any(os.path.getmtime(d) > os.path.getmtime(lib_path) for d in deps):
        extra = ["-DMARCH_EMU_REVERSE"] if request.param == "reverse" else []
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-DMARCH_EMU",
                               "-D__host__=", "-D__device__=", "-D__global__=", "-I/usr/local/cuda/include", "-o", lib_path, SRC] + extra)
    lib = C.CDLL(lib_path)
    lib.march_emu_run.restype = C.c_int
    return lib


def run_emu(lib, o, lz):
    nx, ny, nz = o.nx, o.ny, o.nz
    walls = np.asfortranarray(o.walls)
    phi = np.asfortranarray(o.field("phi"))
    solid = np.ascontiguousarray(o.solid_nodes())
    fluid = np.ascontiguousarray(o.fluid_nodes())
    fl = walls[2:-2, 2:-2, 2:-2] == 0
    nA = int(fl.sum())
    G = np.zeros((4, nA))
    ws = np.zeros((nA + 31) // 32, np.int32)
    flags = np.zeros(4, np.int32)
    vp = C.c_void_p
    rc = lib.march_emu_run(C.c_int(nx), C.c_int(ny), C.c_int(nz), walls.ctypes.data_as(vp), phi.ctypes.data_as(vp),
                           solid.ctypes.data_as(vp), C.c_int(len(solid)), fluid.ctypes.data_as(vp), C.c_int(len(fluid)),
                           C.c_double(o.p.gamma), C.c_int(lz), G.ctypes.data_as(vp), C.c_int(nA), ws.ctypes.data_as(vp),
                           flags.ctypes.data_as(vp))
    assert rc == 0, (rc, flags)
    assert flags[0] == 0, "cell codes inconsistent with the node lists: %d" % flags[0]
    return G, ws, fl


def expected(o, fl):
    """what k_gradient_pack would write, from the oracle's fields after color_gradient (raster order k, j, i)"""
    def at_fluid(a, ghost):
        inner = a[ghost:-ghost, ghost:-ghost, ghost:-ghost]
        return inner.transpose(2, 1, 0)[fl.transpose(2, 1, 0)]
    cn = [at_fluid(o.field(n), 2) for n in ("cn_x", "cn_y", "cn_z")]
    cnorm = at_fluid(o.field("c_norm"), 2)
    curv = at_fluid(o.field("curv"), 1)
    tmp = (0.5 * o.p.gamma) * curv * cnorm
    tmp = np.where(cnorm != 0.0, tmp, 0.0)
    return cn + [tmp]


def check(lib, o, lz):
    G, ws, fl = run_emu(lib, o, lz)
    o.color_gradient()
    exp = expected(o, fl)
    for q in range(4):
        bad = np.flatnonzero(G[q].view(np.int64) != exp[q].view(np.int64))
        # +0 / -0 are the same number to every consumer, but the kernels agree on the sign as well
        assert bad.size == 0, "G[%d]: %d of %d nodes differ, first %d: %r vs %r" % (q, bad.size, G[q].size, bad[0], G[q][bad[0]], exp[q][bad[0]])
    assert np.all(ws == 7)  # every warp of nodes was stamped
    return G


def _random_phi_case(wg, **kw):
    """the reference's benchmark state (initial_fluid_distribution_option 6, unseeded there: injected from here)"""
    nx, ny, nz = wg.shape
    p = default_params(nxG=nx, nyG=ny, nzG=nz, n_exclude_inlet=0, n_exclude_outlet=0, initial_fluid_distribution_option=5, **kw)
    o = Oracle(p)
    o.set_walls(wg); o.geometry_preprocess(); o.init_basic(); o.init_phi()
    rng = np.random.default_rng(nx * 1000 + nz)
    o.field("phi")[...] = np.where(rng.random(o.field("phi").shape) > 0.4, -1.0, 1.0)
    o.init_pdf()
    return o


def test_smem_fits(emu):
    emu.march_emu_smem_bytes.restype = C.c_int
    assert emu.march_emu_smem_bytes() <= 227 * 1024


@pytest.mark.parametrize("lz", [64, 16, 5])
def test_c1_tube_sphere(emu, lz):
    o = make_oracle(modify_geometry_cmd=1)
    for n in range(1, 5):  # a few steps: phi leaves its initial +-1 and the interface gets its tanh profile
        o.step(n)
    G = check(emu, o, lz)
    assert np.count_nonzero(G[3]) > 100


@pytest.mark.parametrize("dims", [(72, 64, 40), (33, 17, 21), (32, 16, 16), (31, 47, 9)])
def test_sphere_pack_random_phi(emu, dims):
    """interface everywhere (the reference's benchmark case 6), tile-unaligned lattices, wetting angle 150 degrees"""
    nx, ny, nz = dims
    wg = geo.sphere_pack(nx, ny, nz, periodic=True, porosity=0.45, rmin=3.0, rmax=7.0, seed=5, buffer=0)
    o = _random_phi_case(wg, kper=1, inlet_BC=0, outlet_BC=0, force_z0=2e-4, la_nu2=0.04, theta_deg=150.0)
    check(emu, o, 16)  # phi = +-1 per node, ghost layers drawn independently of their periodic images
    for n in range(1, 4):
        o.step(n)
    G = check(emu, o, 16)
    assert np.count_nonzero(G[3]) > G.shape[1] // 4


def test_drainage_front_open(emu):
    wg = geo.sphere_pack(48, 40, 56, periodic=False, porosity=0.4, rmin=4.0, rmax=8.0, seed=12, buffer=6)
    o = make_oracle(nxG=48, nyG=40, nzG=56, la_nu2=0.04, interface_z0=8.0, walls_global=wg, n_exclude_inlet=6, n_exclude_outlet=6)
    check(emu, o, 64)  # initial state, before any step
    for n in range(1, 7):
        o.step(n)
    check(emu, o, 16)


def test_y_periodic(emu):
    wg = geo.sphere_pack(24, 32, 24, periodic=True, porosity=0.5, rmin=3.0, rmax=6.0, seed=3, buffer=0)
    o = _random_phi_case(wg, jper=1, kper=1, wsy0=0, wsy1=0, inlet_BC=0, outlet_BC=0, force_z0=1e-4)
    for n in range(1, 4):
        o.step(n)
    G = check(emu, o, 16)
    assert np.count_nonzero(G[3]) > G.shape[1] // 4
