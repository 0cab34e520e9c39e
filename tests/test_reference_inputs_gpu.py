"""GPU parity on the inputs the REFERENCE holds (committed as golden input fixtures by tests/golden/make_fixtures.py):

  tube_sphere.npz                    MF-LBM-extFiles/geometry_files/tube_sphere_example/tube_sphere.dat, 60x60x80 -- the
                                     external-geometry drainage case of the test suite (template parameters)
  bentheimer_in10_240_out10.bits.xz  the reference's Bentheimer wall array (benchmark cases 4-8), here a 96^3 centre
                                     crop so that the CPU oracle finishes in seconds, with the parameters of
                                       case 4  imbibition:  nu2 0.04, theta 150, Ca 1e-4, saturation_injection 0, open z
                                       case 5  fractional flow: z-periodic, body force 2e-4, random phi at S = 0.4 (seeded here)
                                       case 7  singlephase absolute permeability: nu 0.1, body force 1e-5, z-periodic

CUDA path through the C ABI against the oracle: bit-exact in the strict (-fmad=false) build after 10 steps, 1e-12 in the
default FMA build after one odd + one even step."""
import os

import numpy as np
import pytest

import mflbm_b200 as M
from helpers import compare_state, ctx_from_oracle, make_oracle
from oracle.oracle import Oracle, default_params

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LAYOUTS = [pytest.param(1, id="dense"), pytest.param(2, id="sparse")]


def _run(o, ctx, n, t0=1):
    for t in range(t0, t0 + n):
        o.step(t)
    ctx.run(t0, n)
    ctx.sync()


def _bentheimer_crop(n=96):
    from importlib import import_module
    geo = import_module("mflbm_b200.geometry")
    w = geo.load_packed_walls(os.path.join(GOLDEN, "bentheimer_in10_240_out10.bits.xz"), (240, 240, 260))
    assert int((w == 0).sum()) == 3670813
    i0, k0 = (240 - n) // 2, (260 - n) // 2
    return np.ascontiguousarray(w[i0:i0 + n, i0:i0 + n, k0:k0 + n])


def _check(o, strict, layout, steps_strict=10, tol_fma=1e-12):
    ctx = ctx_from_oracle(o, strict=strict, kernel_variant=layout)
    sp = layout == 2
    if o.mp:
        o.color_gradient()
        ctx.color_gradient()
        compare_state(ctx, o, 0.0 if strict else tol_fma, sparse=sp)
    if strict:
        _run(o, ctx, steps_strict)
        compare_state(ctx, o, 0.0, sparse=sp)
    else:
        _run(o, ctx, 2)  # one odd + one even step
        compare_state(ctx, o, tol_fma, sparse=sp)
    ctx.close()


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fma"])
def test_tube_sphere_external_geometry_drainage(strict, layout):
    w = np.load(os.path.join(GOLDEN, "tube_sphere.npz"))["walls"]
    assert w.shape == (60, 60, 80) and int((w == 0).sum()) == 229816
    o = make_oracle(nxG=60, nyG=60, nzG=80, walls_global=w, n_exclude_inlet=5, n_exclude_outlet=5)
    _check(o, strict, layout)


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fma"])
def test_bentheimer_crop_imbibition_case4(strict, layout):
    w = _bentheimer_crop()
    w[:, :, :8] = 0   # the crop has no inlet / outlet buffer of its own: keep 8 open planes at each end like the file does
    w[:, :, -8:] = 0
    o = make_oracle(nxG=96, nyG=96, nzG=96, walls_global=w, la_nu2=0.04, theta_deg=150.0, ca_0=100e-6, sa_inject=0.0,
                    initial_fluid_distribution_option=2, n_exclude_inlet=8, n_exclude_outlet=8)
    # FMA build: the normalised gradient n = grad(phi) / |grad(phi)| and the wetting alteration (1 / sqrt(1 - t^2), theta = 150)
    # amplify last-bit differences where |grad(phi)| is just above the 1e-6 cut-off; measured 8e-9 on n, 2e-10 on the populations
    _check(o, strict, layout, tol_fma=1e-7)


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fma"])
def test_bentheimer_crop_fractional_flow_case5(strict, layout):
    """z-periodic, body-force driven, phase field of initial_fluid_distribution_option 6 (random, S = 0.4) injected from the
    host because the reference's draw is unseeded: interface everywhere, no quiet tile."""
    w = _bentheimer_crop()
    p = default_params(nxG=96, nyG=96, nzG=96, kper=1, inlet_BC=0, outlet_BC=0, la_nu2=0.04, force_z0=200e-6, n_exclude_inlet=0,
                       n_exclude_outlet=0, initial_fluid_distribution_option=5)
    o = Oracle(p)
    o.set_walls(w); o.geometry_preprocess(); o.init_basic(); o.init_phi()
    rng = np.random.default_rng(20261018)
    o.field("phi")[...] = np.where(rng.random(o.field("phi").shape) > 0.4, -1.0, 1.0)
    o.init_pdf()
    _check(o, strict, layout, steps_strict=8, tol_fma=1e-7)


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fma"])
def test_bentheimer_crop_singlephase_case7(strict, layout):
    w = _bentheimer_crop()
    o = make_oracle(multiphase=0, nxG=96, nyG=96, nzG=96, walls_global=w, la_nu1=0.1, kper=1, force_z0=10e-6, inlet_BC=0, outlet_BC=0,
                    n_exclude_inlet=0, n_exclude_outlet=0)
    _check(o, strict, layout)


@pytest.mark.parametrize("layout", LAYOUTS)
def test_monitor_steady_phasefield(layout):
    """monitor_multiphase_steady_phasefield (MP/Monitor.F90:287-365): max |phi - phi_old| over the fluid nodes, then
    phi_old <- phi; phi_old starts as a copy of phi (MP/Init_multiphase.F90:341-347)."""
    o = make_oracle(modify_geometry_cmd=1, steady_state_option=2, ca_0=5e-3)
    ctx = ctx_from_oracle(o, kernel_variant=layout)
    ctx.upload(phi_old=o.field("phi_old"))  # what the driver uploads when steady_state_option == 2 (fortran/mflbm_iso_c.f90)
    o.color_gradient(); ctx.color_gradient()
    t = 1
    for n in (20, 30):
        _run(o, ctx, n, t)
        t += n
        umax, dphi = ctx.monitor_steady_phasefield()
        ref = o.monitor_steady_phasefield()
        assert dphi > 0 and dphi == pytest.approx(ref["d_phi_max"], rel=1e-9)
        assert umax == pytest.approx(ref["umax"], rel=1e-9)
    # a second call without stepping: nothing changed since phi_old was refreshed
    assert ctx.monitor_steady_phasefield()[1] == 0.0 == o.monitor_steady_phasefield()["d_phi_max"]
    got = ctx.download("phi_old")["phi_old"][4:-4, 4:-4, 4:-4]
    m = o.walls[2:-2, 2:-2, 2:-2] == 0
    assert np.max(np.abs(got[m] - o.field("phi_old")[4:-4, 4:-4, 4:-4][m])) <= 1e-9
    ctx.close()


def test_quiet_tiles_appear_in_the_bulk_of_a_large_droplet_box():
    """How far from an interface does |phi -+ 1| <= 1e-7 (the quiet-tile class test) actually hold?  On a 40^3 box with an
    R = 10 droplet nowhere (the tanh tails of the colour-gradient interface reach 1e-7 only ~9 cells out, and a tile needs
    its whole 27-tile neighbourhood uniform); on a 112^3 box with R = 16 the far field is quiet."""
    n, R = 112, 16.0
    p = default_params(nxG=n, nyG=n, nzG=n, kper=1, inlet_BC=0, outlet_BC=0, la_nu1=0.1, la_nu2=0.1, gamma=0.03, theta_deg=90.0,
                       n_exclude_inlet=0, n_exclude_outlet=0, initial_fluid_distribution_option=5)
    o = Oracle(p)
    o.set_walls(None); o.geometry_preprocess(); o.init_basic(); o.init_phi()
    c = (n + 1) / 2.0
    i = np.arange(-3, n + 5)
    X, Y, Z = np.meshgrid(i, i, i, indexing="ij")
    o.field("phi")[...] = np.where(np.sqrt((X - c) ** 2 + (Y - c) ** 2 + (Z - c) ** 2) <= R, 1.0, -1.0)
    o.init_pdf()
    ctx = ctx_from_oracle(o, kernel_variant=2)
    ctx.color_gradient()
    ctx.run(1, 400)
    nt, nq = ctx.tile_stats()
    phi = ctx.download("phi")["phi"][4:-4, 4:-4, 4:-4]
    ii = np.arange(1, n + 1)
    X, Y, Z = np.meshgrid(ii, ii, ii, indexing="ij")
    r = np.sqrt((X - c) ** 2 + (Y - c) ** 2 + (Z - c) ** 2)
    far = np.abs(np.abs(phi) - 1.0) <= 1e-7
    reach = float(np.abs(r[~far] - R).max())  # distance from the nominal interface of the farthest non-uniform node
    assert 6.0 < reach < 20.0, reach
    assert nq > 0.3 * nt, (nq, nt, reach)
    ctx.close()
