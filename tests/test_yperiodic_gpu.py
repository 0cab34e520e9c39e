"""y-periodic domains (first half; MP/Mpi.F90:22-40, :147-207, :633-790): one z slab, the y wrap of the
populations in the adjacency of the sparse layout, the phi ghost rows by k_wrap_y_phi.  Strict build against the oracle,
whose own y exchange is pinned by the tiled-lattice property (tests/test_oracle.py::test_y_periodic_step_equals...)."""
import numpy as np
import pytest

import mflbm_b200 as M
from helpers import compare_state, ctx_from_oracle, make_oracle
from oracle.oracle import Oracle, default_params

pytestmark = pytest.mark.gpu


def _run(o, ctx, n, t0=1):
    for t in range(t0, t0 + n):
        o.step(t)
    ctx.run(t0, n)
    ctx.sync()


def _seam_case(kper, mp=True):
    nx, ny, nz = 20, 24, 28
    rng = np.random.default_rng(5 + kper)
    wg = (rng.random((nx, ny, nz)) < 0.2).astype(np.int8)
    if not kper:
        wg[:, :, :4] = 0
        wg[:, :, -4:] = 0
    p = default_params(multiphase=1 if mp else 0, nxG=nx, nyG=ny, nzG=nz, jper=1, kper=kper, wsy0=0, wsy1=0, inlet_BC=0 if kper else 1,
                       outlet_BC=0 if kper else 1, force_z0=2e-4 if kper else 0.0, la_nu1=0.004 if mp else 0.1, la_nu2=0.04,
                       n_exclude_inlet=0, n_exclude_outlet=0, initial_fluid_distribution_option=5, interface_z0=6.0,
                       Re=0.5, char_length=18.0)
    o = Oracle(p)
    o.set_walls(wg)
    if mp:
        o.geometry_preprocess()
    o.init_basic(); o.init_phi()
    if mp:  # a drop across the y seam (and the z seam when that is periodic too)
        i = np.arange(-3, nx + 5)[:, None, None]
        j = np.arange(-3, ny + 5)[None, :, None]
        k = np.arange(-3, nz + 5)[None, None, :]
        dy = np.minimum((j - 1.0) % ny, ny - (j - 1.0) % ny)
        dz = np.minimum((k - 2.0) % nz, nz - (k - 2.0) % nz) if kper else np.abs(k - 9.0)
        o.field("phi")[...] = np.where((i - 10.5) ** 2 + dy ** 2 + dz ** 2 <= 49.0, 1.0, -1.0)
    o.init_pdf()
    return o


@pytest.mark.parametrize("kper", [1, 0], ids=["yz-periodic", "y-periodic-open-z"])
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fma"])
def test_multiphase_across_the_y_seam(strict, kper):
    o = _seam_case(kper)
    ctx = ctx_from_oracle(o, strict=strict)  # kernel_variant 0: jper forces the sparse layout
    o.color_gradient(); ctx.color_gradient()
    # FMA build: the drop starts as a sharp sphere across both seams; 1.4e-11 on a handful of normals where |grad phi| is
    # close to the cut-off, so 1e-10 here
    compare_state(ctx, o, 0.0 if strict else 1e-10, sparse=True)
    if strict:
        t = 1
        for n in (1, 1, 10):
            _run(o, ctx, n, t)
            t += n
            compare_state(ctx, o, 0.0, sparse=True)
    else:
        _run(o, ctx, 2)
        compare_state(ctx, o, 1e-10, sparse=True)
    ctx.close()


@pytest.mark.parametrize("kper", [1, 0], ids=["yz-periodic", "y-periodic-open-z"])
def test_singlephase_y_periodic(kper):
    o = _seam_case(kper, mp=False)
    ctx = ctx_from_oracle(o, strict=True)
    _run(o, ctx, 9)
    compare_state(ctx, o, 0.0, sparse=True)
    ctx.close()


def test_y_periodic_with_the_porous_plate_is_refused():
    with pytest.raises(M.MflbmError, match="jper"):
        M.Context(solver=1, nx=8, ny=8, nz=8, jper=1, porous_plate_cmd=1, Z_porous_plate=4)


def test_reference_case1_drop_attached_to_the_wall():
    """test_suites/3D_simulation/1.drop_attached_wall: 50x80x80, y and z periodic, fluid-1 drop of radius 18 on the x = 1 wall,
    contact angle 45 degrees (measured through fluid 2, as the control file defines it).  GPU == oracle bit for bit over the
    first steps; then the drop relaxes and the angle of its spherical cap (from its height h and base radius r on the wall,
    tan(theta_1 / 2) = h / r, theta = 180 - theta_1) settles within 8 degrees of the prescribed one (the tolerance class of the
    capillary-tube test: the wetting model is first order at the wall)."""
    nx, ny, nz = 50, 80, 80
    p = default_params(nxG=nx, nyG=ny, nzG=nz, jper=1, kper=1, wsy0=0, wsy1=0, inlet_BC=0, outlet_BC=0, theta_deg=45.0,
                       initial_fluid_distribution_option=3, interface_z0=18.0, n_exclude_inlet=0, n_exclude_outlet=0,
                       steady_state_option=1)
    o = Oracle(p)
    o.setup(None)
    ctx = ctx_from_oracle(o, strict=True)
    o.color_gradient(); ctx.color_gradient()
    _run(o, ctx, 6)
    compare_state(ctx, o, 0.0, sparse=True)
    ctx.run(7, 20000)
    phi = ctx.download("phi")["phi"][4:-4, 4:-4, 4:-4]
    vol = 0.5 * (1.0 + phi)
    vol[o.walls[2:-2, 2:-2, 2:-2] != 0] = 0.0
    V = float(vol.sum())
    h = float(vol[:, ny // 2 - 1:ny // 2 + 1, nz // 2 - 1:nz // 2 + 1].mean(axis=(1, 2)).sum())  # drop height along x through its axis
    # spherical cap: V = pi h (3 r^2 + h^2) / 6  ->  base radius r from V and h
    r = np.sqrt(max((6.0 * V / (np.pi * h) - h * h) / 3.0, 1e-12))
    theta1 = 2.0 * np.degrees(np.arctan(h / r))      # angle through fluid 1 (the drop)
    theta = 180.0 - theta1                           # through fluid 2, the control file's convention (MP/IO_multiphase.F90:467-468)
    assert abs(theta - 45.0) < 8.0, (theta, h, r, V)
    ctx.close()
