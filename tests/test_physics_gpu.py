"""Known-answer test of the hot path against an ANALYTIC solution (independent of the oracle): steady body-force driven
flow in a rectangular duct.  The reference carries the same series for its inlet profile
(inlet_vel_profile_rectangular, MP/Misc.F90:625-665; SP/Misc.F90 likewise):

    w(x, y) = 4 g a^2 / (nu pi^3) * sum_{n odd} (-1)^((n-1)/2) / n^3 * [1 - cosh(n pi y / a) / cosh(n pi b / 2a)] cos(n pi x / a)

for a duct of a x b fluid nodes (the node-based bounce-back of the reference puts the wall half a cell outside the last
fluid node).  6000 AA steps of singlephase_3D on 22x18x4 (20x16 fluid nodes, periodic z) reach the steady state
(viscous time a^2 / (nu pi^2) ~ 400 steps); the MRT collision + in-place streaming + bounce-back then reproduce the
series to better than 0.5 % in peak velocity and flow rate.  This pins collision, streaming and wall treatment of the
CUDA path to a truth that does not pass through our own restatement of the Fortran."""
import numpy as np
import pytest

from helpers import ctx_from_oracle, make_oracle

pytestmark = pytest.mark.gpu


def _series(nx, ny, fluid, g, nu):
    i0 = np.where(fluid.any(axis=1))[0]
    j0 = np.where(fluid.any(axis=0))[0]
    a, b = float(len(i0)), float(len(j0))
    x = np.arange(nx) - 0.5 * (i0[0] + i0[-1])
    y = np.arange(ny) - 0.5 * (j0[0] + j0[-1])
    X, Y = np.meshgrid(x, y, indexing="ij")
    u = np.zeros_like(X)
    for n in range(1, 400, 2):
        u += (-1) ** ((n - 1) // 2) / n ** 3 * (1.0 - np.cosh(n * np.pi * Y / a) / np.cosh(n * np.pi * b / (2 * a))) * np.cos(n * np.pi * X / a)
    return 4.0 * g * a * a / (nu * np.pi ** 3) * u


@pytest.mark.parametrize("layout", [pytest.param(1, id="dense"), pytest.param(2, id="sparse")])
def test_rectangular_duct_poiseuille(layout):
    nx, ny, nz, nu, g = 22, 18, 4, 0.1, 1e-6
    o = make_oracle(multiphase=0, nxG=nx, nyG=ny, nzG=nz, la_nu1=nu, kper=1, force_z0=g, n_exclude_inlet=0, n_exclude_outlet=0)
    ctx = ctx_from_oracle(o, kernel_variant=layout)
    ctx.run(1, 6000)
    ctx.sync()
    ctx.monitor()  # compute_macro_vars + reductions (valid after an even step)
    w = ctx.download("w")["w"][1:-1, 1:-1, 1:-1]
    fluid = o.walls[2:-2, 2:-2, 2:-2][:, :, 0] == 0
    ref = _series(nx, ny, fluid, g, nu)
    for k in range(nz):
        wk = w[:, :, k]
        assert abs(wk[fluid].max() / ref[fluid].max() - 1.0) < 5e-3
        assert abs(wk[fluid].sum() / ref[fluid].sum() - 1.0) < 5e-3
        assert np.max(np.abs(wk[fluid] - ref[fluid])) < 1e-2 * ref[fluid].max()
        assert np.all(wk[~fluid] == 0.0)
    ctx.close()


@pytest.mark.parametrize("layout", [pytest.param(1, id="dense"), pytest.param(2, id="sparse")])
def test_laplace_law_static_droplet(layout):
    """Multiphase path end to end on the GPU (gradient chain with quiet tiles, fused curvature, CSF force, recolouring): a
    static droplet settles to the Young-Laplace jump dp = 2 gamma / R (p = rho / 3) within 5 %, spurious currents < 1e-4."""
    from oracle.oracle import Oracle, default_params
    n, R, gamma = 40, 10.0, 0.03
    p = default_params(nxG=n, nyG=n, nzG=n, kper=1, inlet_BC=0, outlet_BC=0, la_nu1=0.1, la_nu2=0.1, gamma=gamma, theta_deg=90.0,
                       n_exclude_inlet=0, n_exclude_outlet=0, initial_fluid_distribution_option=5)
    o = Oracle(p)
    o.set_walls(None); o.geometry_preprocess(); o.init_basic(); o.init_phi()
    c = (n + 1) / 2.0
    i = np.arange(-3, n + 5)
    X, Y, Z = np.meshgrid(i, i, i, indexing="ij")
    o.field("phi")[...] = np.where(np.sqrt((X - c) ** 2 + (Y - c) ** 2 + (Z - c) ** 2) <= R, 1.0, -1.0)
    o.init_pdf()
    ctx = ctx_from_oracle(o, kernel_variant=layout)
    ctx.color_gradient()
    ctx.run(1, 3000)
    ctx.compute_macro_vars()
    got = ctx.download("rho", "phi", "u", "v", "w")
    rho = got["rho"][1:-1, 1:-1, 1:-1]
    phi = got["phi"][4:-4, 4:-4, 4:-4]
    ii = np.arange(1, n + 1)
    X, Y, Z = np.meshgrid(ii, ii, ii, indexing="ij")
    r = np.sqrt((X - c) ** 2 + (Y - c) ** 2 + (Z - c) ** 2)
    dp = (rho[r < R - 4].mean() - rho[(r > R + 4) & (r < R + 8)].mean()) / 3.0
    r_eff = (3.0 * (0.5 * (1.0 + phi))[r < R + 6].sum() / (4.0 * np.pi)) ** (1.0 / 3.0)
    assert abs(dp / (2.0 * gamma / r_eff) - 1.0) < 0.05, (dp, r_eff)
    assert np.sqrt((got["u"] ** 2 + got["v"] ** 2 + got["w"] ** 2).max()) < 1e-4
    ctx.close()


@pytest.mark.parametrize("theta", [60.0, 120.0])
def test_capillary_tube_young_laplace_with_contact_angle(theta):
    """GPU twin of tests/test_oracle.py::test_capillary_tube_young_laplace_with_contact_angle (wetting model end to end):
    dp = 2 gamma cos(theta) / R across the menisci of a slug in a z-periodic capillary, within 8 %."""
    from oracle.oracle import Oracle, default_params
    nx = ny = 34
    nz, R, gamma = 72, 13.0, 0.03
    i = np.arange(1, nx + 1)
    c = (nx + 1) / 2.0
    X, Y = np.meshgrid(i, i, indexing="ij")
    rr = np.sqrt((X - c) ** 2 + (Y - c) ** 2)
    wg = np.repeat((rr > R).astype(np.int8)[:, :, None], nz, axis=2)
    p = default_params(nxG=nx, nyG=ny, nzG=nz, kper=1, inlet_BC=0, outlet_BC=0, la_nu1=0.1, la_nu2=0.1, gamma=gamma, theta_deg=theta,
                       n_exclude_inlet=0, n_exclude_outlet=0, initial_fluid_distribution_option=5)
    o = Oracle(p)
    o.set_walls(wg); o.geometry_preprocess(); o.init_basic(); o.init_phi()
    kk = np.arange(-3, nz + 5)
    o.field("phi")[...] = np.where(((kk >= nz // 4 + 1) & (kk <= 3 * nz // 4))[None, None, :], 1.0, -1.0)
    o.init_pdf()
    ctx = ctx_from_oracle(o, kernel_variant=2)
    ctx.color_gradient()
    ctx.run(1, 5000)
    ctx.compute_macro_vars()
    rho = ctx.download("rho")["rho"][1:-1, 1:-1, 1:-1]
    core = rr < R - 4
    p_in = rho[core][:, nz // 2 - 3:nz // 2 + 3].mean() / 3.0
    p_out = np.concatenate([rho[core][:, :4], rho[core][:, -4:]], axis=1).mean() / 3.0
    expect = 2.0 * gamma * np.cos(np.radians(theta)) / R
    assert abs((p_in - p_out) / expect - 1.0) < 0.08, (p_in - p_out, expect)
    ctx.close()
