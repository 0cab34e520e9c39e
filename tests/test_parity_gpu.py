"""GPU parity: CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star / BASELINE.md section 4): FP64 fields <= 1e-12 relative after one
odd + one even step in the default (FMA) build; bit-exact in the -fmad=false strict build, where the
kernels evaluate the reference's expressions in source order and no libm call is involved on device.
"""
import numpy as np
import pytest

from helpers import compare_state, ctx_from_oracle, make_oracle

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-12   # FP64 relative, after one odd + one even step (FMA build)
TOL_10 = 1e-9      # after ten steps: rounding differences are amplified through the normalised colour gradient
LAYOUTS = [pytest.param(1, id="dense"), pytest.param(2, id="sparse")]


def _run_both(o, ctx, nsteps, ntime0=1):
    for n in range(ntime0, ntime0 + nsteps):
        o.step(n)
    ctx.run(ntime0, nsteps)
    ctx.sync()


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fma"])
def test_c1_tube_sphere_multiphase_steps(strict, layout):
    """C1: 40x40x60 tube+sphere drainage, velocity inlet / convective outlet (BASELINE configs[0]).
    strict (-fmad=false) build: bit-exact; default build: 1e-12 after 1 odd + 1 even step."""
    o = make_oracle(modify_geometry_cmd=1)
    ctx = ctx_from_oracle(o, strict=strict, kernel_variant=layout)
    sp = layout == 2
    o.color_gradient()
    ctx.color_gradient()
    compare_state(ctx, o, 0.0 if strict else TOL_STEP, sparse=sp)
    t = 1
    for nsteps, tol in ((1, TOL_STEP), (1, TOL_STEP), (8, TOL_10)):  # one odd step, one even step, then eight more
        _run_both(o, ctx, nsteps, t)
        t += nsteps
        compare_state(ctx, o, 0.0 if strict else tol, sparse=sp)
    ctx.close()


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("inlet,outlet", [(2, 2), (1, 2)])
def test_multiphase_pressure_bcs(inlet, outlet, layout):
    o = make_oracle(modify_geometry_cmd=1, inlet_BC=inlet, outlet_BC=outlet, force_z0=1e-5, sa_inject=0.8)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=layout)
    o.color_gradient(); ctx.color_gradient()
    _run_both(o, ctx, 6)
    compare_state(ctx, o, 0.0, sparse=layout == 2)
    ctx.close()


@pytest.mark.parametrize("layout", LAYOUTS)
def test_multiphase_periodic_z_bodyforce(layout):
    """z-periodic, body-force driven (the steady fractional-flow setup of test-suite case 5/6)."""
    rng = np.random.default_rng(5)
    wg = (rng.random((24, 20, 32)) < 0.2).astype(np.int8)
    o = make_oracle(nxG=24, nyG=20, nzG=32, kper=1, force_z0=2e-4, n_exclude_inlet=0, n_exclude_outlet=0,
                    initial_fluid_distribution_option=5, interface_z0=6.0, walls_global=wg, la_nu2=0.04)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=layout)
    o.color_gradient(); ctx.color_gradient()
    _run_both(o, ctx, 7)
    compare_state(ctx, o, 0.0, sparse=layout == 2)
    ctx.close()


@pytest.mark.parametrize("mrt", [1, 3, 4])
def test_multiphase_mrt_variants(mrt):
    o = make_oracle(modify_geometry_cmd=1, mrt=mrt)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=2)
    o.color_gradient(); ctx.color_gradient()
    _run_both(o, ctx, 4)
    compare_state(ctx, o, 0.0, sparse=True)
    ctx.close()


@pytest.mark.parametrize("plate", [1, 2])
def test_porous_plate(plate):
    o = make_oracle(modify_geometry_cmd=0, porous_plate_cmd=plate, Z_porous_plate=40)
    ctx = ctx_from_oracle(o, strict=True)
    o.color_gradient(); ctx.color_gradient()
    _run_both(o, ctx, 6)
    compare_state(ctx, o, 0.0)
    ctx.close()


@pytest.mark.parametrize("cfg", [dict(kper=1, force_z0=1e-5), dict(inlet_BC=1, outlet_BC=1, Re=0.5, char_length=38.0),
                                 dict(inlet_BC=2, outlet_BC=2, rho_drop=1e-3)])
@pytest.mark.parametrize("layout", LAYOUTS)
def test_singlephase_steps(cfg, layout):
    rng = np.random.default_rng(7)
    wg = (rng.random((30, 26, 40)) < 0.25).astype(np.int8)
    wg[:, :, :4] = 0
    wg[:, :, -4:] = 0
    o = make_oracle(multiphase=0, nxG=30, nyG=26, nzG=40, la_nu1=0.1, walls_global=wg, n_exclude_inlet=0,
                    n_exclude_outlet=0, **cfg)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=layout)
    _run_both(o, ctx, 9)
    compare_state(ctx, o, 0.0, sparse=layout == 2)
    ctx.close()


@pytest.mark.parametrize("layout", LAYOUTS)
def test_monitors_match_oracle(layout):
    o = make_oracle(modify_geometry_cmd=1)
    ctx = ctx_from_oracle(o, kernel_variant=layout)
    o.color_gradient(); ctx.color_gradient()
    v1, v2 = ctx.cal_saturation()
    s = o.cal_saturation()
    assert abs(v1 - s["vol1_sum"]) <= 1e-10 * abs(s["vol1_sum"]) and abs(v2 - s["vol2_sum"]) <= 1e-10 * abs(s["vol2_sum"])
    _run_both(o, ctx, 40)
    m = ctx.monitor()
    mo = o.monitor()
    for name in ("fl1", "fl2", "vol1", "vol2", "mass1", "mass2", "pre"):
        ref = o.field(name)
        assert np.max(np.abs(m[name] - ref)) <= 1e-10 * max(1e-30, np.max(np.abs(ref))), name
    assert m["umax"] == pytest.approx(mo["umax"], rel=1e-11)
    assert m["usq1"] == pytest.approx(mo["usq1"], rel=1e-10)
    assert m["usq2"] == pytest.approx(mo["usq2"], rel=1e-10)
    got = ctx.download("u", "v", "w", "rho", "phi")
    for n in ("u", "v", "w", "rho", "phi"):
        g_ = 3 if n == "phi" else 0  # compute_macro_vars covers 1..n only (ghost rho keeps the driver's init value)
        a = got[n][1 + g_:-1 - g_, 1 + g_:-1 - g_, 1 + g_:-1 - g_]
        ref = o.field(n)[1 + g_:-1 - g_, 1 + g_:-1 - g_, 1 + g_:-1 - g_]
        assert np.max(np.abs(a - ref)) <= 1e-12 * np.max(np.abs(ref)), n
    assert ctx.monitor_breakthrough() == o.monitor_breakthrough()["outlet_phase1_sum"]
    c = ctx.monitor_steady_capillarypressure()
    co = o.monitor_steady_capillarypressure()
    assert c["i_w"] == co["i_w"] and c["i_nw"] == co["i_nw"]
    assert c["pre_w"] == pytest.approx(co["pre_w"], rel=1e-11) and c["pre_nw"] == pytest.approx(co["pre_nw"], rel=1e-11)
    ctx.close()


@pytest.mark.parametrize("ca", [1e-4, 5e-3])
def test_quiet_tiles_long_run_bit_exact(ca):
    """Sparse layout, strict build, long run: the quiet-tile skipping of the colour-gradient chain (DESIGN.md "Quiet
    tiles") must stay bit-exact while the interface moves through tiles that were quiet, across monitor calls
    (compute_macro_vars zeroes phi at walls) and across downloads (which refresh the skipped solid-node phi)."""
    o = make_oracle(modify_geometry_cmd=1, ca_0=ca)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=2)
    o.color_gradient(); ctx.color_gradient()
    t = 1
    seen_quiet = 0
    for nsteps, mon in ((50, False), (150, True), (200, False), (200, True)):
        _run_both(o, ctx, nsteps, t)
        t += nsteps
        nt, nq = ctx.tile_stats()
        seen_quiet = max(seen_quiet, nq)
        assert nt > 0
        if mon:
            m, mo = ctx.monitor(), o.monitor()
            assert m["umax"] == pytest.approx(mo["umax"], rel=1e-11)
        compare_state(ctx, o, 0.0, sparse=True)
    assert seen_quiet > 0, "the run never had a quiet tile: the skipping path was not exercised"
    ctx.close()


@pytest.mark.parametrize("force", [None, "30:31", "0:3"], ids=["guessed", "missed", "partly"])
def test_speculative_early_chain_is_exact(force, monkeypatch):
    """Speculative step schedule (DESIGN.md "Step schedule"): the gradient chain of the active tiles runs beside the
    collision of the far planes.  A long duct so that the planner switches it on; strict build, bit-exact against the oracle
    whether the guessed layer range is right, completely wrong (every active tile is "missed" and the chain re-runs on the
    full lists) or partly right."""
    monkeypatch.setenv("MFLBM_SPEC", "1")  # measured: no gain on B200, so the schedule is opt-in; it must still be exact
    if force:
        monkeypatch.setenv("MFLBM_SPEC_FORCE", force)
    rng = np.random.default_rng(23)
    nz = 160
    wg = (rng.random((24, 22, nz)) < 0.22).astype(np.int8)
    wg[:, :, :6] = 0
    wg[:, :, -6:] = 0
    o = make_oracle(nxG=24, nyG=22, nzG=nz, walls_global=wg, la_nu2=0.04, interface_z0=7.0, ca_0=2e-3, n_exclude_inlet=0,
                    n_exclude_outlet=0)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=2)
    o.color_gradient(); ctx.color_gradient()
    t = 1
    for nsteps in (3, 40, 61):
        _run_both(o, ctx, nsteps, t)
        t += nsteps
        compare_state(ctx, o, 0.0, sparse=True)
    assert ctx.spec_steps > 50, ctx.spec_steps
    nt, nq = ctx.tile_stats()
    assert nq > 0
    ctx.close()
