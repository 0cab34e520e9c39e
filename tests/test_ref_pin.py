"""Pins the CPU oracle to the REFERENCE'S OWN SOURCE.

The image has no Fortran compiler, so oracle/_ref is built by translating the reference's hot-path subroutines
statement by statement (oracle/f2c_lite.py; the translator knows nothing about lattice Boltzmann, it maps a small
Fortran-90 subset onto C with fully parenthesised expressions) from /root/reference and compiling the result with
gcc -ffp-contract=off.  Here the hand-written restatement (oracle/mflbm_oracle.c, same flags) and the translated
reference start from the same state and must agree BIT FOR BIT, ghost layers included, after every time step:
main_iteration_kernel = kernel_{odd,even}_color + inlet / outlet / porous-plate routines + color_gradient
(MP/Main_multiphase.F90:341-486), and the init / monitor routines that translate.  Skipped when neither
oracle/_ref/*.so nor /root/reference exists.
"""
import numpy as np
import pytest

from helpers import make_oracle
from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built and /root/reference absent")

from ref_helpers import assert_same_state, copy_state, ref_from_oracle  # noqa: E402


def _lockstep(o, steps, check_every=1):
    r = ref_from_oracle(o)
    copy_state(r, o)
    if o.mp:
        r.call("color_gradient")
        o.color_gradient()
        assert_same_state(r, o, "color_gradient")
    f_before = o.f(5).copy()
    for t in range(1, steps + 1):
        r.set(ntime=t)
        r.call("main_iteration_kernel")
        o.step(t)
        if t % check_every == 0 or t == steps:
            assert_same_state(r, o, "step %d" % t)
    assert not np.array_equal(f_before, o.f(5)), "the run did not change the state: the comparison would be vacuous"
    return r


def test_c1_tube_sphere_drainage_200_steps():
    """BASELINE configs[0]: velocity inlet, convective outlet, wetting (theta = 30), CSF force, shipped MRT rates."""
    o = make_oracle(modify_geometry_cmd=1)
    r = _lockstep(o, 200, check_every=10)
    r.call("cal_saturation")  # MP/Monitor.F90:512-550
    s = o.cal_saturation()
    # OpenMP sum reductions on both sides: same terms, unspecified order
    assert r.get("vol1_sum") == pytest.approx(s["vol1_sum"], rel=1e-12) and r.get("vol2_sum") == pytest.approx(s["vol2_sum"], rel=1e-12)
    assert r.get("saturation_full_domain") == pytest.approx(s["saturation_full_domain"], rel=1e-12)
    r.call("monitor_breakthrough")  # MP/Monitor.F90:472-507
    assert int(r.get("outlet_phase1_sum")) == o.monitor_breakthrough()["outlet_phase1_sum"]
    r.call("compute_macro_vars")  # MP/Misc.F90:372-430
    o.compute_macro_vars()
    for n in ("u", "v", "w", "rho"):  # compute_macro_vars covers 1..n; the ghost layer keeps whatever the initialisation left
        assert np.array_equal(r.array(n)[1:-1, 1:-1, 1:-1], o.field(n)[1:-1, 1:-1, 1:-1]), n
    assert np.array_equal(r.array("phi"), o.field("phi"))  # phi <- 0 at the walls of 1..n (MP/Misc.F90:424)


@pytest.mark.parametrize("inlet,outlet", [(2, 2), (1, 2), (2, 1)])
def test_pressure_boundaries(inlet, outlet):
    o = make_oracle(modify_geometry_cmd=1, inlet_BC=inlet, outlet_BC=outlet, force_z0=1e-5, sa_inject=0.8)
    _lockstep(o, 12)


@pytest.mark.parametrize("plate", [1, 2])
def test_porous_plate(plate):
    o = make_oracle(modify_geometry_cmd=0, porous_plate_cmd=plate, Z_porous_plate=40)
    _lockstep(o, 12)


def test_random_porous_medium_other_viscosity_ratio_and_angle():
    rng = np.random.default_rng(11)
    wg = (rng.random((28, 24, 36)) < 0.3).astype(np.int8)
    wg[:, :, :4] = 0
    wg[:, :, -4:] = 0
    o = make_oracle(nxG=28, nyG=24, nzG=36, walls_global=wg, la_nu1=0.02, la_nu2=0.1, theta_deg=120.0, interface_z0=12.0,
                    n_exclude_inlet=0, n_exclude_outlet=0, ca_0=1e-3)
    _lockstep(o, 20)


def test_initial_populations_and_phase_field():
    """initialization_new_multi (options 1-5) + initialization_new_multi_pdf (MP/Init_multiphase.F90:243-470)"""
    for opt, z0 in ((1, 8.0), (2, 11.0), (3, 9.0), (4, 9.0), (5, 7.5)):
        o = make_oracle(modify_geometry_cmd=1, initial_fluid_distribution_option=opt, interface_z0=z0)
        r = ref_from_oracle(o)
        r.call("initialization_new_multi")  # fills phi, then calls initialization_new_multi_pdf
        assert np.array_equal(r.array("phi"), o.field("phi")), opt
        for q in range(19):
            assert np.array_equal(r.array("f%d" % q), o.f(q)) and np.array_equal(r.array("g%d" % q), o.g(q)), (opt, q)
        for n in ("f_convec_bc", "g_convec_bc", "phi_convec_bc"):
            assert np.array_equal(r.array(n), o.field(n)), (opt, n)


def test_inlet_velocity_profile():
    """initialization_open_velocity_inlet_BC + inlet_vel_profile_rectangular (MP/Init_multiphase.F90:194-240, MP/Misc.F90:625-665)"""
    o = make_oracle(modify_geometry_cmd=1)
    r = ref_from_oracle(o)
    r.set(ca_0=o.p.ca_0, la_nu1=o.p.la_nu1, a_xy=o.get_double("A_xy"), la_x=o.get_double("la_x"), la_y=o.get_double("la_y"),
          pore_sum=o.get_i64("pore_sum"), target_inject_pore_volume=o.p.target_inject_pore_volume, d_vol_monitor=0.01)
    r.call("initialization_open_velocity_inlet_bc")
    assert np.array_equal(r.array("w_in"), o.field("w_in"))
    assert r.get("uin_avg") == o.get_double("uin_avg") and r.get("flowrate") == o.get_double("flowrate")


def test_modify_geometry_tube_sphere():
    """modify_geometry (MP/Misc.F90:213-244) produces the shipped tube + sphere: 73 936 fluid nodes"""
    o = make_oracle(modify_geometry_cmd=1)
    r = ref_from_oracle(o)
    r.array("walls_global")[...] = 0
    r.call("modify_geometry")
    w = r.array("walls_global").copy()
    assert 0 < int((w == 1).sum()) < w.size
    w[0, :, :] = w[-1, :, :] = w[:, 0, :] = w[:, -1, :] = 1  # set_walls then closes the x / y domain boundaries (MP/Misc.F90:40-75)
    assert np.array_equal(w, o.walls_global)
    assert int((w == 0).sum()) == 73936


@pytest.mark.parametrize("cfg", [dict(inlet_BC=1, outlet_BC=1, Re=0.5, char_length=38.0), dict(inlet_BC=2, outlet_BC=2, rho_drop=1e-3),
                                 dict(inlet_BC=1, outlet_BC=2, Re=0.5, char_length=38.0)])
@pytest.mark.parametrize("preset", [1, 2, 3])
def test_singlephase_steps(cfg, preset):
    """singlephase_3D: kernel_odd / kernel_even + its inlet / outlet routines (SP/Kernel.F90, SP/Boundary.F90, SP/Main.F90:291-422)"""
    rng = np.random.default_rng(7)
    wg = (rng.random((30, 26, 40)) < 0.25).astype(np.int8)
    wg[:, :, :4] = 0
    wg[:, :, -4:] = 0
    o = make_oracle(multiphase=0, nxG=30, nyG=26, nzG=40, la_nu1=0.1, walls_global=wg, n_exclude_inlet=0, n_exclude_outlet=0,
                    mrt_para_preset=preset, force_z0=1e-6, **cfg)
    r = _lockstep(o, 16)
    r.call("compute_macro_vars")  # SP/Misc.F90:368-423
    o.compute_macro_vars()
    for n in ("u", "v", "w", "rho"):
        assert np.array_equal(r.array(n)[1:-1, 1:-1, 1:-1], o.field(n)[1:-1, 1:-1, 1:-1]), n
