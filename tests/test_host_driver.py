"""CPU tests of the host driver mirror (C++): its init-time routines must reproduce the oracle's restatement of
set_walls / geometry_preprocessing_new / initialization_* -- integer work bit-exactly, FP64 fields bit-exactly as
well (same expressions, same compiler flags)."""
import numpy as np
import pytest

import mflbm_b200 as M
from helpers import make_oracle


def _driver(tmp_path, multiphase=True, walls=None, idz=0, **kw):
    M.build()
    M.build_host()
    ctl = M.write_control_file(str(tmp_path / "simulation_control.txt"), multiphase=multiphase, **kw)
    d = M.Driver(ctl, idz=idz, walls=walls)
    d.setup()
    return d


def _same_lists(d, o):
    sn, so = d.solid_nodes(), o.solid_nodes()
    fn, fo = d.fluid_nodes(), o.fluid_nodes()
    assert len(sn) == len(so) and len(fn) == len(fo)
    for name in ("ix", "iy", "iz", "i_fluid_num", "neighbor_list", "la_weight"):
        assert np.array_equal(sn[name], so[name]), name
    for name in ("ix", "iy", "iz", "nwx", "nwy", "nwz", "theta"):
        assert np.array_equal(fn[name], fo[name]), name


def test_c1_setup_matches_oracle(tmp_path):
    d = _driver(tmp_path, modify_geometry_cmd=1, breakthrough_check=1)
    o = make_oracle(modify_geometry_cmd=1)
    assert d.i64("pore_sum") == 73936 and d.i64("pore_sum_effective") == 45056 and d.i64("ntime_max") == 68270
    assert d.i64("ntime_monitor") == 682  # SURVEY 8c known answer
    assert np.array_equal(d.walls, o.walls)
    _same_lists(d, o)
    for n in ("la_nui1", "la_nui2", "theta", "phi_inlet", "force_Z", "uin_avg", "flowrate", "A_xy", "A_xy_effective", "rho_in"):
        assert d.f64(n) == o.get_double(n), n
    assert np.array_equal(d.field("w_in"), o.field("w_in"))
    assert np.array_equal(d.field("phi"), o.field("phi"))
    for q in range(19):
        assert np.array_equal(d.field("f", q), o.f(q)) and np.array_equal(d.field("g", q), o.g(q))
    assert np.array_equal(d.field("f_convec_bc"), o.field("f_convec_bc"))
    assert np.array_equal(d.field("phi_convec_bc"), o.field("phi_convec_bc"))


def test_slab_lists_match_oracle(tmp_path):
    """z-slab slicing of the boundary-node lists and wall ghost layers (idz of npz), periodic and open."""
    rng = np.random.default_rng(11)
    wg = (rng.random((18, 16, 32)) < 0.25).astype(np.int8)
    for kper, idz in ((1, 0), (1, 3), (0, 0), (0, 2), (0, 3)):
        zs = "0,0"
        d = _driver(tmp_path, walls=wg, idz=idz, lattice_dimensions="18,16,32", MPI_process_num="1,1,4",
                    periodic_indicator="0,0,%d" % kper, domain_wall_status_z=zs, excluded_layers="0,0",
                    inlet_BC=1 if not kper else 0, outlet_BC=1 if not kper else 0, body_force_0="1d-5")
        o = make_oracle(nxG=18, nyG=16, nzG=32, npz=4, idz=idz, kper=kper, walls_global=wg, n_exclude_inlet=0,
                        n_exclude_outlet=0, force_z0=1e-5, inlet_BC=1 if not kper else 0, outlet_BC=1 if not kper else 0)
        assert np.array_equal(d.walls, o.walls), (kper, idz)
        _same_lists(d, o)
        assert np.array_equal(d.field("phi"), o.field("phi"))
        d.close()


def test_singlephase_setup_matches_oracle(tmp_path):
    rng = np.random.default_rng(2)
    wg = (rng.random((20, 20, 30)) < 0.3).astype(np.int8)
    d = _driver(tmp_path, multiphase=False, walls=wg, lattice_dimensions="20,20,30", fluid_viscosity=0.1,
                body_force_0="1d-5", MRT_collision_parameter_preset=2)
    o = make_oracle(multiphase=0, nxG=20, nyG=20, nzG=30, la_nu1=0.1, kper=1, force_z0=1e-5, walls_global=wg,
                    mrt_para_preset=2)
    assert np.array_equal(d.walls, o.walls)
    for n in ("s_e", "s_e2", "s_q", "s_nu", "s_pi", "s_t", "force_Z"):
        assert d.f64(n) == o.get_double(n), n
    for q in range(19):
        assert np.array_equal(d.field("f", q), o.f(q))
    assert d.i64("pore_sum") == o.get_i64("pore_sum")


def test_control_file_errors_mirror_reference(tmp_path):
    M.build(); M.build_host()
    for bad, msg in ((dict(periodic_indicator="1,0,0"), "X direction periodic"),
                     (dict(MPI_process_num="2,1,1"), "MPI_process_num_X"),
                     (dict(inlet_BC=2, outlet_BC=1), "Inlet pressure \\+ outlet convective"),
                     (dict(periodic_indicator="0,0,1", domain_wall_status_z="1,1"), "z = zmin")):
        ctl = M.write_control_file(str(tmp_path / "c.txt"), **bad)
        with pytest.raises(M.MflbmError, match=msg):
            M.Driver(ctl)


def test_wall_array_file_roundtrip(tmp_path):
    import os
    from conftest import ROOT
    w = np.load(os.path.join(ROOT, "tests", "golden", "tube_sphere.npz"))["walls"]
    path = M.write_wall_array(str(tmp_path / "walls.dat"), w)
    M.build(); M.build_host()
    ctl = M.write_control_file(str(tmp_path / "c.txt"), lattice_dimensions="60,60,80", external_geometry_read_cmd=1,
                               excluded_layers="5,5")
    d = M.Driver(ctl, wall_file=path)
    d.setup()
    assert d.i64("pore_sum") == 229816


def test_windowed_geometry_matches_whole_lattice(tmp_path):
    """Each rank may hold only a z window of the wall array (slab + 12 planes): identical slab results."""
    rng = np.random.default_rng(4)
    nzG, npz = 96, 4
    wg = (rng.random((14, 12, nzG)) < 0.3).astype(np.int8)
    for kper in (0, 1):
        for idz in range(npz):
            kw = dict(lattice_dimensions="14,12,%d" % nzG, MPI_process_num="1,1,%d" % npz, periodic_indicator="0,0,%d" % kper,
                      excluded_layers="0,0", inlet_BC=0 if kper else 1, outlet_BC=0 if kper else 1, body_force_0="1d-5")
            k0, k1 = M.Driver.window_range(idz, npz, nzG, kper)
            planes = [(k - 1) % nzG for k in range(k0, k1 + 1)]
            M.build(); M.build_host()
            ctl = M.write_control_file(str(tmp_path / "c.txt"), **kw)
            dw = M.Driver(ctl, idz=idz, walls_window=(wg[:, :, planes], k0))
            dw.setup()
            o = make_oracle(nxG=14, nyG=12, nzG=nzG, npz=npz, idz=idz, kper=kper, walls_global=wg, n_exclude_inlet=0,
                            n_exclude_outlet=0, force_z0=1e-5, inlet_BC=0 if kper else 1, outlet_BC=0 if kper else 1)
            assert np.array_equal(dw.walls, o.walls), (kper, idz)
            _same_lists(dw, o)
            assert dw.i64("pore_sum_local") == int((o.walls[2:-2, 2:-2, 2:-2] == 0).sum())
            dw.close()


@pytest.mark.parametrize("kper", [0, 1])
def test_windowed_modify_geometry_matches_whole_lattice(tmp_path, kper):
    """modify_geometry_cmd = 1 on a z window: the tube + sphere is carved into every held plane, the wrapped ghost planes of a
    z-periodic window included (the reference modifies the whole global array before distributing it, MP/Misc.F90:213-244),
    so that walls and boundary-node lists agree across the periodic seam."""
    nzG, npz = 96, 4
    for idz in range(npz):
        kw = dict(lattice_dimensions="24,24,%d" % nzG, MPI_process_num="1,1,%d" % npz, periodic_indicator="0,0,%d" % kper,
                  excluded_layers="0,0", inlet_BC=0 if kper else 1, outlet_BC=0 if kper else 1, body_force_0="1d-5", modify_geometry_cmd=1)
        ctl = M.write_control_file(str(tmp_path / "c.txt"), **kw)
        k0, k1 = M.Driver.window_range(idz, npz, nzG, kper)
        dw = M.Driver(ctl, idz=idz, walls_window=(np.zeros((24, 24, k1 - k0 + 1), np.int8), k0))
        dw.setup()
        o = make_oracle(nxG=24, nyG=24, nzG=nzG, npz=npz, idz=idz, kper=kper, n_exclude_inlet=0, n_exclude_outlet=0, force_z0=1e-5,
                        inlet_BC=0 if kper else 1, outlet_BC=0 if kper else 1, modify_geometry_cmd=1)
        assert np.array_equal(dw.walls, o.walls), (kper, idz)
        _same_lists(dw, o)
        dw.close()


def test_too_small_wall_window_is_refused(tmp_path):
    ctl = M.write_control_file(str(tmp_path / "c.txt"), lattice_dimensions="14,12,96", MPI_process_num="1,1,4")
    with pytest.raises(M.MflbmError, match="wall window too small"):
        M.Driver(ctl, idz=1, walls_window=(np.zeros((14, 12, 30), np.int8), 22))  # slab 25..48 needs planes 15..58


@pytest.mark.parametrize("kper", [0, 1])
def test_windowed_wall_file_read_equals_the_whole_file_path(tmp_path, kper):
    """every rank reads only the planes around its slab from the reference wall-array file (seek, no
    whole-lattice array, no broadcast) and ends up with exactly the local walls, pore counts and boundary-node lists that
    the whole-file path produces.  The sample is smaller than the lattice in x / y (the reference pads it with solid)."""
    from importlib import import_module
    geo = import_module("mflbm_b200.geometry")
    nxs, nys, nzG, npz = 36, 30, 96, 4
    sample = geo.sphere_pack(nxs, nys, nzG, periodic=bool(kper), porosity=0.5, rmin=3.0, rmax=7.0, seed=31, buffer=0 if kper else 4)
    wf = M.write_wall_array(str(tmp_path / "walls.dat"), sample)
    over = dict(lattice_dimensions="40,34,%d" % nzG, MPI_process_num="1,1,%d" % npz, external_geometry_read_cmd=1,
                excluded_layers="0,0")
    if kper:
        over.update(periodic_indicator="0,0,1", inlet_BC=0, outlet_BC=0)
    ctl = M.write_control_file(str(tmp_path / "ctl.txt"), multiphase=True, **over)
    for idz in range(npz):
        a = M.Driver(ctl, idz=idz, wall_file=wf)
        b = M.Driver(ctl, idz=idz, wall_file=wf, wall_file_window=True)
        a.setup(); b.setup()
        assert np.array_equal(a.walls, b.walls)
        assert a.i64("pore_sum_local") == b.i64("pore_sum_local")
        assert a.solid_nodes().tobytes() == b.solid_nodes().tobytes()
        assert a.fluid_nodes().tobytes() == b.fluid_nodes().tobytes()
        assert np.array_equal(a.field("phi"), b.field("phi"))
        a.close(); b.close()


@pytest.mark.parametrize("option", [1, 2, 3, 4, 5])
def test_initial_fluid_distributions_match_oracle(tmp_path, option):
    """initialization_new_multi, every deterministic initial_fluid_distribution_option (MP/Init_multiphase.F90:243-330;
    option 6 draws unseeded random numbers in the reference and is injected from the host instead)."""
    rng = np.random.default_rng(40 + option)
    wg = (rng.random((16, 14, 24)) < 0.2).astype(np.int8)
    d = _driver(tmp_path, walls=wg, lattice_dimensions="16,14,24", initial_fluid_distribution_option=option,
                initial_interface_position=7.0, excluded_layers="2,2")
    o = make_oracle(nxG=16, nyG=14, nzG=24, walls_global=wg, initial_fluid_distribution_option=option, interface_z0=7.0,
                    n_exclude_inlet=2, n_exclude_outlet=2)
    assert np.array_equal(d.field("phi"), o.field("phi"))
    for q in range(19):
        assert np.array_equal(d.field("f", q), o.f(q)) and np.array_equal(d.field("g", q), o.g(q))
    d.close()


@pytest.mark.parametrize("inlet,outlet", [(1, 1), (2, 2), (1, 2)])
def test_boundary_condition_setups_match_oracle(tmp_path, inlet, outlet):
    """initialization_basic_multi for the velocity / pressure inlet-outlet combinations: derived scalars and the inlet
    profile (MP/Init_multiphase.F90:116-236, MP/Misc.F90:625-665)."""
    d = _driver(tmp_path, modify_geometry_cmd=1, inlet_BC=inlet, outlet_BC=outlet)
    o = make_oracle(modify_geometry_cmd=1, inlet_BC=inlet, outlet_BC=outlet)
    for n in ("la_nui1", "la_nui2", "phi_inlet", "force_Z", "uin_avg", "flowrate", "rho_in", "rho_out", "relaxation"):
        assert d.f64(n) == o.get_double(n), n
    assert np.array_equal(d.field("w_in"), o.field("w_in"))
    assert np.array_equal(d.field("phi"), o.field("phi"))
    d.close()


def test_y_periodic_setup_matches_oracle(tmp_path):
    """reference test-suite case 1 shape (1.drop_attached_wall: y and z periodic, domain_wall_status_y = 0): walls incl. the
    wrapped y ghost rows, boundary-node lists and the initial state"""
    rng = np.random.default_rng(12)
    wg = (rng.random((18, 16, 20)) < 0.15).astype(np.int8)
    d = _driver(tmp_path, walls=wg, lattice_dimensions="18,16,20", periodic_indicator="0,1,1", domain_wall_status_y="0,0",
                initial_fluid_distribution_option=3, initial_interface_position=6.0, theta=45, inlet_BC=0, outlet_BC=0,
                excluded_layers="0,0")
    o = make_oracle(nxG=18, nyG=16, nzG=20, jper=1, kper=1, wsy0=0, wsy1=0, walls_global=wg, initial_fluid_distribution_option=3,
                    interface_z0=6.0, theta_deg=45.0, inlet_BC=0, outlet_BC=0, n_exclude_inlet=0, n_exclude_outlet=0)
    assert np.array_equal(d.walls, o.walls)
    _same_lists(d, o)
    assert np.array_equal(d.field("phi"), o.field("phi"))
    d.close()
