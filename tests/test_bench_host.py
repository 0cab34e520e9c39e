"""CPU checks of the host-side plumbing around the hot path that bench.py and the multi-GPU runs rely on: the synthetic
media are deterministic functions of (shape, seed) so that every rank can rasterise just its own z window and still see
the same medium; the workload table names every BASELINE.json config; the slab windows cover what the geometry
preprocessing needs."""
from importlib import import_module

import numpy as np
import pytest

import bench
import mflbm_b200 as M

geo = import_module("mflbm_b200.geometry")


def test_windows_of_the_sphere_pack_equal_slices_of_the_whole():
    kw = dict(porosity=0.36, rmin=4.0, rmax=9.0, seed=5, buffer=6)
    full = geo.sphere_pack(48, 40, 96, periodic=False, **kw)
    for k0, k1 in ((1, 96), (1, 30), (25, 72), (60, 96)):
        w = geo.sphere_pack_window(48, 40, 96, k0, k1, periodic=False, **kw)
        assert np.array_equal(w, full[:, :, k0 - 1:k1])
    # periodic lattice: a window may leave 1..nz and wraps
    fullp = geo.sphere_pack(48, 40, 96, periodic=True, **kw)
    w = geo.sphere_pack_window(48, 40, 96, -3, 10, periodic=True, **kw)
    assert np.array_equal(w[:, :, 4:], fullp[:, :, :10]) and np.array_equal(w[:, :, :4], fullp[:, :, -4:])
    assert np.all(full[:, :, :6] == 0) and np.all(full[:, :, -6:] == 0)  # buffer layers stay open


def test_sphere_pack_hits_the_target_porosity():
    w = geo.sphere_pack(96, 96, 116, periodic=False, porosity=0.36, rmin=8.0, rmax=20.0, seed=1, buffer=10)
    core = w[:, :, 10:-10]
    assert abs((core == 0).mean() - 0.36) < 0.04  # Boolean model on a small box; the 512^3 realisation lands within 0.005


def test_c2rock_workload_reaches_the_driver_with_the_reference_pore_count(tmp_path):
    """bench.py --workload c2rock: the host side (control file, wall window, set_walls) on the reference's rock"""
    spec = bench.workload_spec("c2rock", 1)
    ctl = M.write_control_file(str(tmp_path / "ctl.txt"), multiphase=False, lattice_dimensions="240,240,260", MPI_process_num="1,1,1",
                               external_geometry_read_cmd=1, **spec["control"])
    w = geo.load_packed_walls(spec["walls_file"], (240, 240, 260))
    drv = M.Driver(ctl, idz=0, walls_window=(w, 1), lazy_pdfs=True)
    drv.setup()
    assert drv.i64("pore_sum_local") == 3670813
    drv.close()


@pytest.mark.parametrize("name,mp,cross", [("c1", True, (40, 40)), ("c2", False, (240, 240)), ("c3", True, (512, 512)),
                                            ("c5", True, (1536, 1536))])
def test_workload_table_names_every_baseline_config(name, mp, cross):
    for n in (1, 2, 8):
        s = bench.workload_spec(name, n)
        assert s["multiphase"] == mp and (s["nx"], s["ny"]) == cross
        assert s["nz"] % n == 0 and s["label"]
    assert bench.BYTES_PER_UPDATE[True] == 624.0 and bench.BYTES_PER_UPDATE[False] == 304.0


def test_strong_scaling_workload_keeps_the_global_lattice():
    for n in (1, 2, 4, 8):
        s = bench.workload_spec("c4", n)  # BASELINE configs[3]
        assert (s["nx"], s["ny"], s["nz"]) == (512, 512, 1024) and s["scaling"] == "strong"


def test_slab_windows_cover_the_preprocessing_stencils():
    """Driver.window_range: every slab's window reaches the lattice end or extends >= 10 planes beyond the slab, which is
    what mflbm_geometry_preprocess demands (classification radius 1 + four smoothing passes + ISO8 radius 2)."""
    for nzG, npz in ((512, 2), (1024, 4), (1536, 8), (96, 4)):
        nz = nzG // npz
        for idz in range(npz):
            k0, k1 = M.Driver.window_range(idz, npz, nzG, False)
            lo, hi = idz * nz + 1, idz * nz + nz
            assert k0 == 1 or k0 <= lo - 10
            assert k1 == nzG or k1 >= hi + 10
            assert 1 <= k0 <= lo and hi <= k1 <= nzG


def test_bentheimer_fixture_is_the_reference_rock():
    """tests/golden/bentheimer_in10_240_out10.bits.xz (made by tests/golden/make_fixtures.py): 240x240x260, 3 670 813 pore
    nodes; identical to the reference's file whenever that is mounted."""
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    w = geo.load_packed_walls(os.path.join(here, "golden", "bentheimer_in10_240_out10.bits.xz"), (240, 240, 260))
    assert int((w == 0).sum()) == 3670813 and set(np.unique(w)) == {0, 1}
    # ten open buffer layers at each end inside the x / y side walls of the sample
    assert np.all(w[1:-1, 1:-1, :10] == 0) and np.all(w[1:-1, 1:-1, -10:] == 0) and np.all(w[0] == 1) and np.all(w[:, 0] == 1)
    ref = "/root/reference/MF-LBM-extFiles/geometry_files/sample_rock_geometry_wallarray/bentheimer_in10_240_240_240_out10.dat"
    if os.path.exists(ref):
        from oracle.oracle import read_wall_array
        assert np.array_equal(w, read_wall_array(ref))
