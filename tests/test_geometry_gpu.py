"""GPU parity of the device geometry preprocessing against the CPU oracle's restatement of
geometry_preprocessing_new (MP/Geometry_preprocessing.F90:9-512).

Integer work (classification, list order, coordinates, neighbour lists, counts) must be bit-exact; la_weight and the ISO8
wall normals are bit-exact too, in both library builds: the device code spells every FP64 operation with round-to-nearest
intrinsics in the reference's term order, so there is nothing for the compiler to contract.  Covers the whole-lattice
path, z windows of a slab decomposition (the lists must equal the oracle's local lists of that slab), periodic
extension, and degenerate media (no solid at all, everything solid)."""
from importlib import import_module

import numpy as np
import pytest

import mflbm_b200 as M
from helpers import make_oracle

pytestmark = pytest.mark.gpu
geo = import_module("mflbm_b200.geometry")


def _same_lists(solid, fluid, o):
    so, fo = o.solid_nodes(), o.fluid_nodes()
    assert len(solid) == len(so) and len(fluid) == len(fo), (len(solid), len(so), len(fluid), len(fo))
    for name in ("ix", "iy", "iz", "i_fluid_num", "neighbor_list"):
        assert np.array_equal(solid[name], so[name]), name
    assert np.array_equal(solid["la_weight"], so["la_weight"])
    for name in ("ix", "iy", "iz"):
        assert np.array_equal(fluid[name], fo[name]), name
    for name in ("nwx", "nwy", "nwz", "theta"):
        assert np.array_equal(fluid[name], fo[name]), name   # bit-exact FP64


@pytest.mark.parametrize("strict", [False, True], ids=["fma", "strict"])
def test_c1_tube_sphere_lists(strict):
    o = make_oracle(modify_geometry_cmd=1)
    solid, fluid, gs, gf = M.geometry_preprocess(o.walls_global, theta=o.get_double("theta"), strict=strict)
    _same_lists(solid, fluid, o)
    assert gs == o.get_int("num_solid_global") and gf == o.get_int("num_fluid_global")


@pytest.mark.parametrize("kper", [0, 1])
def test_spherepack_whole_lattice(kper):
    wg = geo.sphere_pack(56, 48, 64, periodic=bool(kper), porosity=0.4, rmin=4.0, rmax=9.0, seed=21, buffer=5)
    o = make_oracle(nxG=56, nyG=48, nzG=64, kper=kper, force_z0=1e-5 if kper else 0.0, walls_global=wg, n_exclude_inlet=0,
                    n_exclude_outlet=0)
    solid, fluid, gs, gf = M.geometry_preprocess(o.walls_global, periodic=(0, 0, kper), theta=o.get_double("theta"))
    _same_lists(solid, fluid, o)
    assert gs == o.get_int("num_solid_global") and gf == o.get_int("num_fluid_global")


@pytest.mark.parametrize("npz", [2, 4])
def test_slab_windows_equal_the_global_run(npz):
    """every rank preprocesses only planes [slab - 12, slab + 12] and must get exactly its local lists"""
    nzG = 96
    wg = geo.sphere_pack(40, 36, nzG, periodic=False, porosity=0.45, rmin=4.0, rmax=8.0, seed=22, buffer=4)
    for idz in range(npz):
        o = make_oracle(nxG=40, nyG=36, nzG=nzG, npz=npz, idz=idz, walls_global=wg, n_exclude_inlet=0, n_exclude_outlet=0)
        full = o.walls_global
        nz = nzG // npz
        k0, k1 = max(1, idz * nz + 1 - 12), min(nzG, idz * nz + nz + 12)
        solid, fluid, _, _ = M.geometry_preprocess(full[:, :, k0 - 1:k1], nzGlobal=nzG, wk0=k0, idz=idz, npz=npz,
                                                   theta=o.get_double("theta"))
        _same_lists(solid, fluid, o)


def test_degenerate_media_and_errors():
    empty = np.zeros((12, 10, 16), np.int8)
    solid, fluid, gs, gf = M.geometry_preprocess(empty)
    assert len(solid) == 0 and len(fluid) == 0 and gs == 0 and gf == 0
    full = np.ones((12, 10, 16), np.int8)
    solid, fluid, gs, gf = M.geometry_preprocess(full)
    assert len(solid) == 0 and len(fluid) == 0
    with pytest.raises(M.MflbmError, match="window too small"):
        M.geometry_preprocess(np.zeros((12, 10, 12), np.int8), nzGlobal=64, wk0=20, idz=1, npz=4)


def test_host_driver_device_geometry_equals_host_path(tmp_path):
    """the driver mirror with geometry_preprocessing_new on the GPU hands the hot path the same node lists"""
    wg = geo.sphere_pack(48, 40, 72, periodic=False, porosity=0.4, rmin=4.0, rmax=9.0, seed=23, buffer=5)
    ctl = M.write_control_file(str(tmp_path / "ctl.txt"), multiphase=True, lattice_dimensions="48,40,72", MPI_process_num="1,1,1",
                               external_geometry_read_cmd=1, excluded_layers="5,5")
    lists = []
    for dev in (None, 0):
        drv = M.Driver(ctl, idz=0, walls_window=(wg, 1), device_geometry=dev)
        drv.setup()
        lists.append((drv.solid_nodes().copy(), drv.fluid_nodes().copy(), drv.i64("num_solid_boundary_global"),
                      drv.i64("num_fluid_boundary_global")))
        drv.close()
    (s0, f0, gs0, gf0), (s1, f1, gs1, gf1) = lists
    assert len(s0) > 0 and len(f0) > 0
    assert s0.tobytes() == s1.tobytes()
    assert f0.tobytes() == f1.tobytes()
    assert (gs0, gf0) == (gs1, gf1)
