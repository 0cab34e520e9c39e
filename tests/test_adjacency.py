"""CPU test of the integer machinery of the sparse population layout: node numbering, compact link slots and the
per-warp compressed adjacency of the odd step (mf-lbm_b200/csrc/mflbm_internal.cuh "Adjacency").  The library's host-only
self-test decodes every (node, direction) with the very function the CUDA kernel uses and compares it with the direct
neighbour lookup; integer work must be bit-exact."""
import ctypes as C

import numpy as np
import pytest

import mflbm_b200 as M
from helpers import make_oracle


def _selftest(walls):
    M.build()
    lib = M.load()
    fn = lib.mflbmx_adjacency_selftest
    fn.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_longlong * 8)]
    w = np.asfortranarray(walls, dtype=np.int8)
    nx, ny, nz = (s - 4 for s in w.shape)
    counts = (C.c_longlong * 8)()
    rc = fn(nx, ny, nz, w.ctypes.data, C.byref(counts))
    assert rc == 0, lib.mflbm_last_error(None).decode()
    return list(counts)[:4]


def _with_ghosts(core, wall_xy=True):
    """(nx,ny,nz) 0/1 array -> walls(-1:n+2)^3 with solid x/y ghost layers like a walled channel and open z ghosts"""
    nx, ny, nz = core.shape
    w = np.ones((nx + 4, ny + 4, nz + 4), np.int8) if wall_xy else np.zeros((nx + 4, ny + 4, nz + 4), np.int8)
    w[2:-2, 2:-2, :] = 0
    w[2:-2, 2:-2, 2:-2] = core
    return w


def test_c1_tube_sphere_adjacency():
    o = make_oracle(modify_geometry_cmd=1)
    nA, nAct, links, ovf = _selftest(o.walls)
    assert nA == 73936                      # pore_sum of C1
    assert nAct >= nA and links > 0


@pytest.mark.parametrize("shape,p,seed", [((33, 17, 40), 0.3, 1), ((64, 8, 9), 0.6, 2), ((7, 5, 4), 0.5, 3), ((40, 40, 12), 0.05, 4),
                                          ((31, 31, 31), 0.9, 5)])
def test_random_media_adjacency(shape, p, seed):
    rng = np.random.default_rng(seed)
    core = (rng.random(shape) < p).astype(np.int8)
    for wall_xy in (True, False):
        nA, nAct, links, ovf = _selftest(_with_ghosts(core, wall_xy))
        interior = core if wall_xy else core
        assert nA == int((interior == 0).sum())


def test_all_fluid_and_all_solid():
    nA, nAct, links, ovf = _selftest(_with_ghosts(np.zeros((20, 12, 10), np.int8)))
    assert nA == 20 * 12 * 10
    nA, nAct, links, ovf = _selftest(_with_ghosts(np.ones((6, 6, 6), np.int8)))
    assert nA == 0 and links == 0


@pytest.mark.parametrize("shape", [(1536, 1536, 2), (1000, 3, 5), (17, 1023, 3), (240, 240, 20), (512, 512, 4)])
def test_cell_decomposition_on_large_cross_sections(shape):
    """the self-test also checks Grid::coords3 (multiply-high division by the row / plane strides) on every cell: run it
    on the cross-sections of the BASELINE configs (strides 1552 x 1544, 256 x 248, 528 x 520) and on awkward ones"""
    core = np.zeros(shape, np.int8)
    core[::7, ::5, :] = 1
    nA, nAct, links, ovf = _selftest(_with_ghosts(core))
    assert nA == int((core == 0).sum())


BENTHEIMER = "/root/reference/MF-LBM-extFiles/geometry_files/sample_rock_geometry_wallarray/bentheimer_in10_240_240_240_out10.dat"


@pytest.mark.skipif(not __import__("os").path.exists(BENTHEIMER), reason="reference fixture not mounted (GPU box)")
def test_reference_bentheimer_rock_counts_and_adjacency():
    """The reference's own C2 geometry (240x240x260 Bentheimer sandstone with 10 buffer layers each end; read here only,
    never copied): 3 670 813 pore nodes (SURVEY 8(c) known answer) through the wall-array reader, and the node numbering /
    link slots / compressed adjacency of the sparse layout self-check on a real rock."""
    from oracle.oracle import read_wall_array
    w = read_wall_array(BENTHEIMER)
    assert w.shape == (240, 240, 260) and set(np.unique(w)) == {0, 1}
    assert int((w == 0).sum()) == 3670813
    nA, nAct, links, ovf = _selftest(_with_ghosts(w))
    assert nA == 3670813 and links > 0 and nAct >= nA
