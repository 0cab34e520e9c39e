"""The fused colour-gradient kernel (csrc/march.cuh: K3..K7 on chip, sparse multiphase layout) on the GPU.

Every sparse-layout multiphase test of the suite runs through it by default (the populations after a step depend on what it
wrote); here it is additionally compared, entry by entry of the packed gradient, with the five reference-order kernels
(mflbm_chain_selfcheck: strict build, bit for bit) in both of its launch shapes -- everything (after an upload, and while most
tiles hold an interface) and the work items around the active tiles (a drainage front) -- and the first case also runs with
the list kernels (the default: the fused kernel is measured slower and therefore opt-in, MFLBM_MARCH=1).  CPU twin on the same source: tests/test_march_emu.py."""
from importlib import import_module

import numpy as np
import pytest

from helpers import compare_state, ctx_from_oracle, make_oracle
from oracle.oracle import Oracle, default_params

pytestmark = pytest.mark.gpu
geo = import_module("mflbm_b200.geometry")


def _run_both(o, ctx, nsteps, t):
    for n in range(t, t + nsteps):
        o.step(n)
    ctx.run(t, nsteps)
    ctx.sync()
    return t + nsteps


@pytest.fixture(autouse=True)
def _march_on(monkeypatch):
    """the fused kernel is opt-in (measured slower than the list kernels): every test here selects it itself"""
    monkeypatch.setenv("MFLBM_MARCH", "1")


def _random_phi(wg, wrap_y=False, **kw):
    nx, ny, nz = wg.shape
    p = default_params(nxG=nx, nyG=ny, nzG=nz, n_exclude_inlet=0, n_exclude_outlet=0, initial_fluid_distribution_option=5, **kw)
    o = Oracle(p)
    o.set_walls(wg); o.geometry_preprocess(); o.init_basic(); o.init_phi()
    rng = np.random.default_rng(nx * 1000 + nz)
    phi = np.where(rng.random(o.field("phi").shape) > 0.4, -1.0, 1.0)
    if wrap_y:  # the populations of the y ghost rows are images on the device (adjacency): phi must be periodic in y to start with
        phi[:, :4, :] = phi[:, ny:ny + 4, :]
        phi[:, ny + 4:, :] = phi[:, 4:8, :]
    o.field("phi")[...] = phi
    o.init_pdf()
    return o


@pytest.mark.parametrize("fused", [1, 2, 0], ids=["march", "hybrid", "lists"])
def test_drainage_front_both_shapes(fused, monkeypatch):
    """C1-like duct long enough for quiet tiles: first every tile is evaluated, later only the items around the front
    (hybrid = MFLBM_MARCH=2: the fused kernel for those items, the flat sweeps of the list kernels for "every tile")"""
    monkeypatch.setenv("MFLBM_MARCH", str(fused))  # (overrides the module fixture)
    wg = geo.sphere_pack(72, 40, 96, periodic=False, porosity=0.4, rmin=4.0, rmax=8.0, seed=21, buffer=6)
    o = make_oracle(nxG=72, nyG=40, nzG=96, la_nu2=0.04, interface_z0=8.0, ca_0=2e-3, walls_global=wg, n_exclude_inlet=6, n_exclude_outlet=6)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=2)
    assert ctx.chain_info() == (1 if fused else 0, 0)
    o.color_gradient(); ctx.color_gradient()
    assert ctx.chain_selfcheck() == 0
    compare_state(ctx, o, 0.0, sparse=True)
    t = 1
    for nsteps in (1, 1, 2, 30, 31):
        t = _run_both(o, ctx, nsteps, t)
        assert ctx.chain_selfcheck() == 0, "after step %d" % (t - 1)
    nt, nq = ctx.tile_stats()
    assert 0 < nq < nt  # the tile-driven shape ran
    compare_state(ctx, o, 0.0, sparse=True)
    t = _run_both(o, ctx, 7, t)  # the dense arrays the download above materialised do not disturb the next steps
    compare_state(ctx, o, 0.0, sparse=True)
    ctx.close()


@pytest.mark.parametrize("mode", ["1", "2"], ids=["march", "hybrid"])
@pytest.mark.parametrize("dims", [(72, 64, 40), (33, 17, 21), (96, 80, 72)])
def test_random_phi_periodic(dims, mode, monkeypatch):
    """interface everywhere (reference benchmark case 6): the flat shape, tile-unaligned lattices, theta = 150 degrees"""
    monkeypatch.setenv("MFLBM_MARCH", mode)
    nx, ny, nz = dims
    wg = geo.sphere_pack(nx, ny, nz, periodic=True, porosity=0.45, rmin=3.0, rmax=7.0, seed=5, buffer=0)
    o = _random_phi(wg, kper=1, inlet_BC=0, outlet_BC=0, force_z0=2e-4, la_nu2=0.04, theta_deg=150.0)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=2)
    assert ctx.chain_info() == (1, 0)
    o.color_gradient(); ctx.color_gradient()
    assert ctx.chain_selfcheck() == 0
    t = 1
    for nsteps in (1, 1, 6):
        t = _run_both(o, ctx, nsteps, t)
        assert ctx.chain_selfcheck() == 0
        compare_state(ctx, o, 0.0, sparse=True)
    ctx.close()


def test_y_periodic_and_fma_build():
    wg = geo.sphere_pack(24, 32, 24, periodic=True, porosity=0.5, rmin=3.0, rmax=6.0, seed=3, buffer=0)
    for strict in (True, False):
        o = _random_phi(wg, wrap_y=True, jper=1, kper=1, wsy0=0, wsy1=0, inlet_BC=0, outlet_BC=0, force_z0=1e-4)
        ctx = ctx_from_oracle(o, strict=strict, kernel_variant=2)
        assert ctx.chain_info() == (1, 0)
        o.color_gradient(); ctx.color_gradient()
        t = _run_both(o, ctx, 2, 1)
        if strict:
            assert ctx.chain_selfcheck() == 0
        compare_state(ctx, o, 0.0 if strict else 1e-9, sparse=True)
        ctx.close()


def test_foreign_node_lists_fall_back_to_the_list_kernels():
    """la_weight that is not the sum over the listed neighbours: the fused kernel (which recomputes it) steps aside"""
    o = make_oracle(modify_geometry_cmd=1)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=2)
    solid = o.solid_nodes().copy()
    solid["la_weight"][::7] *= 1.25
    ctx.upload(solid_boundary_nodes=solid)
    assert ctx.chain_info() == (0, 2)
    ctx.color_gradient()
    ctx.run(1, 4)
    ctx.sync()
    assert ctx.chain_selfcheck() == 0  # list kernels against themselves: the path still works
    ctx.close()


@pytest.mark.parametrize("brick", ["", "32,8,4", "16,4,2", "0,4,4", "k7:96,4,2", "k7:24,2,3"],
                         ids=["raster", "b32x8x4", "b16x4x2", "rows4x4", "k7_96x4x2", "k7_24x2x3"])
def test_flat_sweeps_of_the_list_kernels(brick, monkeypatch):
    """the default chain (list kernels) in its flat shape -- every tile holds an interface: K4 + K5 fused in one sweep, the lists
    walked in brick order (MFLBM_BRICK) -- against the oracle and against the separate reference-order kernels (self-check)"""
    monkeypatch.setenv("MFLBM_MARCH", "0")
    if brick.startswith("k7:"):  # K7 + packing alone in brick order
        monkeypatch.setenv("MFLBM_BRICK7", brick[3:])
    elif brick:
        monkeypatch.setenv("MFLBM_BRICK", brick)
    wg = geo.sphere_pack(72, 64, 40, periodic=True, porosity=0.45, rmin=3.0, rmax=7.0, seed=5, buffer=0)
    o = _random_phi(wg, kper=1, inlet_BC=0, outlet_BC=0, force_z0=2e-4, la_nu2=0.04, theta_deg=150.0)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=2)
    assert ctx.chain_info() == (0, 0)
    o.color_gradient(); ctx.color_gradient()
    assert ctx.chain_selfcheck() == 0
    t = 1
    for nsteps in (1, 1, 6):
        t = _run_both(o, ctx, nsteps, t)
        assert ctx.chain_selfcheck() == 0
        compare_state(ctx, o, 0.0, sparse=True)
    nt, nq = ctx.tile_stats()
    assert nq * 4 < nt  # most tiles active: the flat sweeps ran
    ctx.close()
