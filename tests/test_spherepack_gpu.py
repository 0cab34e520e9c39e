"""GPU parity on SPHERE-PACK media (the geometry class of BASELINE configs[1..4]) against the CPU oracle.

The random-noise media of test_parity_gpu.py have runs of 3-4 fluid nodes, which sends nearly every warp of the sparse
layout down the IRREGULAR adjacency path (verbatim index rows); a sphere pack has long runs, so these cases exercise the
compressed REGULAR path (mflbm_internal.cuh adj_index_fast), the compact link slots and the quiet-tile logic the way the
benchmarks do -- at a size the oracle steps in seconds.  Strict (-fmad=false) build: bit-exact.
"""
from importlib import import_module

import numpy as np
import pytest

from helpers import compare_state, ctx_from_oracle, make_oracle

pytestmark = pytest.mark.gpu
geo = import_module("mflbm_b200.geometry")
LAYOUTS = [pytest.param(1, id="dense"), pytest.param(2, id="sparse")]


def _pack(nx, ny, nz, periodic, seed):
    return geo.sphere_pack(nx, ny, nz, periodic=periodic, porosity=0.4, rmin=5.0, rmax=10.0, seed=seed, buffer=6)


def _run_both(o, ctx, nsteps, ntime0=1):
    for n in range(ntime0, ntime0 + nsteps):
        o.step(n)
    ctx.run(ntime0, nsteps)
    ctx.sync()


@pytest.mark.parametrize("layout", LAYOUTS)
def test_singlephase_spherepack_periodic(layout):
    wg = _pack(72, 64, 80, True, 11)
    o = make_oracle(multiphase=0, nxG=72, nyG=64, nzG=80, la_nu1=0.1, kper=1, force_z0=1e-5, walls_global=wg,
                    n_exclude_inlet=0, n_exclude_outlet=0)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=layout)
    t = 1
    for nsteps in (1, 1, 10):
        _run_both(o, ctx, nsteps, t)
        t += nsteps
        compare_state(ctx, o, 0.0, sparse=layout == 2)
    # the per-slice mass profile of SP/Monitor.F90:36-38 (valid after an even step)
    m = ctx.monitor()
    o.monitor()
    ref = o.field("pre")
    assert np.max(np.abs(m["pre"] - ref)) <= 1e-12 * np.max(np.abs(ref))
    ctx.close()


@pytest.mark.parametrize("layout", LAYOUTS)
def test_multiphase_spherepack_drainage(layout):
    wg = _pack(72, 64, 80, False, 12)
    o = make_oracle(nxG=72, nyG=64, nzG=80, la_nu2=0.04, interface_z0=8.0, walls_global=wg, n_exclude_inlet=6,
                    n_exclude_outlet=6)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=layout)
    o.color_gradient(); ctx.color_gradient()
    compare_state(ctx, o, 0.0, sparse=layout == 2)
    t = 1
    for nsteps in (1, 1, 12):
        _run_both(o, ctx, nsteps, t)
        t += nsteps
        compare_state(ctx, o, 0.0, sparse=layout == 2)
    v1, v2 = ctx.cal_saturation()
    s = o.cal_saturation()
    assert abs(v1 - s["vol1_sum"]) <= 1e-10 * abs(s["vol1_sum"]) and abs(v2 - s["vol2_sum"]) <= 1e-10 * abs(s["vol2_sum"])
    ctx.close()
