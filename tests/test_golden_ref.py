"""Golden vectors of the reference's own source (tests/golden/ref_c1_vectors.json, made by tests/golden/make_ref_vectors.py
from oracle/_ref = the reference's Fortran translated mechanically and compiled here) on BASELINE configs[0].

CPU: the hand-written oracle reproduces every digest (bit-exact) and the 1000-step quantities.
GPU (-m gpu, through the C ABI): the strict (-fmad=false) build reproduces the digests on both population layouts; the
default FMA build matches the integrated quantities after 1000 steps to 1e-8 and the breakthrough count exactly
(MP/Monitor.F90:158-238, :472-507)."""
import importlib.util
import json
import os

import numpy as np
import pytest

from helpers import ctx_from_oracle, make_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_c1_vectors.json")))
_spec = importlib.util.spec_from_file_location("make_ref_vectors", os.path.join(HERE, "golden", "make_ref_vectors.py"))
V = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(V)


def _oracle():
    o = make_oracle(modify_geometry_cmd=1)
    assert int((o.walls[2:-2, 2:-2, 2:-2] == 0).sum()) == GOLD["fluid_nodes"] == 73936
    return o


def test_oracle_reproduces_reference_vectors():
    o = _oracle()
    masks = V.fluid_masks(o.walls)
    o.color_gradient()
    t = 0
    for upto in (1, 2, 10, 100, 1000):
        while t < upto:
            t += 1
            o.step(t)
        if upto < 1000:
            assert V.state_digest(o.f, o.g, o.field, masks) == GOLD["steps"][str(upto)], "state after %d steps" % upto
    fin = GOLD["after_1000_steps"]
    s = o.cal_saturation()
    assert s["vol1_sum"] == pytest.approx(fin["vol1_sum"], rel=1e-12) and s["vol2_sum"] == pytest.approx(fin["vol2_sum"], rel=1e-12)
    assert o.monitor_breakthrough()["outlet_phase1_sum"] == fin["outlet_phase1_sum"]
    o.compute_macro_vars()
    got = V.integrated(o.field("u"), o.field("v"), o.field("w"), o.field("rho"), o.field("phi"), masks)
    for k, v in got.items():
        assert v == pytest.approx(fin[k], rel=1e-12), k


@pytest.mark.gpu
@pytest.mark.parametrize("layout", [1, 2], ids=["dense", "sparse"])
def test_cuda_strict_reproduces_reference_digests(layout):
    o = _oracle()
    masks = V.fluid_masks(o.walls)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=layout)
    ctx.color_gradient()
    t = 1
    for upto in (1, 2, 10, 100):
        ctx.run(t, upto - t + 1)
        t = upto + 1
        got = ctx.download("f", "g", "phi", "cn_x", "cn_y", "cn_z", "c_norm", "curv")
        d = V.state_digest(lambda q: got["f"][q], lambda q: got["g"][q], lambda n: got[n], masks)
        bad = [k for k in d if d[k] != GOLD["steps"][str(upto)][k]]
        assert not bad, "after %d steps: %s" % (upto, bad)
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("layout", [1, 2], ids=["dense", "sparse"])
def test_cuda_fma_1000_steps_integrated_quantities(layout):
    """default (FMA-contracting) build: 1e-8 on the integrated quantities after 1000 steps, breakthrough count identical"""
    o = _oracle()
    masks = V.fluid_masks(o.walls)
    ctx = ctx_from_oracle(o, strict=False, kernel_variant=layout)
    ctx.color_gradient()
    ctx.run(1, 1000)
    fin = GOLD["after_1000_steps"]
    v1, v2 = ctx.cal_saturation()
    assert v1 == pytest.approx(fin["vol1_sum"], rel=1e-8) and v2 == pytest.approx(fin["vol2_sum"], rel=1e-8)
    assert v1 / (v1 + v2) == pytest.approx(fin["saturation_full_domain"], rel=1e-8)
    assert ctx.monitor_breakthrough() == fin["outlet_phase1_sum"]
    ctx.compute_macro_vars()
    got = ctx.download("u", "v", "w", "rho", "phi")
    q = V.integrated(got["u"], got["v"], got["w"], got["rho"], got["phi"], masks)
    for k, v in q.items():
        assert v == pytest.approx(fin[k], rel=1e-8), k
    ctx.close()
