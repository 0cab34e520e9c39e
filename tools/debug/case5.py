"""debug: Bentheimer crop case 5 (z-periodic, random phi), sparse strict, step by step"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import compare_state, ctx_from_oracle
from oracle.oracle import Oracle, default_params
from test_reference_inputs_gpu import _bentheimer_crop

def run(n, layout, tag):
    w = _bentheimer_crop(n)
    p = default_params(nxG=n, nyG=n, nzG=n, kper=1, inlet_BC=0, outlet_BC=0, la_nu2=0.04, force_z0=200e-6, n_exclude_inlet=0,
                       n_exclude_outlet=0, initial_fluid_distribution_option=5)
    o = Oracle(p)
    o.set_walls(w); o.geometry_preprocess(); o.init_basic(); o.init_phi()
    rng = np.random.default_rng(20261018)
    o.field("phi")[...] = np.where(rng.random(o.field("phi").shape) > 0.4, -1.0, 1.0)
    o.init_pdf()
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=layout)
    o.color_gradient(); ctx.color_gradient()
    for t in range(0, 9):
        if t:
            o.step(t); ctx.run(t, 1); ctx.sync()
        try:
            compare_state(ctx, o, 0.0, sparse=layout == 2)
            print(tag, "n", n, "step", t, "ok", flush=True)
        except AssertionError as e:
            print(tag, "n", n, "step", t, "FAIL", str(e)[:600], flush=True)
            break
    ctx.close()

for n in (48, 96):
    run(n, 2, "sparse env=%s" % {k: v for k, v in os.environ.items() if k.startswith("MFLBM_")})
