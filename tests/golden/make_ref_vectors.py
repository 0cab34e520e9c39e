#!/usr/bin/env python
"""Golden vectors produced by THE REFERENCE'S OWN SOURCE (run in the build container, where /root/reference exists):

    python tests/golden/make_ref_vectors.py        ->  tests/golden/ref_c1_vectors.json

oracle/_ref (the reference's subroutines translated mechanically by oracle/f2c_lite.py and compiled with gcc
-ffp-contract=off) runs BASELINE configs[0] -- the 40x40x60 tube + sphere drainage with the template parameters --
for 1000 time steps.  The geometry lists come from the oracle's geometry preprocessing (that routine is outside the
translatable subset); the phase field and populations are initialised by the reference's initialization_new_multi.
Recorded: SHA-256 of the fluid-node values of every population array, phi, the interface normal, |grad phi| and the
curvature after 1, 2, 10 and 100 steps (bit-exact contract for un-contracted FP64 arithmetic), and integrated
quantities after 1000 steps (1e-8 contract for FMA builds).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def fluid_masks(walls):
    """walls (-1:n+2)^3 -> masks of the fluid nodes of 1..n for arrays with ghost widths 1, 2 and 4"""
    inner = walls[2:-2, 2:-2, 2:-2] == 0
    out = {}
    for o in (1, 2, 4):
        m = np.zeros(tuple(s + 2 * o for s in inner.shape), bool)
        m[o:-o, o:-o, o:-o] = inner
        out[o] = m
    return out


def state_digest(get_f, get_g, get_field, masks):
    h = {}
    for q in range(19):
        h["f%d" % q] = hashlib.sha256(np.ascontiguousarray(get_f(q)[masks[1]] + 0.0).tobytes()).hexdigest()
        h["g%d" % q] = hashlib.sha256(np.ascontiguousarray(get_g(q)[masks[1]] + 0.0).tobytes()).hexdigest()
    for n, o in (("phi", 4), ("cn_x", 2), ("cn_y", 2), ("cn_z", 2), ("c_norm", 2), ("curv", 1)):
        h[n] = hashlib.sha256(np.ascontiguousarray(get_field(n)[masks[o]] + 0.0).tobytes()).hexdigest()  # + 0.0: -0 -> +0
    return h


def integrated(u, v, w, rho, phi, masks):
    """what the reference's monitor integrates (MP/Monitor.F90:33-85), over the fluid nodes of 1..n"""
    m = masks[1]
    ph = phi[3:-3, 3:-3, 3:-3][m]
    ww, rr = w[m], rho[m]
    plane = lambda k: rho[:, :, k][m[:, :, k]].mean()  # mean density of the pore space of plane k
    return {"flow1": float(np.sum(ww * 0.5 * (1.0 + ph))), "flow2": float(np.sum(ww * 0.5 * (1.0 - ph))),
            "umax_sq": float(np.max(u[m] ** 2 + v[m] ** 2 + ww ** 2)), "rho_sum": float(np.sum(rr)),
            "pressure_drop": float((plane(11) - plane(50)) / 3.0)}


def main():
    from helpers import make_oracle
    from ref_helpers import ref_from_oracle
    o = make_oracle(modify_geometry_cmd=1)
    r = ref_from_oracle(o)
    r.call("initialization_new_multi")
    r.array("w_in")[...] = o.field("w_in")
    masks = fluid_masks(o.walls)
    out = {"case": "C1 tube+sphere 40x40x60, template parameters (nu1 0.004, nu2 0.4, gamma 0.03, theta 30, beta 0.95, Ca 1e-4)",
           "fluid_nodes": int(masks[1].sum()), "steps": {}}
    r.call("color_gradient")
    t = 0
    for upto in (1, 2, 10, 100, 1000):
        while t < upto:
            t += 1
            r.set(ntime=t)
            r.call("main_iteration_kernel")
        if upto < 1000:
            out["steps"][str(upto)] = state_digest(lambda q: r.array("f%d" % q), lambda q: r.array("g%d" % q), r.array, masks)
    r.call("cal_saturation")
    r.call("monitor_breakthrough")
    fin = {"vol1_sum": r.get("vol1_sum"), "vol2_sum": r.get("vol2_sum"), "saturation_full_domain": r.get("saturation_full_domain"),
           "outlet_phase1_sum": int(r.get("outlet_phase1_sum"))}
    r.call("compute_macro_vars")
    fin.update(integrated(r.array("u"), r.array("v"), r.array("w"), r.array("rho"), r.array("phi"), masks))
    out["after_1000_steps"] = fin
    with open(os.path.join(HERE, "ref_c1_vectors.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print(json.dumps(fin, indent=1))


if __name__ == "__main__":
    main()
