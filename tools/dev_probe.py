"""Developer probe (NOT product, NOT bench): quick MLUPS of the CUDA path on a synthetic case, set up with the oracle."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import make_oracle, ctx_from_oracle

def spheres(n, nz, porosity, seed=1, rmin=6, rmax=14):
    rng = np.random.default_rng(seed)
    w = np.zeros((n, n, nz), np.int8)
    core = slice(10, nz - 10)
    while (w[:, :, core] == 0).mean() > porosity:
        for _ in range(50):
            r = rng.uniform(rmin, rmax); c = rng.uniform(0, [n, n, nz])
            lo = np.maximum(np.floor(c - r).astype(int), [0, 0, 10]); hi = np.minimum(np.ceil(c + r).astype(int) + 1, [n, n, nz - 10])
            if np.any(hi <= lo): continue
            g = np.ogrid[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
            m = (g[0] - c[0]) ** 2 + (g[1] - c[1]) ** 2 + (g[2] - c[2]) ** 2 <= r * r
            w[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]][m] = 1
    return w

def run(mp, n, nz, porosity, steps=100):
    wg = spheres(n, nz, porosity) if porosity < 1 else None
    t = time.time()
    kw = dict(multiphase=1 if mp else 0, nxG=n, nyG=n, nzG=nz, walls_global=wg)
    if mp: kw.update(la_nu2=0.04)
    else: kw.update(la_nu1=0.1, kper=1, force_z0=1e-5, n_exclude_inlet=0, n_exclude_outlet=0)
    o = make_oracle(**kw)
    pore = o.get_i64("pore_sum")
    ctx = ctx_from_oracle(o)
    print("setup %.1fs pore_sum=%d porosity=%.3f dev_bytes=%.2f GB" % (time.time() - t, pore, pore / (n * n * nz), ctx.device_bytes / 1e9), flush=True)
    if mp: ctx.color_gradient()
    ctx.run(1, 20); ctx.sync()
    best = 1e9
    for r in range(3):
        ctx.timer_start(); ctx.run(1, steps); ms = ctx.timer_stop(); best = min(best, ms)
    mlups = pore * steps / (best * 1e-3) / 1e6
    bpn = 624 if mp else 304
    print("%s n=%d nz=%d porosity=%.2f: %.3f ms/step  %.0f MLUPS (fluid)  %.0f GB/s algorithmic  dense-MLUPS %.0f" % (
        "MP" if mp else "SP", n, nz, porosity, best / steps, mlups, mlups * bpn / 1e3, n * n * nz * steps / (best * 1e-3) / 1e6), flush=True)
    ctx.close(); o.close()

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    if which in ("all", "mp"): run(True, n, n, 1.0)
    if which in ("all", "sp"): run(False, n, n, 1.0)
    if which in ("all", "mpp"): run(True, n, n, 0.36)
    if which in ("all", "spp"): run(False, n, n, 0.2)
