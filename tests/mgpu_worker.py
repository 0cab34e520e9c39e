"""Worker of tests/test_multi_gpu.py: one rank per GPU (torchrun), one z slab each, NCCL halo exchange inside
libmflbm.so.  Sets the slab up from the CPU oracle's per-slab initial state (test infrastructure), runs the CUDA path
and writes the slab's interior fields for the parent test to compare with a single-domain oracle run."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    case = json.load(open(sys.argv[1]))
    outdir = sys.argv[2]
    import mflbm_b200 as M
    from importlib import import_module
    Ranks = import_module("mflbm_b200.dist").Ranks
    from helpers import ctx_from_oracle, make_oracle
    rk = Ranks(backend="gloo")  # rendezvous only; the data path uses the library's own NCCL communicator
    wg = np.load(case["walls"])
    o = make_oracle(walls_global=wg, npz=rk.world, idz=rk.rank, **case["oracle"])
    nid = rk.broadcast_bytes(M.nccl_unique_id() if rk.rank == 0 else b"", 128)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=case["layout"], device=rk.local_rank, use_nccl=1, nccl_unique_id=nid)
    if o.mp:
        ctx.color_gradient()
    # an output download in the middle of the run (save_phi / save_macro of the reference's driver): it grows the
    # library's staging buffer after the halo buffers exist, and the run must carry on unaffected
    half = case["steps"] // 2
    ctx.run(1, half)
    ctx.download(*(["phi", "f"] if o.mp else ["f"]))
    ctx.run(half + 1, case["steps"] - half)
    ctx.sync()
    names = ["f"] + (["g", "phi", "cn_x", "cn_y", "cn_z", "c_norm"] if o.mp else [])
    got = ctx.download(*names)
    out = {}
    for q in range(19):
        out["f%d" % q] = got["f"][q][1:-1, 1:-1, 1:-1]
        if o.mp:
            out["g%d" % q] = got["g"][q][1:-1, 1:-1, 1:-1]
    if o.mp:
        out["phi"] = got["phi"][4:-4, 4:-4, 4:-4]
        for n in ("cn_x", "cn_y", "cn_z", "c_norm"):
            out[n] = got[n][2:-2, 2:-2, 2:-2]
        v1, v2 = ctx.cal_saturation()
        out["sat"] = np.array([v1, v2])
    out["tk"] = ctx.monitor()["tk"]
    np.savez(os.path.join(outdir, "slab%d.npz" % rk.rank), **out)
    ctx.close()
    rk.close()


if __name__ == "__main__":
    main()
