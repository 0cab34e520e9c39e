"""world_size-2 gloo test (CPU) of the multi-slab host logic used by bench.py and the multi-GPU tests."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import mflbm_b200 as M
from conftest import ROOT


@pytest.mark.parametrize("periodic", [0, 1])
def test_two_ranks_gloo(tmp_path, periodic):
    M.build(); M.build_host()
    port = 29600 + (os.getpid() % 300) + periodic
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_cpu_worker.py"), str(tmp_path), str(periodic)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = [json.load(open(tmp_path / ("r%d.json" % k))) for k in range(2)]
    from importlib import import_module
    geo = import_module("mflbm_b200.geometry")
    w = geo.sphere_pack(24, 20, 64, porosity=0.5, rmin=3.0, rmax=6.0, seed=9, buffer=4, periodic=bool(periodic))
    w[0, :, :] = 1; w[-1, :, :] = 1; w[:, 0, :] = 1; w[:, -1, :] = 1   # domain_wall_status_x/y = 1 (template default)
    assert res[0]["total"] == res[1]["total"] == int((w == 0).sum())
    assert res[0]["local"] == int((w[:, :, :32] == 0).sum()) and res[1]["local"] == int((w[:, :, 32:] == 0).sum())
    assert all(x["id_ok"] and x["tmax"] == 2.0 and x["nz"] == 32 for x in res)
    assert (res[0]["first"], res[0]["last"], res[1]["first"], res[1]["last"]) == (1, 32, 33, 64)


def test_slab_partition_requires_equal_slabs():
    from importlib import import_module
    dist = import_module("mflbm_b200.dist")
    assert dist.slab_partition(1536, 8)[-1] == (1345, 1536)
    with pytest.raises(ValueError):
        dist.slab_partition(100, 8)
