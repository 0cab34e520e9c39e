"""N-GPU == 1-domain parity: z-slab decomposition with NCCL halo exchange of the
10 (5) population planes and 4 phi planes per direction, boundary slabs first and overlapped with the interior
(MP/Main_multiphase.F90:358-387).  Each rank runs one slab on its own GPU; the union of the slab interiors must equal
the single-domain CPU oracle bit-exactly in the strict build (identical arithmetic per node)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from helpers import make_oracle

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


CASES = {
    "mp_open": dict(oracle=dict(nxG=20, nyG=18, nzG=32, la_nu2=0.04, interface_z0=6.0, n_exclude_inlet=0, n_exclude_outlet=0)),
    "mp_periodic": dict(oracle=dict(nxG=20, nyG=18, nzG=32, kper=1, force_z0=2e-4, la_nu2=0.04, initial_fluid_distribution_option=5,
                                    interface_z0=6.0, n_exclude_inlet=0, n_exclude_outlet=0)),
    "sp_periodic": dict(oracle=dict(multiphase=0, nxG=20, nyG=18, nzG=32, kper=1, force_z0=1e-5, la_nu1=0.1, n_exclude_inlet=0,
                                    n_exclude_outlet=0)),
    # y-periodic lattices on z slabs: the x edges between ranks come out of the y wrap in the adjacency + the z halo planes
    "mp_yz_periodic": dict(oracle=dict(nxG=20, nyG=18, nzG=32, jper=1, kper=1, wsy0=0, wsy1=0, inlet_BC=0, outlet_BC=0, force_z0=2e-4,
                                       la_nu2=0.04, initial_fluid_distribution_option=3, interface_z0=7.0, n_exclude_inlet=0,
                                       n_exclude_outlet=0)),
    "mp_y_periodic_open": dict(oracle=dict(nxG=20, nyG=18, nzG=32, jper=1, wsy0=0, wsy1=0, la_nu2=0.04, interface_z0=6.0,
                                           n_exclude_inlet=0, n_exclude_outlet=0)),
}


@pytest.mark.parametrize("layout", [1, 2], ids=["dense", "sparse"])
@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("npz", [2, 4])
def test_slabs_match_single_domain(tmp_path, case, layout, npz):
    if _ngpu() < npz:
        pytest.skip("needs %d GPUs" % npz)
    if npz == 4 and case != "mp_open":
        pytest.skip("4-slab run only for the open multiphase case")
    if layout == 1 and "y_" in case or layout == 1 and "yz_" in case:
        pytest.skip("y-periodic lattices always run the sparse layout")
    spec = dict(CASES[case], layout=layout, steps=9)
    rng = np.random.default_rng(3)
    n = spec["oracle"]
    wg = (rng.random((n["nxG"], n["nyG"], n["nzG"])) < 0.25).astype(np.int8)
    wg[:, :, :3] = 0
    wg[:, :, -3:] = 0
    np.save(tmp_path / "walls.npy", wg)
    spec["walls"] = str(tmp_path / "walls.npy")
    json.dump(spec, open(tmp_path / "case.json", "w"))
    port = 29500 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(npz), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"), str(tmp_path / "case.json"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    # single-domain oracle
    o = make_oracle(walls_global=wg, **spec["oracle"])
    if o.mp:
        o.color_gradient()
    for t in range(1, spec["steps"] + 1):
        o.step(t)
    nzl = n["nzG"] // npz
    fluid = o.walls[2:-2, 2:-2, 2:-2] == 0
    for rnk in range(npz):
        d = np.load(tmp_path / ("slab%d.npz" % rnk))
        ks = slice(rnk * nzl, (rnk + 1) * nzl)
        m = fluid[:, :, ks]
        for q in range(19):
            for fam in (("f", o.f), ("g", o.g)) if o.mp else (("f", o.f),):
                ref = fam[1](q)[1:-1, 1:-1, 1:-1][:, :, ks]
                assert np.array_equal(d["%s%d" % (fam[0], q)][m], ref[m]), (case, rnk, fam[0], q)
        if o.mp:
            assert np.array_equal(d["phi"][m], o.field("phi")[4:-4, 4:-4, 4:-4][:, :, ks][m])
            for nm in ("cn_x", "cn_y", "cn_z", "c_norm"):
                assert np.array_equal(d[nm][m], o.field(nm)[2:-2, 2:-2, 2:-2][:, :, ks][m]), (case, rnk, nm)
    # monitors: per-slab tk profiles concatenate to the single-domain profiles (per-plane sums: same order -> 1e-12)
    mo = o.monitor()
    names = ("fl1", "fl2", "vol1", "vol2", "mass1", "mass2", "pre") if o.mp else ("fl1", "pre")  # the oracle keeps SP "fl" in fl1
    for idx, nm in enumerate(names):
        got = np.concatenate([np.load(tmp_path / ("slab%d.npz" % rnk))["tk"][idx * nzl:(idx + 1) * nzl] for rnk in range(npz)])
        ref = o.field(nm)
        assert np.max(np.abs(got - ref)) <= 1e-10 * max(1e-30, np.max(np.abs(ref))), nm
