"""Asynchronous output staging: mflbm_output_begin snapshots phi / u,v,w,rho at the current step and
copies them to the host while the step loop continues; mflbm_output_end hands the caller exactly what the blocking
path (compute_macro_vars + mflbm_download, i.e. the reference's save_macro / save_phi, MP/IO_multiphase.F90:646-712)
would have produced at that step, and the run itself is unaffected (state compared with the oracle, which performs the
same compute_macro_vars at the same step)."""
import numpy as np
import pytest

import mflbm_b200 as M
from helpers import compare_state, ctx_from_oracle, make_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("layout", [pytest.param(1, id="dense"), pytest.param(2, id="sparse")])
def test_staged_output_equals_blocking_download(layout):
    o = make_oracle(modify_geometry_cmd=1)
    a = ctx_from_oracle(o, strict=True, kernel_variant=layout)   # asynchronous path
    b = ctx_from_oracle(o, strict=True, kernel_variant=layout)   # blocking path
    o.color_gradient(); a.color_gradient(); b.color_gradient()
    for n in range(1, 7):
        o.step(n)
    a.run(1, 6); b.run(1, 6)
    a.output_begin(a.OUT_PHI | a.OUT_MACRO)
    a.run(7, 6)                                                   # the step loop goes on while the copy is in flight
    b.compute_macro_vars()
    ref = b.download("phi", "u", "v", "w", "rho")
    got = a.output_end("phi", "u", "v", "w", "rho")
    for nme in ("phi", "u", "v", "w", "rho"):
        assert np.array_equal(got[nme], ref[nme]), nme
    # the oracle's macroscopic fields at step 6 (valid after an even step)
    o.compute_macro_vars()
    inner = (slice(1, -1),) * 3  # compute_macro_vars covers 1..n; the ghost layer keeps each side's initial value
    for nme in ("u", "v", "w", "rho"):
        r = o.field(nme)[inner]
        assert np.max(np.abs(got[nme][inner] - r)) <= 1e-13 * max(1e-30, np.max(np.abs(r))), nme
    # ... and the run carried on as if nothing had happened
    for n in range(7, 13):
        o.step(n)
    compare_state(a, o, 0.0, sparse=layout == 2)
    with pytest.raises(M.MflbmError, match="no output in flight"):
        a.output_end("phi")
    a.close(); b.close()


def test_staged_output_singlephase_and_errors():
    rng = np.random.default_rng(9)
    wg = (rng.random((24, 20, 28)) < 0.2).astype(np.int8)
    o = make_oracle(multiphase=0, nxG=24, nyG=20, nzG=28, la_nu1=0.1, kper=1, force_z0=1e-5, walls_global=wg, n_exclude_inlet=0,
                    n_exclude_outlet=0)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=2)
    ctx.run(1, 4)
    with pytest.raises(M.MflbmError, match="multiphase field"):
        ctx.output_begin(ctx.OUT_PHI)
    ctx.output_begin(ctx.OUT_MACRO)
    with pytest.raises(M.MflbmError, match="already in flight"):
        ctx.output_begin(ctx.OUT_MACRO)
    ctx.run(5, 4)
    got = ctx.output_end("w", "rho")
    for n in range(1, 5):
        o.step(n)
    o.compute_macro_vars()
    inner = (slice(1, -1),) * 3
    for nme in ("w", "rho"):
        r = o.field(nme)[inner]
        assert np.max(np.abs(got[nme][inner] - r)) <= 1e-13 * max(1e-30, np.max(np.abs(r))), nme
    ctx.close()


def test_streamed_steps_equal_blocking_steps():
    """mflbm_step_streamed (host inlet profile copied beside the previous step, saturation sums handed out one call later) leaves
    exactly the state of mflbm_upload(w_in) + mflbm_step, and every step's sums equal mflbm_cal_saturation after that step"""
    import ctypes as C
    import torch
    o = make_oracle(modify_geometry_cmd=1, ca_0=5e-3)
    a = ctx_from_oracle(o, strict=True, kernel_variant=2)
    b = ctx_from_oracle(o, strict=True, kernel_variant=2)
    a.color_gradient(); b.color_gradient()
    w0 = np.asfortranarray(o.field("w_in")).copy(order="F")
    pin = torch.zeros(w0.size, dtype=torch.float64).pin_memory()
    want, got = [], []
    for t in range(1, 14):
        w = w0 * (1.0 + 0.01 * t)  # a time-dependent inlet profile
        a.upload(w_in=w)
        a.step(t)
        want.append(a.cal_saturation())
        if t > 1:  # the pinned buffer of step t-1 may be reused once that step's copy is done: the call for t-1 has returned
            b.sync()
        pin.copy_(torch.from_numpy(np.ascontiguousarray(w.ravel(order="F"))))
        r = b.step_streamed(t, pin.data_ptr())
        assert (r is None) == (t == 1)
        if r is not None:
            got.append(r)
    got.append(b.stream_flush())
    assert got == want
    fa, fb = a.download("f", "g", "phi"), b.download("f", "g", "phi")
    for n in ("f", "g"):
        for q in range(19):
            assert np.array_equal(fa[n][q], fb[n][q]), (n, q)
    assert np.array_equal(fa["phi"], fb["phi"])
    a.close(); b.close()
