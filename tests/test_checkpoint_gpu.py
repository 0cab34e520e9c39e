"""Checkpoint staging (MP/IO_multiphase.F90:562-642, MP/Init_multiphase.F90:477-557).

  * mflbm_checkpoint_begin freezes the state of the current step; the step loop goes on; mflbm_checkpoint_fetch later
    returns exactly what a blocking mflbm_download would have returned at that step (== the oracle at that step);
  * restarting from the fetched arrays (mflbm_upload + mflbm_color_gradient, MP/Main_multiphase.F90:120) continues bit
    for bit like the uninterrupted run (strict build);
  * the frozen-context mode (no room for a device snapshot) gives the same arrays and refuses to step meanwhile;
  * the driver mirror writes / reads the reference's per-rank stream file."""
import os

import numpy as np
import pytest

import mflbm_b200 as M
from helpers import compare_state, ctx_from_oracle, make_oracle, slot_mask, live_phi_mask

pytestmark = pytest.mark.gpu
LAYOUTS = [pytest.param(1, id="dense"), pytest.param(2, id="sparse")]


def _restart_ctx(o0, ck, layout):
    """new context on the same geometry, state = the checkpoint arrays (what initialization_old_multi reads back)"""
    ctx = ctx_from_oracle(o0, strict=True, kernel_variant=layout)  # uploads o0's (initial) state first ...
    ctx.upload(f=ck["f"], g=ck["g"], phi=ck["phi"], f_convec_bc=ck["f_convec_bc"], g_convec_bc=ck["g_convec_bc"],
               phi_convec_bc=ck["phi_convec_bc"])                   # ... then the checkpointed one over it
    ctx.color_gradient()
    return ctx


@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("direct", [False, True], ids=["staged", "frozen"])
def test_checkpoint_equals_state_at_begin_and_restart_is_bit_exact(layout, direct, monkeypatch):
    if direct:
        monkeypatch.setenv("MFLBM_CKPT_DIRECT", "1")
    o = make_oracle(modify_geometry_cmd=1, ca_0=5e-3)
    ctx = ctx_from_oracle(o, strict=True, kernel_variant=layout)
    o.color_gradient(); ctx.color_gradient()
    for t in range(1, 8):
        o.step(t)
    ctx.run(1, 7)
    mode = ctx.checkpoint_begin()
    assert mode == (1 if direct else 0)
    if direct:
        with pytest.raises(M.MflbmError):
            ctx.step(8)  # frozen until mflbm_checkpoint_end
    else:
        ctx.run(8, 5)    # the loop goes on while the snapshot waits to be fetched
    ck = ctx.checkpoint_fetch("f", "g")                 # two calls: arrays may be collected piecewise
    ck = ctx.checkpoint_fetch("phi", "f_convec_bc", "g_convec_bc", "phi_convec_bc", into=ck)
    ctx.checkpoint_end()
    with pytest.raises(M.MflbmError):
        ctx.checkpoint_end()
    # (1) the frozen state is the state after step 7
    sp = layout == 2
    for q in range(19):
        m = slot_mask(o, q) if sp else np.ones(o.f(q).shape, bool)
        assert np.array_equal(ck["f"][q][m], o.f(q)[m]) and np.array_equal(ck["g"][q][m], o.g(q)[m]), q
    pm = live_phi_mask(o) if sp else np.ones(o.field("phi").shape, bool)
    assert np.array_equal(ck["phi"][pm], o.field("phi")[pm])
    for n in ("f_convec_bc", "g_convec_bc", "phi_convec_bc"):
        assert np.array_equal(ck[n][1:-1, 1:-1], o.field(n)[1:-1, 1:-1]), n
    # (2) restart from it == uninterrupted run
    if direct:
        ctx.run(8, 5)
    for t in range(8, 13):
        o.step(t)
    compare_state(ctx, o, 0.0, sparse=sp)
    o0 = make_oracle(modify_geometry_cmd=1, ca_0=5e-3)
    ctx2 = _restart_ctx(o0, ck, layout)
    ctx2.run(8, 5)
    compare_state(ctx2, o, 0.0, sparse=sp)
    ctx.close(); ctx2.close()


def test_driver_checkpoint_file_roundtrip(tmp_path):
    """save_checkpoint / initialization_old_multi of the driver mirror: the reference's stream-file layout and a restart
    through it that continues like the uninterrupted run"""
    ctl = M.write_control_file(str(tmp_path / "simulation_control.txt"), multiphase=True, modify_geometry_cmd=1,
                               capillary_number="5000d-6")

    def driver():
        d = M.Driver(ctl)
        d.setup()
        d.create_context(kernel_variant=2)
        d.upload()
        d.color_gradient()
        return d

    a = driver()
    a.run(1, 6)
    path = str(tmp_path / "id0000")
    a.save_checkpoint(path, 6)
    nx, ny, nz = a.nx, a.ny, a.nz
    n1, n4, n2 = (nx + 2) * (ny + 2) * (nz + 2), (nx + 8) * (ny + 8) * (nz + 8), (nx + 2) * (ny + 2)
    assert os.path.getsize(path) == 4 + 8 + 8 + 8 * (38 * n1 + n4 + 2 * 19 * n2 + n2)
    hdr = np.fromfile(path, dtype="<i4", count=1)
    assert int(hdr[0]) == 7  # ntime + 1
    a.run(7, 6)
    a.sync()
    sat_a = a.cal_saturation()
    b = driver()
    assert b.initialization_old(path) == 7
    b.color_gradient()
    b.run(7, 6)
    b.sync()
    assert b.cal_saturation() == sat_a
    from mflbm_b200.binding import Arrays  # noqa: F401  (both contexts expose the same C ABI: compare phi through it)
    a.close(); b.close()
