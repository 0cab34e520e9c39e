"""Shared test plumbing: build an oracle case and the matching GPU context through the C ABI."""
import numpy as np

import mflbm_b200 as M
from oracle.oracle import Oracle, default_params

PDF_NAMES = ["f", "g"]


def make_oracle(**kw):
    walls_global = kw.pop("walls_global", None)
    p = default_params(**kw)
    o = Oracle(p)
    o.setup(walls_global)
    return o


def ctx_from_oracle(o, strict=False, **over):
    """mflbm_create + mflbm_upload from the oracle's state (what the Fortran driver would hand over)."""
    p = o.p
    mp = bool(p.multiphase)
    solid = o.solid_nodes() if mp else np.zeros(0, M.SOLID_DTYPE)
    fluid = o.fluid_nodes() if mp else np.zeros(0, M.FLUID_DTYPE)
    cfg = dict(solver=1 if mp else 0, nx=o.nx, ny=o.ny, nz=o.nz, nxGlobal=p.nxG, nyGlobal=p.nyG, nzGlobal=p.nzG,
               idz=p.idz, npz=p.npz, jper=p.jper, kper=p.kper, domain_wall_status_z_min=p.wsz0,
               domain_wall_status_z_max=p.wsz1, inlet_BC=p.inlet_BC, outlet_BC=p.outlet_BC,
               porous_plate_cmd=p.porous_plate_cmd, Z_porous_plate=p.Z_porous_plate, mrt=p.mrt, iz_async=4,
               num_solid_boundary=len(solid), num_fluid_boundary=len(fluid),
               la_nui1=o.get_double("la_nui1"), la_nui2=o.get_double("la_nui2"), gamma=p.gamma, beta=p.beta,
               force_Z=o.get_double("force_Z"), phi_inlet=o.get_double("phi_inlet"), sa_inject=p.sa_inject,
               relaxation=o.get_double("relaxation"), uin_avg=o.get_double("uin_avg"), rho_in=o.get_double("rho_in"),
               rho_out=o.get_double("rho_out"), s_e=o.get_double("s_e"), s_e2=o.get_double("s_e2"),
               s_q=o.get_double("s_q"), s_nu=o.get_double("s_nu"), s_pi=o.get_double("s_pi"), s_t=o.get_double("s_t"))
    cfg.update(over)
    ctx = M.Context(strict=strict, **cfg)
    upload_from_oracle(ctx, o)
    return ctx


def upload_from_oracle(ctx, o):
    mp = o.mp
    arrays = dict(f=[o.f(q) for q in range(19)], walls=o.walls, w_in=o.field("w_in"), f_convec_bc=o.field("f_convec_bc"))
    if mp:
        arrays.update(g=[o.g(q) for q in range(19)], phi=o.field("phi"), g_convec_bc=o.field("g_convec_bc"),
                      phi_convec_bc=o.field("phi_convec_bc"), solid_boundary_nodes=o.solid_nodes(),
                      fluid_boundary_nodes=o.fluid_nodes())
    ctx.upload(**arrays)


def rel_err(a, b):
    """max |a-b| / max(|b|) -- field-level relative error (the populations span many decades)."""
    a = np.asarray(a)
    b = np.asarray(b)
    scale = np.max(np.abs(b))
    if scale == 0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - b)) / scale)


EX = [0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0]
EY = [0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1]
EZ = [0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1]


def fluid_mask(o):
    """(0:n+1)^3 boolean: nodes the collision kernel processes (walls==0 inside 1..n)."""
    w = o.walls[1:-1, 1:-1, 1:-1]
    a = np.zeros(w.shape, bool)
    a[1:-1, 1:-1, 1:-1] = w[1:-1, 1:-1, 1:-1] == 0
    return a


def active_mask(o):
    """fluid nodes plus every node of the 0..n+1 box they stream into (the sparse layout's storage set)."""
    a = fluid_mask(o)
    act = a.copy()
    nx, ny, nz = a.shape
    for q in range(1, 19):
        src = a[max(0, -EX[q]):nx - max(0, EX[q]), max(0, -EY[q]):ny - max(0, EY[q]), max(0, -EZ[q]):nz - max(0, EZ[q])]
        act[max(0, EX[q]):nx - max(0, -EX[q]), max(0, EY[q]):ny - max(0, -EY[q]), max(0, EZ[q]):nz - max(0, -EZ[q])] |= src
    return act


def slot_mask(o, q):
    """Sparse layout: cells whose slot q is LIVE storage = the fluid nodes (the even step reads all 19 own slots) and every
    non-fluid cell y of the 0..n+1 box whose slot q a fluid node streams through, i.e. y + e_q is fluid (the only node that
    reads or writes slot q at y in the odd step is y + e_q; the inlet / outlet kernels and the z exchange feed exactly those
    slots).  All other slots are dead storage: nothing ever consumes them, the reference's exchange merely copies them
    around (e.g. into the slots of a solid node behind the periodic seam), and mflbm_download may leave them untouched."""
    a = fluid_mask(o)
    nx, ny, nz = a.shape
    m = a.copy()
    if q:
        # y + e_q in A  <=>  shift the fluid mask by -e_q
        ex, ey, ez = EX[q], EY[q], EZ[q]
        src = a[max(0, ex):nx - max(0, -ex), max(0, ey):ny - max(0, -ey), max(0, ez):nz - max(0, -ez)]
        m[max(0, -ex):nx - max(0, ex), max(0, -ey):ny - max(0, ey), max(0, -ez):nz - max(0, ez)] |= src
    if o.p.jper:
        # y-periodic: the population storage of the y ghost rows is aliased to the periodic images (the exchange is part of
        # the adjacency), so mflbm_download leaves those rows of the caller's arrays alone -- where z is exchanged too, and
        # in the rows k = 1..nz otherwise (the reference exchanges no others, MP/Mpi.F90:147-180)
        ks = slice(None) if o.p.kper else slice(1, -1)
        m[:, 0, ks] = False
        m[:, -1, ks] = False
    return m


def compare_state(ctx, o, tol, fields=("phi", "cn_x", "cn_y", "cn_z", "c_norm", "curv"), sparse=False):
    """Compare all populations (+ fields) of the GPU context with the oracle.  Returns the worst error.

    sparse=True: populations are compared on the active-node set only (the rest of the caller's array is left
    untouched by mflbm_download) and the curvature on fluid nodes only (its values at solid nodes
    have no consumer)."""
    names = ["f"] + (["g"] if o.mp else [])
    names += [n for n in fields if o.mp]
    got = ctx.download(*names)
    report = {}
    act = active_mask(o) if sparse else None
    fl = fluid_mask(o) if sparse else None

    def err(a, b, m):
        return rel_err(a[m], b[m]) if m is not None else rel_err(a, b)

    for q in range(19):
        mq = slot_mask(o, q) if sparse else None
        report["f%d" % q] = err(got["f"][q], o.f(q), mq)
        if o.mp:
            report["g%d" % q] = err(got["g"][q], o.g(q), mq)
    where = {}
    if o.mp:
        for n in fields:
            m = fl if (n == "curv" and sparse) else None
            if n == "phi" and sparse:
                m = live_phi_mask(o)
            report[n] = err(got[n], o.field(n), m)
            if not (report[n] <= tol):  # first few offending cells, as array indices of the field
                d = np.abs(got[n] - o.field(n)) > tol * max(1e-300, np.max(np.abs(o.field(n))))
                if m is not None:
                    d &= m
                where[n] = np.argwhere(d)[:6].tolist()
    worst = max(report.values())
    bad = {k: v for k, v in report.items() if not (v <= tol)}
    assert not bad, "fields beyond tol %g: %s at %s" % (tol, bad, where)
    return worst


def live_phi_mask(o):
    """phi(-3:n+4)^3 cells with a consumer.  Dead storage: the outermost z ghost planes (k = -3 and k = nz+4) in columns
    whose boundary node is solid.  The inlet / outlet kernels copy phi(k=0) / phi(k=nz+1) there (wall_indicator blend,
    MP/Boundary_multiphase_inlet.F90:26-29, _outlet.F90:33-40); the cell is solid, lies outside the solid-boundary list
    (3 ghost layers) and outside every gradient stencil (K4 runs on -1..n+2), so nothing ever reads it.  Its value is the
    PREVIOUS step's K3 result of the cell below, which an implementation that skips K3 where phi is uniform cannot (and
    need not) reproduce."""
    phi = o.field("phi")
    m = np.ones(phi.shape, bool)
    w = o.walls  # (-1:n+2)
    nz = o.nz
    lo = w[2:-2, 2:-2, 2 + 0] != 0      # walls(i,j,1), i,j in 1..n
    hi = w[2:-2, 2:-2, 2 + nz - 1] != 0  # walls(i,j,nz)
    m[4:-4, 4:-4, 0][lo] = False
    m[4:-4, 4:-4, -1][hi] = False
    if o.p.kper and o.p.npz == 1:
        # z-periodic: the same two planes are images of the planes nz-3 and 4 (MP/Mpi.F90:624-631), copied BEFORE this step's
        # K3: at solid cells they hold the previous step's K3 result of the image cell -- equally without a consumer (solid, so
        # no K3 of plane -2 / nz+3 reads them; outside the list; outside every gradient stencil), and the fused chain kernel
        # (csrc/march.cuh) never stores K3 results at all, so only the fluid cells of these planes are compared
        fl_lo = np.zeros(phi.shape[:2], bool)
        fl_hi = np.zeros(phi.shape[:2], bool)
        fl_lo[2:-2, 2:-2] = w[:, :, 2 + (nz - 3) - 1] == 0   # walls(i,j,nz-3), i,j in -1..n+2
        fl_hi[2:-2, 2:-2] = w[:, :, 2 + 4 - 1] == 0          # walls(i,j,4)
        m[:, :, 0] &= fl_lo
        m[:, :, -1] &= fl_hi
    if o.p.jper:
        # y-periodic: likewise the outermost ghost rows j = -3 and j = ny+4, images of the rows ny-3 and 4 (MP/Mpi.F90:633-790)
        ny = o.ny
        fl_lo = np.zeros((phi.shape[0], phi.shape[2]), bool)
        fl_hi = np.zeros((phi.shape[0], phi.shape[2]), bool)
        fl_lo[2:-2, 2:-2] = w[:, 2 + (ny - 3) - 1, :] == 0
        fl_hi[2:-2, 2:-2] = w[:, 2 + 4 - 1, :] == 0
        m[:, 0, :] &= fl_lo
        m[:, -1, :] &= fl_hi
    return m
