"""Plumbing for tests/test_ref_pin.py: drive oracle/_ref (the reference's own subroutines, translated from
/root/reference by oracle/f2c_lite.py) from the state of an Oracle instance."""
import numpy as np

from oracle.oracle import FLUID_DTYPE, SOLID_DTYPE
from oracle.ref import Ref


def ref_from_oracle(o, fast=False):
    """Module variables and arrays of the reference program <- the oracle's state (after its setup): what the reference's
    own initialisation would leave behind for main_iteration_kernel."""
    p = o.p
    r = Ref("mp" if o.mp else "sp", fast=fast)
    r.set(nx=o.nx, ny=o.ny, nz=o.nz, nxglobal=p.nxG, nyglobal=p.nyG, nzglobal=p.nzG, idx=0, idy=0, idz=p.idz, npx=1, npy=1,
          npz=p.npz, id=0, iper=0, jper=p.jper, kper=p.kper, domain_wall_status_x_min=p.wsx0, domain_wall_status_x_max=p.wsx1,
          domain_wall_status_y_min=p.wsy0, domain_wall_status_y_max=p.wsy1, domain_wall_status_z_min=p.wsz0,
          domain_wall_status_z_max=p.wsz1, inlet_bc=p.inlet_BC, outlet_bc=p.outlet_BC, mpi_x=0, mpi_y=0, mpi_z=0, ix_async=0,
          iy_async=4, iz_async=4, steady_state_option=0, n_exclude_inlet=p.n_exclude_inlet, n_exclude_outlet=p.n_exclude_outlet,
          force_z=o.get_double("force_Z"), relaxation=o.get_double("relaxation"), uin_avg=o.get_double("uin_avg"),
          rho_in=o.get_double("rho_in"), rho_out=o.get_double("rho_out"))
    if o.mp:
        r.set(porous_plate_cmd=p.porous_plate_cmd, z_porous_plate=p.Z_porous_plate, la_nui1=o.get_double("la_nui1"),
              la_nui2=o.get_double("la_nui2"), gamma=p.gamma, beta=p.beta, phi_inlet=o.get_double("phi_inlet"),
              sa_inject=p.sa_inject, interface_z0=p.interface_z0,
              initial_fluid_distribution_option=p.initial_fluid_distribution_option)
        solid, fluid = o.solid_nodes(), o.fluid_nodes()
        r.set(num_solid_boundary=len(solid), num_fluid_boundary=len(fluid))
    else:
        r.set(la_nui=o.get_double("la_nui1"), **{k: o.get_double(k) for k in ("s_e", "s_e2", "s_q", "s_nu", "s_pi", "s_t")})
    r.call("memallocate_geometry", 1)
    r.call("memallocate_multi" if o.mp else "memallocate", 1)
    if o.mp:
        r.alloc("solid_boundary_nodes", (1, len(solid)))
        r.alloc("fluid_boundary_nodes", (1, len(fluid)))
        if len(solid):
            r.array("solid_boundary_nodes", SOLID_DTYPE)[:] = solid
        if len(fluid):
            r.array("fluid_boundary_nodes", FLUID_DTYPE)[:] = fluid
    r.array("walls")[...] = o.walls
    return r


def copy_state(r, o):
    for q in range(19):
        r.array("f%d" % q)[...] = o.f(q)
        if o.mp:
            r.array("g%d" % q)[...] = o.g(q)
    names = ["w_in"]
    if o.p.outlet_BC == 1:
        names += ["f_convec_bc"] + (["g_convec_bc", "phi_convec_bc"] if o.mp else [])
    if o.mp:
        names += ["phi"]
    for n in names:
        r.array(n)[...] = o.field(n)


STATE_MP = ("phi", "cn_x", "cn_y", "cn_z", "c_norm", "curv")


def assert_same_state(r, o, tag=""):
    """bit-for-bit: every population array and field, ghost layers included"""
    for q in range(19):
        assert np.array_equal(r.array("f%d" % q), o.f(q)), "%s f%d" % (tag, q)
        if o.mp:
            assert np.array_equal(r.array("g%d" % q), o.g(q)), "%s g%d" % (tag, q)
    names = list(STATE_MP) if o.mp else []
    if o.p.outlet_BC == 1:
        names += ["f_convec_bc"] + (["g_convec_bc", "phi_convec_bc"] if o.mp else [])
    for n in names:
        assert np.array_equal(r.array(n), o.field(n)), "%s %s" % (tag, n)
