"""Worker of tests/test_dist_cpu.py (gloo, no GPU): the host-side logic of an N-slab run -- windowed geometry per rank,
global pore count by all-reduce, broadcast of the 128-byte id that seeds the slab ring."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    outdir = sys.argv[1]
    import mflbm_b200 as M
    from importlib import import_module
    dist = import_module("mflbm_b200.dist")
    geo = import_module("mflbm_b200.geometry")
    rk = dist.Ranks(backend="gloo")
    nx, ny, nzG, periodic = 24, 20, 64, bool(int(sys.argv[2]))
    g = dict(porosity=0.5, rmin=3.0, rmax=6.0, seed=9, buffer=4)
    k0, k1 = M.Driver.window_range(rk.rank, rk.world, nzG, periodic)
    w = geo.sphere_pack_window(nx, ny, nzG, k0, k1, periodic=periodic, **g)
    ctl = M.write_control_file(os.path.join(outdir, "c%d.txt" % rk.rank), multiphase=True,
                               lattice_dimensions="%d,%d,%d" % (nx, ny, nzG), MPI_process_num="1,1,%d" % rk.world,
                               external_geometry_read_cmd=1, excluded_layers="4,4", periodic_indicator="0,0,%d" % int(periodic),
                               inlet_BC=0 if periodic else 1, outlet_BC=0 if periodic else 1, body_force_0="1d-5")
    d = M.Driver(ctl, idz=rk.rank, walls_window=(w, k0))
    d.setup()
    local = d.i64("pore_sum_local")
    total = int(round(rk.allreduce(float(local))))
    tmax = rk.allreduce(float(rk.rank + 1), "max")
    payload = bytes(range(128)) if rk.rank == 0 else b""
    got = rk.broadcast_bytes(payload, 128)
    first, last = dist.slab_partition(nzG, rk.world)[rk.rank]
    json.dump(dict(rank=rk.rank, world=rk.world, local=int(local), total=total, tmax=tmax, id_ok=got == bytes(range(128)),
                   nz=d.nz, first=first, last=last, num_solid=d.i64("num_solid_boundary")),
              open(os.path.join(outdir, "r%d.json" % rk.rank), "w"))
    d.close()
    rk.close()


if __name__ == "__main__":
    main()
