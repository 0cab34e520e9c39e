"""Synthetic porous geometries of the BASELINE.json configs, written as reference wall arrays.

All generators are deterministic functions of (shape, seed): the sphere list is drawn once with
``numpy.random.Generator(PCG64(seed))`` for the GLOBAL lattice, so every rank of a multi-GPU run can rasterise
just its own z window and still see the same medium.  The number of overlapping spheres follows the Boolean
(Poisson) model, porosity = exp(-n E[V] / V), which lands within ~0.005 of the target on these sizes; the realised
porosity is always reported next to the numbers.
"""
import math

import numpy as np


def sphere_list(nx, ny, nz, porosity, rmin, rmax, seed, buffer):
    """Centres/radii of overlapping solid spheres filling the core nz-2*buffer planes to ~porosity."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ev = 4.0 / 3.0 * math.pi * (rmax ** 4 - rmin ** 4) / (4.0 * (rmax - rmin))  # E[4/3 pi r^3], r ~ U[rmin,rmax]
    # centres live in the core extended by rmax so that the core sees a homogeneous Boolean model
    lo = np.array([-rmax, -rmax, buffer - rmax], dtype=float)
    hi = np.array([nx + rmax, ny + rmax, nz - buffer + rmax], dtype=float)
    vol = float(np.prod(hi - lo))
    n = int(round(-math.log(porosity) * vol / ev))
    c = rng.uniform(lo, hi, size=(n, 3))
    r = rng.uniform(rmin, rmax, size=n)
    return c, r


def rasterize(nx, ny, k0, k1, centres, radii, nz, buffer):
    """int8 walls for the global planes k0..k1 (1-based, inclusive): 1 inside a sphere, 0 elsewhere; the
    ``buffer`` planes at each end of the lattice stay fluid (inlet / outlet reservoirs)."""
    nk = k1 - k0 + 1
    w = np.zeros((nx, ny, nk), dtype=np.int8)
    zlo, zhi = max(k0, buffer + 1), min(k1, nz - buffer)  # planes that may hold solid
    if zhi < zlo:
        return w
    sel = np.nonzero((centres[:, 2] + radii >= zlo - 0.5) & (centres[:, 2] - radii <= zhi + 0.5))[0]
    for n in sel:
        cx, cy, cz = centres[n]
        r = radii[n]
        # cell (i,j,k) has coordinates (i,j,k), 1-based like the reference
        i0, i1 = max(1, int(math.ceil(cx - r))), min(nx, int(math.floor(cx + r)))
        j0, j1 = max(1, int(math.ceil(cy - r))), min(ny, int(math.floor(cy + r)))
        z0, z1 = max(zlo, int(math.ceil(cz - r))), min(zhi, int(math.floor(cz + r)))
        if i1 < i0 or j1 < j0 or z1 < z0:
            continue
        gi = np.arange(i0, i1 + 1)[:, None, None] - cx
        gj = np.arange(j0, j1 + 1)[None, :, None] - cy
        gk = np.arange(z0, z1 + 1)[None, None, :] - cz
        m = gi * gi + gj * gj + gk * gk <= r * r
        sub = w[i0 - 1:i1, j0 - 1:j1, z0 - k0:z1 - k0 + 1]
        sub[m] = 1
    return w


def sphere_pack_window(nx, ny, nz, k0, k1, porosity=0.36, rmin=8.0, rmax=20.0, seed=1, buffer=10, periodic=False):
    """Planes k0..k1 of the global sphere pack; on a periodic lattice k may leave 1..nz and is wrapped."""
    c, r = sphere_list(nx, ny, nz, porosity, rmin, rmax, seed, buffer)
    if not periodic or (k0 >= 1 and k1 <= nz):
        return rasterize(nx, ny, k0, k1, c, r, nz, buffer)
    planes = [((k - 1) % nz) + 1 for k in range(k0, k1 + 1)]
    out = np.zeros((nx, ny, len(planes)), dtype=np.int8)
    lo, hi = min(planes), max(planes)
    full = rasterize(nx, ny, lo, hi, c, r, nz, buffer)
    for n, k in enumerate(planes):
        out[:, :, n] = full[:, :, k - lo]
    return out


def stacked_window(nx, ny, unit_nz, k0, k1, **kw):
    """Planes k0..k1 of a lattice made of identical copies of ONE unit_nz-plane sphere pack (its buffer layers included)
    stacked along z: plane k of the stack is plane ((k-1) mod unit_nz)+1 of the unit.  Every z slab of unit_nz planes
    then holds exactly the same medium (equal fluid-node counts: weak-scaling numbers compare communication, not
    porosity drift)."""
    unit = sphere_pack_window(nx, ny, unit_nz, 1, unit_nz, **kw)
    idx = [(k - 1) % unit_nz for k in range(k0, k1 + 1)]
    return np.ascontiguousarray(unit[:, :, idx]) if idx != list(range(unit_nz)) else unit


def sphere_pack(nx, ny, nz, **kw):
    return sphere_pack_window(nx, ny, nz, 1, nz, **kw)


def load_packed_walls(path, shape):
    """Wall array stored one bit per node (i fastest) and lzma-compressed, e.g. tests/golden/bentheimer_in10_240_out10.bits.xz
    (the reference's own Bentheimer geometry, see tests/golden/make_fixtures.py)."""
    import lzma
    with open(path, "rb") as fh:
        bits = np.frombuffer(lzma.decompress(fh.read()), dtype=np.uint8)
    n = int(np.prod(shape))
    return np.unpackbits(bits)[:n].astype(np.int8).reshape(shape, order="F")
