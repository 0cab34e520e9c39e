#!/usr/bin/env python
"""Regenerates the golden INPUT fixtures from the reference's shipped geometry files (run in the container where
/root/reference is mounted; the GPU box only sees the committed outputs).

  bentheimer_in10_240_out10.bits.xz   MF-LBM-extFiles/geometry_files/sample_rock_geometry_wallarray/
                                      bentheimer_in10_240_240_240_out10.dat (240x240x260 int8 wall array, the geometry of
                                      the reference's own benchmark cases 6 and 8): one bit per node, i fastest, lzma.
                                      Known answer: 3 670 813 pore nodes.
  tube_sphere.npz                     MF-LBM-extFiles/geometry_files/tube_sphere_example/tube_sphere.dat (60x60x80) with the
                                      list checksums the oracle produced for it (written by an earlier session of this
                                      round; kept as is).
"""
import lzma
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference/MF-LBM-extFiles/geometry_files"


def main():
    from oracle.oracle import read_wall_array
    w = read_wall_array(os.path.join(REF, "sample_rock_geometry_wallarray", "bentheimer_in10_240_240_240_out10.dat"))
    assert w.shape == (240, 240, 260) and int((w == 0).sum()) == 3670813
    bits = np.packbits(w.ravel(order="F"))
    with open(os.path.join(HERE, "bentheimer_in10_240_out10.bits.xz"), "wb") as fh:
        fh.write(lzma.compress(bits.tobytes(), preset=9))
    print("bentheimer: %d bytes" % os.path.getsize(os.path.join(HERE, "bentheimer_in10_240_out10.bits.xz")))


if __name__ == "__main__":
    main()
