"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/mflbm.h declares, struct layouts agree between C and the ctypes mirror, and the product refuses
to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

import mflbm_b200 as M
from conftest import HAS_GPU, ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mflbm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mflbm_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    M.build()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for strict in (False, True):
        lib = M.load(strict=strict)
        for s in syms:
            assert hasattr(lib, s), s
    assert set(syms) == set(M.EXPORTS)


def test_struct_layouts_match_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mflbm.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(mflbm_config),sizeof(mflbm_arrays),sizeof(mflbm_solid_node),sizeof(mflbm_fluid_node),'
                   'offsetof(mflbm_config,la_nui1),offsetof(mflbm_config,nccl_unique_id),sizeof(mflbm_geometry_config),'
                   'offsetof(mflbm_geometry_config,theta));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got == [ctypes.sizeof(M.Config), ctypes.sizeof(M.Arrays), M.SOLID_DTYPE.itemsize, M.FLUID_DTYPE.itemsize,
                   M.Config.la_nui1.offset, M.Config.nccl_unique_id.offset, ctypes.sizeof(M.GeometryConfig),
                   M.GeometryConfig.theta.offset]
    assert M.SOLID_DTYPE.itemsize == 96 and M.FLUID_DTYPE.itemsize == 48


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(M.MflbmError, match="no CUDA device"):
        M.Context(solver=1, nx=8, ny=8, nz=8)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mf-lbm_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) in ("build", "lib"):
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in txt.lower().replace("cpu oracle", "").replace("the oracle", ""), (dirpath, fn)


def test_fortran_offset_table_is_current(tmp_path):
    """fortran/mflbm_abi_offsets.txt (what a maintainer compares the type, bind(c) declarations with) is the output of
    tools/gen_abi_offsets.c for the header as it is now, and the ctypes mirror agrees with it"""
    import subprocess
    exe = str(tmp_path / "gen")
    subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "gen_abi_offsets.c"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    assert out == open(os.path.join(ROOT, "fortran", "mflbm_abi_offsets.txt")).read()
    import mflbm_b200 as M
    table = {}
    for line in out.splitlines()[1:]:
        st, mem, off, size = line.split()
        table[(st, mem)] = (off, int(size))
    for name, typ in (("mflbm_config", M.Config), ("mflbm_arrays", M.Arrays), ("mflbm_geometry_config", M.GeometryConfig)):
        assert table[(name, "(sizeof)")][1] == __import__("ctypes").sizeof(typ)
        for fname, _ in typ._fields_:
            if (name, fname) in table:
                assert int(table[(name, fname)][0]) == getattr(typ, fname).offset, (name, fname)
