"""ctypes binding of oracle/_ref: the reference's own subroutines, translated mechanically from /root/reference by
oracle/f2c_lite.py and compiled with gcc (TEST INFRASTRUCTURE ONLY -- same rules as oracle/oracle.py).

The library is built by ``make -C oracle ref`` wherever /root/reference is present (this container); the built
``oracle/_ref/libmflbm_ref_{mp,sp}.so`` travels to the GPU box with the snapshot.  ``available()`` says whether it can be
loaded; nothing reads /root/reference at run time.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
REFERENCE_ROOT = "/root/reference"


class _Desc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("base", C.c_void_p), ("lo", C.c_longlong * 4), ("n", C.c_longlong * 4), ("rank", C.c_int),
                ("elem", C.c_int)]


def lib_path(solver, fast=False):
    return os.path.join(REF_DIR, "libmflbm_ref_%s%s.so" % (solver, "_fast" if fast else ""))


def build(solver="mp", fast=False):
    """translate + compile (needs /root/reference); returns the .so path or None"""
    if not os.path.isdir(REFERENCE_ROOT):
        return lib_path(solver, fast) if os.path.exists(lib_path(solver, fast)) else None
    subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return lib_path(solver, fast)


def available(solver="mp", fast=False):
    return os.path.exists(lib_path(solver, fast)) or os.path.isdir(REFERENCE_ROOT)


_LIBS = {}


def _load(solver, fast=False):
    key = (solver, fast)
    if key in _LIBS:
        return _LIBS[key]
    path = build(solver, fast)
    if path is None or not os.path.exists(path):
        raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
    lib = C.CDLL(path)
    lib.ref_set_int.argtypes = [C.c_char_p, C.c_longlong]
    lib.ref_set_double.argtypes = [C.c_char_p, C.c_double]
    lib.ref_get.argtypes = [C.c_char_p]
    lib.ref_get.restype = C.c_double
    lib.ref_has.argtypes = [C.c_char_p]
    lib.ref_array.argtypes = [C.c_char_p]
    lib.ref_array.restype = C.POINTER(_Desc)
    lib.ref_alloc.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    lib.ref_call.argtypes = [C.c_char_p, C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
    lib.ref_sub_name.argtypes = [C.c_int]
    lib.ref_sub_name.restype = C.c_char_p
    _LIBS[key] = lib
    return lib


class Ref:
    """The module state of ONE reference solver ("mp" = multiphase_3D, "sp" = singlephase_3D) in this process.
    Module variables are process globals, exactly like in the Fortran program: one instance at a time."""

    def __init__(self, solver="mp", fast=False):
        self.lib = _load(solver, fast)
        self.solver = solver

    def subroutines(self):
        out, i = [], 0
        while True:
            n = self.lib.ref_sub_name(i)
            if not n:
                return out
            out.append(n.decode())
            i += 1

    def set(self, **kw):
        for k, v in kw.items():
            name = k.lower().encode()
            if isinstance(v, (bool, int, np.integer)):
                rc = self.lib.ref_set_int(name, int(v))
            else:
                rc = self.lib.ref_set_double(name, float(v))
            if rc:
                raise KeyError("module variable %s (rc=%d)" % (k, rc))

    def get(self, name):
        if self.lib.ref_has(name.lower().encode()) != 1:
            raise KeyError(name)
        return self.lib.ref_get(name.lower().encode())

    def call(self, name, *args):
        ia = [int(a) for a in args if isinstance(a, (int, np.integer))]
        da = [float(a) for a in args if isinstance(a, float)]
        IA = (C.c_longlong * max(len(ia), 1))(*ia)
        DA = (C.c_double * max(len(da), 1))(*da)
        if self.lib.ref_call(name.lower().encode(), IA, DA):
            raise KeyError("subroutine %s was not translated" % name)

    def alloc(self, name, *bounds):
        lo = (C.c_longlong * len(bounds))(*[b[0] for b in bounds])
        hi = (C.c_longlong * len(bounds))(*[b[1] for b in bounds])
        if self.lib.ref_alloc(name.lower().encode(), len(bounds), lo, hi):
            raise KeyError(name)

    def array(self, name, dtype=None):
        """numpy view (Fortran order, the reference's own bounds) of a module array"""
        d = self.lib.ref_array(name.lower().encode())
        if not d or not d.contents.base:
            raise KeyError("array %s is not allocated" % name)
        d = d.contents
        shape = tuple(int(d.n[m]) for m in range(d.rank))
        n = int(np.prod(shape))
        if dtype is None:
            dtype = {8: np.float64, 1: np.int8, 4: np.int32}[d.elem]
        dt = np.dtype(dtype)
        buf = (C.c_char * (n * d.elem)).from_address(d.base)
        a = np.frombuffer(buf, dtype=dt)
        if dt.itemsize == d.elem and dt.fields is None:
            return a.reshape(shape, order="F")
        return a  # record arrays (derived types): one record per element

    def lower_bounds(self, name):
        d = self.lib.ref_array(name.lower().encode()).contents
        return tuple(int(d.lo[m]) for m in range(d.rank))
