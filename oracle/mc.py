"""ORACLE -- test infrastructure only.  ctypes front end of oracle/mc_oracle.c (CPU marching cubes).
Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU baseline legs."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libr3g_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "mc_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "r3g_mc_tables.h")
    if (not force and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(src)
            and os.path.getmtime(_LIB) >= os.path.getmtime(hdr)):
        return _LIB
    os.makedirs(os.path.dirname(_LIB), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-shared", "-fPIC", src, "-lm", "-o", _LIB],
                   check=True)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.r3g_oracle_marching_cubes.restype = C.c_int
        _lib.r3g_oracle_marching_cubes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                                                   C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                                   C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_void_p]
        _lib.r3g_oracle_free.argtypes = [C.c_void_p]
    return _lib


def marching_cubes(volume, level=0.0, bounds=None, return_cases=False):
    """Mirror of skimage.measure.marching_cubes(volume, level, method='lewiner')[:2] (+ optional Hunyuan
    rescale).  Raises ValueError / RuntimeError where skimage does."""
    lib = _load()
    vol = np.ascontiguousarray(volume, dtype=np.float32)
    n0, n1, n2 = vol.shape
    vp, fp = C.c_void_p(), C.c_void_p()
    nv, nf = C.c_int64(), C.c_int64()
    cases = np.zeros(((n0 - 1) * (n1 - 1) * (n2 - 1),), np.uint8) if return_cases else None
    b = None
    if bounds is not None:
        b = np.ascontiguousarray(bounds, dtype=np.float64)
    rc = lib.r3g_oracle_marching_cubes(vol.ctypes.data, n0, n1, n2, float(level),
                                       b.ctypes.data if b is not None else None, C.byref(vp), C.byref(nv),
                                       C.byref(fp), C.byref(nf), cases.ctypes.data if return_cases else None)
    if rc == 1:
        raise ValueError("Surface level must be within volume data range.")
    if rc == 2:
        raise RuntimeError("No surface found at the given iso value.")
    verts = np.ctypeslib.as_array(C.cast(vp, C.POINTER(C.c_float)), shape=(nv.value, 3)).copy()
    faces = np.ctypeslib.as_array(C.cast(fp, C.POINTER(C.c_int32)), shape=(nf.value, 3)).copy()
    lib.r3g_oracle_free(vp)
    lib.r3g_oracle_free(fp)
    if return_cases:
        return verts, faces, cases.reshape(n0 - 1, n1 - 1, n2 - 1)
    return verts, faces
