"""
TEST INFRASTRUCTURE ONLY -- ctypes front end of ``oracle/piranha_cpu.c`` (the C restatement of piranha's CPU
multiplier).  Used by tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference``.
Never imported by ``piranha_b200``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

from . import pyoracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpiranha_cpu_oracle.so")

PC_OK, PC_E_ARGS, PC_E_KEY_RANGE, PC_E_NOMEM = 0, 1, 2, 4


class _PcResult(C.Structure):
    _fields_ = [
        ("n", C.c_uint64),
        ("key", C.POINTER(C.c_int64)),
        ("mag", C.POINTER(C.c_uint64)),
        ("sign", C.POINTER(C.c_int8)),
        ("limbs", C.c_uint32),
        ("seconds", C.c_double),
        ("buckets", C.c_uint64),
        ("threads_used", C.c_uint32),
    ]


def build(force: bool = False) -> str:
    newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("piranha_cpu.c", "pc_dispatch.c", "Makefile"))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.pc_mul.restype = C.c_int
        _lib.pc_mul.argtypes = [
            C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
            C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
            C.c_uint32, C.c_void_p, C.c_int64, C.c_uint32, C.c_uint64,
            C.c_int, C.c_int64, C.c_void_p, C.POINTER(_PcResult),
        ]
        _lib.pc_result_free.argtypes = [C.POINTER(_PcResult)]
    return _lib


class CpuResult:
    def __init__(self, key, mag, sign, seconds, buckets, threads_used):
        self.key, self.mag, self.sign = key, mag, sign
        self.seconds, self.buckets, self.threads_used = seconds, buckets, threads_used

    def canonical(self):
        """Sorted by Kronecker code, magnitude trimmed to the limbs actually needed."""
        order = np.argsort(self.key, kind="stable")
        key, mag, sign = self.key[order], self.mag[order], self.sign[order]
        L = mag.shape[1]
        while L > 1 and not mag[:, L - 1].any():
            L -= 1
        return key, np.ascontiguousarray(mag[:, :L]), sign


def cpu_mul(a, b, n_vars: int, n_threads: int = 1, min_work: int = 0, trunc_mode: int = 0, max_degree: int = 0,
            deg_mask: Optional[Sequence[int]] = None) -> CpuResult:
    """a, b: (key int64[n], mag uint64[n, limbs], sign int8[n]).  Raises OverflowError like the reference's
    std::overflow_error when the exponent bounds check fails."""
    ka, ma, sa = (np.ascontiguousarray(x) for x in a)
    kb, mb, sb = (np.ascontiguousarray(x) for x in b)
    ma = ma.reshape(len(ka), -1) if len(ka) else ma.reshape(0, max(1, ma.shape[-1] if ma.ndim > 1 else 1))
    mb = mb.reshape(len(kb), -1) if len(kb) else mb.reshape(0, max(1, mb.shape[-1] if mb.ndim > 1 else 1))
    M, hmin, _, _ = pyoracle.kron_limits(n_vars) if n_vars else ((), 0, 0, 0)
    lim = np.asarray(M, dtype=np.int64)
    mask = np.asarray(deg_mask if deg_mask is not None else [1] * n_vars, dtype=np.uint8)
    res = _PcResult()
    rc = lib().pc_mul(
        len(ka), ka.ctypes.data, ma.ctypes.data, ma.shape[1], sa.ctypes.data,
        len(kb), kb.ctypes.data, mb.ctypes.data, mb.shape[1], sb.ctypes.data,
        n_vars, lim.ctypes.data, hmin, n_threads, min_work, trunc_mode, max_degree, mask.ctypes.data, C.byref(res))
    if rc == PC_E_KEY_RANGE:
        raise OverflowError("Kronecker monomial components are out of bounds")
    if rc == PC_E_ARGS:
        raise ValueError("invalid arguments")
    if rc != PC_OK:
        raise MemoryError(f"pc_mul failed: {rc}")
    n = res.n
    if n:
        key = np.ctypeslib.as_array(res.key, shape=(n,)).copy()
        mag = np.ctypeslib.as_array(res.mag, shape=(n, res.limbs)).copy()
        sign = np.ctypeslib.as_array(res.sign, shape=(n,)).copy()
    else:
        key = np.zeros(0, np.int64)
        mag = np.zeros((0, 1), np.uint64)
        sign = np.zeros(0, np.int8)
    out = CpuResult(key, mag, sign, res.seconds, res.buckets, res.threads_used)
    lib().pc_result_free(C.byref(res))
    return out
