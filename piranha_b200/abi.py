"""ctypes binding of the C ABI declared in ``include/piranha_b200.h``.

This is plumbing for the tests and the benchmark harness: every call goes straight to the CUDA library
(``piranha_b200/lib/libpiranha_b200.so``).  There is no Python or CPU implementation behind these functions;
if the library is missing or no B200 is present the calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpiranha_b200.so")

PK_OK, PK_E_ARGS, PK_E_KEY_RANGE, PK_E_CF_WIDTH, PK_E_NOMEM, PK_E_CUDA, PK_E_DEGREE, PK_E_UNSUPPORTED = range(8)
PK_ALGO_AUTO, PK_ALGO_DENSE, PK_ALGO_HASH, PK_ALGO_ZONE, PK_ALGO_BLOCK = range(5)
ALGO_NAMES = {0: "auto", 1: "dense", 2: "hash", 3: "zone", 4: "block"}

EXPORTED_SYMBOLS = [
    "pk_ctx_create", "pk_ctx_create_multi", "pk_ctx_device_count", "pk_device_count", "pk_ctx_synchronize", "pk_ctx_destroy", "pk_ctx_trim", "pk_last_error", "pk_ctx_set_limits", "pk_ctx_get_limits", "pk_get_stats", "pk_ctx_set_tuning",
    "pk_mul", "pk_result_free", "pk_upload", "pk_mul_dev", "pk_check_bounds", "pk_download", "pk_download_grouped", "pk_download_grouped_progress", "pk_dpoly_free", "pk_dpoly_size",
    "pk_dpoly_limbs", "pk_dpoly_nvars", "pk_dpoly_export", "pk_dpoly_export_padded", "pk_dpoly_digest",
    "pk_dpoly_publish_size", "pk_dpoly_export_peers",
]


class PkPoly(C.Structure):
    _fields_ = [("n", C.c_uint64), ("n_vars", C.c_uint32), ("limbs", C.c_uint32), ("key", C.c_void_p),
                ("mag", C.c_void_p), ("sign", C.c_void_p)]


class PkResult(C.Structure):
    _fields_ = [("n", C.c_uint64), ("n_vars", C.c_uint32), ("limbs", C.c_uint32), ("key", C.POINTER(C.c_int64)),
                ("mag", C.POINTER(C.c_uint64)), ("sign", C.POINTER(C.c_int8)), ("owner", C.c_void_p)]


PROGRESS_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(PkResult), C.c_uint64)  # pk_progress_fn


class PkOpts(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("trunc_mode", C.c_int32), ("max_degree", C.c_int64),
                ("degree_mask", C.c_void_p), ("part_index", C.c_uint32), ("part_count", C.c_uint32),
                ("algo", C.c_uint32), ("reserved", C.c_uint32)]


class PkStats(C.Structure):
    _fields_ = [("n1", C.c_uint64), ("n2", C.c_uint64), ("term_products", C.c_uint64), ("result_terms", C.c_uint64),
                ("code_range", C.c_uint64), ("table_slots", C.c_uint64), ("algo", C.c_uint32), ("acc_lanes", C.c_uint32),
                ("chunk_bits", C.c_uint32), ("result_limbs", C.c_uint32), ("kernel_launches", C.c_uint32),
                ("retries", C.c_uint32), ("total_launches", C.c_uint64), ("mul_kernel_ms", C.c_float),
                ("device_ms", C.c_float)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class PkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"pk error {code}: {msg}")
        self.code = code


def _raise(code: int, msg: str):
    """Map status codes to the exception types the reference throws."""
    if code == PK_E_ARGS:
        raise ValueError(msg)  # std::invalid_argument
    if code in (PK_E_KEY_RANGE, PK_E_DEGREE, PK_E_CF_WIDTH):
        raise OverflowError(msg)  # std::overflow_error (the C++ mirror maps PK_E_CF_WIDTH the same way)
    if code == PK_E_NOMEM:
        raise MemoryError(msg)  # std::bad_alloc
    raise PkError(code, msg)


_lib = None


def load_library():
    """Load the CUDA library.  Raises if it has not been built: the product has no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with ./build.sh (or __graft_entry__.build()); "
                           "piranha_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, u32, u64, i64 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int64
    lib.pk_ctx_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
    lib.pk_ctx_create_multi.argtypes = [u32, C.POINTER(C.c_int), C.POINTER(vp)]
    lib.pk_device_count.argtypes = []
    lib.pk_ctx_synchronize.argtypes = [vp]
    lib.pk_ctx_device_count.argtypes = [vp]
    lib.pk_ctx_device_count.restype = u32
    lib.pk_ctx_destroy.argtypes = [vp]
    lib.pk_ctx_destroy.restype = None
    lib.pk_ctx_trim.argtypes = [vp]
    lib.pk_last_error.argtypes = [vp]
    lib.pk_last_error.restype = C.c_char_p
    lib.pk_ctx_set_limits.argtypes = [vp, u32, vp]
    lib.pk_ctx_get_limits.argtypes = [vp, u32, vp, C.POINTER(i64), C.POINTER(i64)]
    lib.pk_get_stats.argtypes = [vp, C.POINTER(PkStats)]
    lib.pk_ctx_set_tuning.argtypes = [vp, C.c_char_p, u64]
    lib.pk_mul.argtypes = [vp, C.POINTER(PkPoly), C.POINTER(PkPoly), C.POINTER(PkOpts), C.POINTER(PkResult)]
    lib.pk_result_free.argtypes = [vp, C.POINTER(PkResult)]
    lib.pk_result_free.restype = None
    lib.pk_upload.argtypes = [vp, C.POINTER(PkPoly), C.POINTER(vp)]
    lib.pk_mul_dev.argtypes = [vp, vp, vp, C.POINTER(PkOpts), C.POINTER(vp)]
    lib.pk_check_bounds.argtypes = [vp, vp, vp]
    lib.pk_download.argtypes = [vp, vp, C.POINTER(PkResult)]
    lib.pk_download_grouped.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.POINTER(PkResult)]
    lib.pk_download_grouped_progress.argtypes = [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, PROGRESS_FN, vp, vp, C.POINTER(PkResult)]
    lib.pk_dpoly_free.argtypes = [vp, vp]
    lib.pk_dpoly_free.restype = None
    lib.pk_dpoly_size.argtypes = [vp]
    lib.pk_dpoly_size.restype = u64
    lib.pk_dpoly_limbs.argtypes = [vp]
    lib.pk_dpoly_limbs.restype = u32
    lib.pk_dpoly_nvars.argtypes = [vp]
    lib.pk_dpoly_nvars.restype = u32
    lib.pk_dpoly_export.argtypes = [vp, vp, vp, vp, vp]
    lib.pk_dpoly_export_padded.argtypes = [vp, vp, vp, vp, vp, u32]
    lib.pk_dpoly_publish_size.argtypes = [vp, vp, u32, vp]
    lib.pk_dpoly_export_peers.argtypes = [vp, vp, u32, u32, vp, vp, vp, vp, u64, u32]
    lib.pk_dpoly_digest.argtypes = [vp, vp, C.POINTER(u64)]
    _lib = lib
    return lib


Terms = Tuple[np.ndarray, np.ndarray, np.ndarray]  # key int64[n], mag uint64[n, limbs], sign int8[n]


def _as_terms(t: Terms) -> Terms:
    key = np.ascontiguousarray(t[0], dtype=np.int64)
    mag = np.ascontiguousarray(t[1], dtype=np.uint64)
    sign = np.ascontiguousarray(t[2], dtype=np.int8)
    if mag.ndim == 1:
        mag = mag.reshape(len(key), -1) if len(key) else mag.reshape(0, 1)
    assert mag.shape[0] == len(key) == len(sign)
    return key, mag, sign


def make_opts(trunc_mode: int = 0, max_degree: int = 0, degree_mask: Optional[Sequence[int]] = None, part_index: int = 0,
              part_count: int = 0, algo: int = PK_ALGO_AUTO):
    o = PkOpts()
    o.struct_size = C.sizeof(PkOpts)
    o.trunc_mode = trunc_mode
    o.max_degree = max_degree
    keep = None
    if degree_mask is not None:
        keep = np.ascontiguousarray(degree_mask, dtype=np.uint8)
        o.degree_mask = keep.ctypes.data
    o.part_index, o.part_count, o.algo = part_index, part_count, algo
    return o, keep


class DevicePoly:
    """A polynomial resident in HBM (handle to pk_dpoly)."""

    def __init__(self, ctx: "Context", handle: int):
        self.ctx, self.handle = ctx, handle

    def __len__(self):
        return int(self.ctx.lib.pk_dpoly_size(self.handle))

    @property
    def limbs(self) -> int:
        return int(self.ctx.lib.pk_dpoly_limbs(self.handle))

    @property
    def n_vars(self) -> int:
        return int(self.ctx.lib.pk_dpoly_nvars(self.handle))

    def download(self) -> Terms:
        return self.ctx.download(self)

    def digest(self) -> int:
        d = C.c_uint64()
        self.ctx._check(self.ctx.lib.pk_dpoly_digest(self.ctx.handle, self.handle, C.byref(d)))
        return int(d.value)

    def export_to(self, key_ptr: int, mag_ptr: int, sign_ptr: int, mag_limbs: int = 0):
        """Copy the terms into caller-owned DEVICE buffers (e.g. torch tensors' data_ptr()); mag rows of `mag_limbs`
        words (default: this polynomial's own limb count)."""
        self.ctx._check(self.ctx.lib.pk_dpoly_export_padded(self.ctx.handle, self.handle, key_ptr, mag_ptr, sign_ptr,
                                                            mag_limbs or self.limbs))

    def publish_size(self, slot_ptrs: Sequence[int]):
        """Write {term count, limbs} into slot_ptrs[j] (peer-mapped device address, 2 x uint64) on every peer j."""
        tab = (C.c_uint64 * len(slot_ptrs))(*slot_ptrs)
        self.ctx._check(self.ctx.lib.pk_dpoly_publish_size(self.ctx.handle, self.handle, len(slot_ptrs), tab))

    def export_peers(self, rank: int, sizes_ptr: int, key_ptrs: Sequence[int], mag_ptrs: Sequence[int],
                     sign_ptrs: Sequence[int], cap_terms: int, cap_limbs: int):
        """Push the terms into the result arrays of every peer (peer-mapped device addresses, index = rank) at the term
        offset the kernel derives from the size table at sizes_ptr (pk_dpoly_export_peers)."""
        n = len(key_ptrs)
        k, m, g = (C.c_uint64 * n)(*key_ptrs), (C.c_uint64 * n)(*mag_ptrs), (C.c_uint64 * n)(*sign_ptrs)
        self.ctx._check(self.ctx.lib.pk_dpoly_export_peers(self.ctx.handle, self.handle, n, rank, sizes_ptr, k, m, g,
                                                           cap_terms, cap_limbs))

    def free(self):
        if self.handle:
            self.ctx.lib.pk_dpoly_free(self.ctx.handle, self.handle)
            self.handle = None

    def __del__(self):
        try:
            if self.handle and self.ctx.handle:
                self.free()
        except Exception:
            pass


class Context:
    """pk_ctx: one device and one stream, or (Context.multi) several devices behind one handle."""

    def __init__(self, device: int = 0, stream: int = 0, devices: Optional[Sequence[int]] = None):
        self.lib = load_library()
        h = C.c_void_p()
        if devices is not None:
            ids = (C.c_int * len(devices))(*devices)
            rc = self.lib.pk_ctx_create_multi(len(devices), ids, C.byref(h))
        else:
            rc = self.lib.pk_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != PK_OK:
            raise PkError(rc, "pk_ctx_create failed: no usable sm_100 GPU (there is no CPU fallback)")
        self.handle = h
        self.stream = int(stream) if (stream and devices is None) else None  # None: streams private to the context

    def synchronize(self):
        """Wait for everything queued on the context's stream(s)."""
        self._check(self.lib.pk_ctx_synchronize(self.handle))

    @classmethod
    def multi(cls, devices: Sequence[int]) -> "Context":
        """pk_ctx_create_multi: every product is cut into len(devices) output-code ranges, one per device."""
        return cls(devices=list(devices))

    @property
    def device_count(self) -> int:
        return int(self.lib.pk_ctx_device_count(self.handle))

    def close(self):
        if self.handle:
            self.lib.pk_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trim(self):
        """Release cached workspaces (staging arrays, idle pinned buffers, pooled device memory)."""
        self._check(self.lib.pk_ctx_trim(self.handle))

    def _check(self, rc: int):
        if rc != PK_OK:
            _raise(rc, self.lib.pk_last_error(self.handle).decode())

    def limits(self, n_vars: int):
        lim = np.zeros(max(n_vars, 1), dtype=np.int64)
        hmin, hmax = C.c_int64(), C.c_int64()
        self._check(self.lib.pk_ctx_get_limits(self.handle, n_vars, lim.ctypes.data, C.byref(hmin), C.byref(hmax)))
        return lim[:n_vars].copy(), int(hmin.value), int(hmax.value)

    def set_limits(self, n_vars: int, limits: Sequence[int]):
        lim = np.ascontiguousarray(limits, dtype=np.int64)
        self._check(self.lib.pk_ctx_set_limits(self.handle, n_vars, lim.ctypes.data))

    def set_tuning(self, **kv):
        for k, v in kv.items():
            self._check(self.lib.pk_ctx_set_tuning(self.handle, k.encode(), int(v)))

    def stats(self) -> dict:
        s = PkStats()
        self._check(self.lib.pk_get_stats(self.handle, C.byref(s)))
        return s.as_dict()

    @staticmethod
    def _poly_struct(t: Terms, n_vars: int):
        key, mag, sign = _as_terms(t)
        p = PkPoly()
        p.n, p.n_vars, p.limbs = len(key), n_vars, max(1, mag.shape[1])
        p.key, p.mag, p.sign = key.ctypes.data, mag.ctypes.data, sign.ctypes.data
        return p, (key, mag, sign)

    @staticmethod
    def _take_result(lib, ctx_handle, r: PkResult) -> Terms:
        n, L = int(r.n), max(1, int(r.limbs))
        if n:
            key = np.ctypeslib.as_array(r.key, shape=(n,)).copy()
            mag = np.ctypeslib.as_array(r.mag, shape=(n, L)).copy()
            sign = np.ctypeslib.as_array(r.sign, shape=(n,)).copy()
        else:
            key, mag, sign = np.zeros(0, np.int64), np.zeros((0, L), np.uint64), np.zeros(0, np.int8)
        lib.pk_result_free(ctx_handle, C.byref(r))
        return key, mag, sign

    # ---- host-to-host product -------------------------------------------------------------------------------
    def mul(self, a: Terms, b: Terms, n_vars: int, **opt) -> Terms:
        pa, keep_a = self._poly_struct(a, n_vars)
        pb, keep_b = self._poly_struct(b, n_vars)
        o, keep_o = make_opts(**opt)
        r = PkResult()
        self._check(self.lib.pk_mul(self.handle, C.byref(pa), C.byref(pb), C.byref(o), C.byref(r)))
        del keep_a, keep_b, keep_o
        return self._take_result(self.lib, self.handle, r)

    def mul_raw(self, pa: PkPoly, pb: PkPoly, o: PkOpts, r: PkResult):
        """pk_mul on pre-built structs, result left in library-owned pinned memory (benchmark path)."""
        self._check(self.lib.pk_mul(self.handle, C.byref(pa), C.byref(pb), C.byref(o), C.byref(r)))

    def result_free(self, r: PkResult):
        self.lib.pk_result_free(self.handle, C.byref(r))

    # ---- device-resident ------------------------------------------------------------------------------------
    def upload(self, t: Terms, n_vars: int) -> DevicePoly:
        p, keep = self._poly_struct(t, n_vars)
        h = C.c_void_p()
        self._check(self.lib.pk_upload(self.handle, C.byref(p), C.byref(h)))
        del keep
        return DevicePoly(self, h)

    def mul_dev(self, a: DevicePoly, b: DevicePoly, **opt) -> DevicePoly:
        o, keep_o = make_opts(**opt)
        h = C.c_void_p()
        self._check(self.lib.pk_mul_dev(self.handle, a.handle, b.handle, C.byref(o), C.byref(h)))
        del keep_o
        return DevicePoly(self, h)

    def download(self, p: DevicePoly) -> Terms:
        r = PkResult()
        self._check(self.lib.pk_download(self.handle, p.handle, C.byref(r)))
        return self._take_result(self.lib, self.handle, r)

    def download_grouped(self, p: DevicePoly, shift: int, bits: int, n_chunks: int = 0, progress=None, offsets=None) -> Terms:
        """The terms ordered by ((uint64)key >> shift) & (2**bits - 1), ascending keys inside a group (pk_download_grouped).
        n_chunks > 0: the copy runs in pieces and progress(terms_ready) is called after each (pk_download_grouped_progress);
        offsets: a uint64 numpy array of 2**bits + 1 entries that receives the first term of every group."""
        r = PkResult()
        if n_chunks:
            fn = PROGRESS_FN((lambda user, out, ready: progress(int(ready))) if progress else (lambda user, out, ready: None))
            if offsets is not None:
                assert offsets.dtype == np.uint64 and offsets.flags["C_CONTIGUOUS"] and len(offsets) == (1 << bits) + 1
            optr = offsets.ctypes.data_as(C.c_void_p) if offsets is not None else None
            self._check(self.lib.pk_download_grouped_progress(self.handle, p.handle, shift, bits, n_chunks, fn, None, optr, C.byref(r)))
        else:
            self._check(self.lib.pk_download_grouped(self.handle, p.handle, shift, bits, C.byref(r)))
        return self._take_result(self.lib, self.handle, r)


def digest_terms(key: np.ndarray, mag: np.ndarray, sign: np.ndarray) -> int:
    """Host (numpy) evaluation of the order-independent digest that pk_dpoly_digest computes on the device."""
    key = np.asarray(key, dtype=np.int64)
    mag = np.asarray(mag, dtype=np.uint64).reshape(len(key), -1)
    sign = np.asarray(sign, dtype=np.int8)

    def mix(x):
        x = x ^ (x >> np.uint64(33))
        x = x * np.uint64(0xFF51AFD7ED558CCD)
        x = x ^ (x >> np.uint64(33))
        x = x * np.uint64(0xC4CEB9FE1A85EC53)
        x = x ^ (x >> np.uint64(33))
        return x

    with np.errstate(over="ignore"):
        h = mix(key.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        h = mix(h ^ sign.astype(np.uint8).astype(np.uint64))
        for l in range(mag.shape[1]):
            m = mag[:, l]
            hh = mix(h ^ (m + np.uint64((0x632BE59BD9B4E019 * (l + 1)) & 0xFFFFFFFFFFFFFFFF)))
            h = np.where(m != 0, hh, h)
        return int(h.sum(dtype=np.uint64))
