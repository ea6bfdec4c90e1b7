"""ctypes binding of libpgica.so (the C ABI declared in include/pgica.h).

The library is the product: if it is missing or fails to load, every op raises — there is no
PyTorch/CPU fallback behind these calls.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpgica.so")

_lock = threading.Lock()
_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_int64 = ctypes.c_int64
c_float = ctypes.c_float
c_size_t = ctypes.c_size_t
c_uint32 = ctypes.c_uint32

# name -> (restype, argtypes); must list every symbol include/pgica.h declares (tests check this).
SIGNATURES = {
    "pgica_abi_version": (c_int, []),
    "pgica_last_error": (ctypes.c_char_p, []),
    "pgica_device_check": (c_int, []),
    "pgica_sm_count": (c_int, []),
    "pgica_gemm_lse_workspace_bytes": (c_int, [c_int64, c_int64, c_int64, ctypes.POINTER(c_size_t)]),
    "pgica_gemm_lse": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float, c_void_p, c_int64,
                               c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "pgica_probe_umma": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_uint32, c_uint32,
                                 c_void_p, c_void_p]),
}


class PgicaError(RuntimeError):
    pass


def load(build_if_missing=True):
    """Load (building first if the in-tree .so is absent and nvcc is available)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            if not build_if_missing:
                raise PgicaError(f"{LIB_PATH} is missing; run `python __graft_entry__.py build`")
            from . import _build
            _build.build()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch: fail loudly
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.pgica_abi_version() != 1:
            raise PgicaError("libpgica.so ABI version mismatch")
        _lib = lib
        return lib


def check(rc, lib=None):
    if rc != 0:
        lib = lib or load()
        msg = lib.pgica_last_error()
        raise PgicaError(f"pgica error {rc}: {msg.decode() if msg else '?'}")
