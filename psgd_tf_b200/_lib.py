"""ctypes binding of the C ABI declared in ``include/psgd_b200.h``.

There is deliberately no fallback here: if the shared library is missing, or no CUDA device is
present when a context is requested, the caller gets an exception -- never a CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from .build import LIB_PATH

PSGD_OK = 0
PSGD_ERR_UNSUPPORTED = 3

FACTOR_DENSE, FACTOR_SCALE, FACTOR_NORM = 0, 1, 2

c_float_p = C.c_void_p  # device pointers travel as integers

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p)


class KronLayer(C.Structure):
    """Mirror of ``psgd_kron_layer``."""
    _fields_ = [
        ("kind_l", C.c_int32), ("kind_r", C.c_int32), ("M", C.c_int64), ("N", C.c_int64),
        ("Ql", C.c_void_p), ("Qr", C.c_void_p), ("dX", C.c_void_p), ("dG", C.c_void_p), ("G", C.c_void_p),
        ("Ql_out", C.c_void_p), ("Qr_out", C.c_void_p), ("out", C.c_void_p),
    ]


class ParamUpdate(C.Structure):
    """Mirror of ``psgd_param_update``."""
    _fields_ = [("W", C.c_void_p), ("pre", C.c_void_p), ("v", C.c_void_p), ("count", C.c_int64)]


class DiffItem(C.Structure):
    """Mirror of ``psgd_diff_item``."""
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("out", C.c_void_p), ("count", C.c_int64)]


# name -> (restype, argtypes); every symbol include/psgd_b200.h declares
SIGNATURES = {
    "psgd_abi_version": (C.c_int, []),
    "psgd_last_error": (C.c_char_p, []),
    "psgd_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "psgd_destroy": (C.c_int, [C.c_void_p]),
    "psgd_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "psgd_launch_count": (C.c_int64, [C.c_void_p]),
    "psgd_workspace_bytes": (C.c_int64, [C.c_void_p]),
    "psgd_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "psgd_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_int]),
    "psgd_set_allreduce": (C.c_int, [C.c_void_p, ALLREDUCE_FN, C.c_void_p]),
    "psgd_comm_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "psgd_comm_attach": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "psgd_comm_detach": (C.c_int, [C.c_void_p]),
    "psgd_comm_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "psgd_uvd_update": (C.c_int, [C.c_void_p] + [c_float_p] * 5 + [C.c_int64, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int]),
    "psgd_uvd_update_apply": (C.c_int, [C.c_void_p] + [c_float_p] * 7 + [C.c_int64, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int]),
    "psgd_uvd_plan_info": (C.c_int, [C.c_int] + [C.POINTER(C.c_int)] * 4),
    "psgd_uvd_apply": (C.c_int, [C.c_void_p] + [c_float_p] * 5 + [C.c_int64, C.c_int]),
    "psgd_uvd_step_tail": (C.c_int, [C.c_void_p] + [c_float_p] * 7 + [C.c_int64, C.c_int, C.c_float, C.c_float, C.c_float]),
    "psgd_ipuvt_matvec": (C.c_int, [C.c_void_p] + [c_float_p] * 4 + [C.c_int64, C.c_int, C.c_int]),
    "psgd_diag_update": (C.c_int, [C.c_void_p] + [c_float_p] * 3 + [C.c_int64, C.c_float, C.c_float]),
    "psgd_diag_apply": (C.c_int, [C.c_void_p] + [c_float_p] * 3 + [C.c_int64]),
    "psgd_xmat_update": (C.c_int, [C.c_void_p] + [c_float_p] * 4 + [C.c_int64, C.c_float, C.c_float]),
    "psgd_xmat_apply": (C.c_int, [C.c_void_p] + [c_float_p] * 4 + [C.c_int64]),
    "psgd_dense_update": (C.c_int, [C.c_void_p] + [c_float_p] * 4 + [C.c_int64, C.c_float, C.c_float]),
    "psgd_dense_apply": (C.c_int, [C.c_void_p] + [c_float_p] * 3 + [C.c_int64]),
    "psgd_splu_update": (C.c_int, [C.c_void_p] + [c_float_p] * 10 + [C.c_int64, C.c_int, C.c_float, C.c_float]),
    "psgd_splu_apply": (C.c_int, [C.c_void_p] + [c_float_p] * 6 + [C.c_int64, C.c_int]),
    "psgd_gemm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p, C.c_int, C.c_int, c_float_p, C.c_int,
                            C.c_int, c_float_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "psgd_kron_update": (C.c_int, [C.c_void_p, C.c_int, C.c_int] + [c_float_p] * 6 + [C.c_int64, C.c_int64, C.c_float, C.c_float]),
    "psgd_kron_apply": (C.c_int, [C.c_void_p, C.c_int, C.c_int] + [c_float_p] * 4 + [C.c_int64, C.c_int64]),
    "psgd_kron_update_batched": (C.c_int, [C.c_void_p, C.POINTER(KronLayer), C.c_int, C.c_float, C.c_float]),
    "psgd_kron_apply_batched": (C.c_int, [C.c_void_p, C.POINTER(KronLayer), C.c_int]),
    "psgd_apply_updates": (C.c_int, [C.c_void_p, C.POINTER(ParamUpdate), C.c_int, C.c_float, C.c_float]),
    "psgd_multi_sub": (C.c_int, [C.c_void_p, C.POINTER(DiffItem), C.c_int]),
}

_lib = None
_lib_lock = threading.Lock()


class PsgdError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"psgd_b200 status {status}: {message}")
        self.status = status


def load_library():
    """dlopen ``_C/libpsgd_b200.so`` (built by ``psgd_tf_b200.build.build``) and type every symbol."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        # PSGD_B200_LIB: another build of the same library (kernel-tuning A/B runs, tools/uvd_variants.py)
        path = os.environ.get("PSGD_B200_LIB") or LIB_PATH
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing: build it with `python -m psgd_tf_b200.build` "
                "(psgd_tf_b200 has no CPU or pure-Python fallback)")
        lib = C.CDLL(path)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError if the .so does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.psgd_abi_version() != 1:
            raise ImportError(f"{LIB_PATH}: ABI version {lib.psgd_abi_version()} != 1; rebuild")
        _lib = lib
        return _lib


def check(status: int):
    if status != PSGD_OK:
        msg = load_library().psgd_last_error()
        raise PsgdError(status, msg.decode() if msg else "")


class Context:
    """One ``psgd_ctx`` per (thread, device); not thread-safe, stream-ordered."""

    def __init__(self, device: int, stream: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        check(self.lib.psgd_create(int(device), C.c_void_p(stream), C.byref(h)))
        self.handle = h
        self.device = int(device)
        self._hook = None

    def close(self):
        if self.handle:
            self.lib.psgd_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream: int):
        check(self.lib.psgd_set_stream(self.handle, C.c_void_p(stream)))

    def set_option(self, key: str, value: int):
        check(self.lib.psgd_set_option(self.handle, key.encode(), int(value)))

    @property
    def launch_count(self) -> int:
        return int(self.lib.psgd_launch_count(self.handle))

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.psgd_workspace_bytes(self.handle))

    def profile_read(self, cap: int = 4096):
        """[(kernel id, milliseconds, work)] recorded since the last read (needs set_option("profile", 1))."""
        ids = (C.c_int * cap)()
        ms = (C.c_float * cap)()
        work = (C.c_double * cap)()
        n = self.lib.psgd_profile_read(self.handle, ids, ms, work, cap)
        return [(int(ids[i]), float(ms[i]), float(work[i])) for i in range(n)]

    # ---- peer-memory exchange (include/psgd_b200.h, csrc/comm.cu) -----------------------------------
    COMM_HANDLE_BYTES = 64

    def comm_export(self) -> bytes:
        """Allocate this rank's exchange slab and return its CUDA IPC handle (64 bytes)."""
        buf = C.create_string_buffer(self.COMM_HANDLE_BYTES)
        check(self.lib.psgd_comm_export(self.handle, buf))
        return buf.raw

    def comm_attach(self, rank: int, world: int, handles):
        """``handles``: every rank's ``comm_export()`` bytes, in rank order."""
        blob = b"".join(bytes(h) for h in handles)
        if len(blob) != world * self.COMM_HANDLE_BYTES:
            raise ValueError(f"comm_attach: expected {world} handles of {self.COMM_HANDLE_BYTES} bytes")
        check(self.lib.psgd_comm_attach(self.handle, int(rank), int(world), blob))

    def comm_detach(self):
        check(self.lib.psgd_comm_detach(self.handle))

    def comm_status(self) -> int:
        """Synchronise and return the number of exchanges completed; raises if a peer wait timed out."""
        e = C.c_int64(0)
        check(self.lib.psgd_comm_status(self.handle, C.byref(e)))
        return int(e.value)

    def set_allreduce(self, pyfunc):
        """pyfunc(device_ptr:int, count:int, op:int, stream:int) -> int, or None to clear."""
        if pyfunc is None:
            self._hook = None
            check(self.lib.psgd_set_allreduce(self.handle, C.cast(None, ALLREDUCE_FN), None))
            return

        def tramp(user, buf, count, op, stream):
            try:
                return int(pyfunc(int(buf or 0), int(count), int(op), int(stream or 0)) or 0)
            except Exception:  # never unwind through C
                import traceback
                traceback.print_exc()
                return 1

        self._hook = ALLREDUCE_FN(tramp)
        check(self.lib.psgd_set_allreduce(self.handle, self._hook, None))
