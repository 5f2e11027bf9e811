"""
ctypes binding of libpgx.so (include/pgx.h).  This is the whole FFI layer: plain C
pointers and sizes, no torch types.  There is no CPU fallback: if the shared object
is missing or there is no CUDA device, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGX_LIB") or os.path.join(_HERE, "csrc", "libpgx.so")  # PGX_LIB: A/B builds

PGX_OK, PGX_ERR_INVALID, PGX_ERR_CUDA, PGX_ERR_NO_DEVICE, PGX_ERR_NOMEM = 0, -1, -2, -3, -4
PGX_FLAG_MIXDOWN_INPUT = 1
PGX_PULL_MIX, PGX_PULL_INPUT_RESIDENT, PGX_PULL_X_DEVICE, PGX_PULL_X_PCM16, PGX_PULL_Y_PCM16 = 1, 2, 4, 8, 16
PGX_PULL_REDUCE = 32
PGX_OSC_SNAPSHOT = 64
PGX_CTL_HOST = 1
PGX_CTL_AMP_OSC = 2
PGX_COMM_HANDLE_BYTES = 128
PGX_OSC_SINE, PGX_OSC_BLIT = 0, 1
ABI_VERSION = 2


class Layout(C.Structure):
    _fields_ = [("stream", C.c_int64), ("chan", C.c_int64), ("samp", C.c_int64)]


class BankConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("device", "n_streams", "c_in", "c_out", "filter_len",
                                          "filter_channels", "n_filters", "block", "max_pull")] + \
               [("flags", C.c_uint32), ("tail_block", C.c_int32)]


class BankInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_streams", "c_in", "c_x", "c_out", "filter_len", "filter_channels",
                                          "n_filters", "block", "partitions", "max_pull", "device", "head",
                                          "fill")] + \
               [(n, C.c_int64) for n in ("state_bytes", "kernel_launches", "block_steps")] + \
               [(n, C.c_int32) for n in ("mac_grid", "mac_split", "mac_stream_tile", "mac_occupancy",
                                          "tail_block", "tail_partitions", "mac_tile", "submit_depth")] + \
               [("graph_pulls", C.c_int64)]


class OscConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("device", "kind", "n_voices", "unison", "channels", "sample_rate",
                                          "max_pull", "reserved")] + [("leak", C.c_double)]


class Profile(C.Structure):
    _fields_ = [("ms_r2c", C.c_double), ("ms_mac", C.c_double), ("ms_c2r", C.c_double), ("steps", C.c_int64),
                ("ms_fold", C.c_double), ("ms_now", C.c_double), ("n_mac", C.c_int64),
                ("ms_mac_union", C.c_double), ("ms_conv1", C.c_double)]


_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)

# name -> (restype, argtypes): every symbol include/pgx.h declares
PROTOTYPES = {
    "pgx_abi_version": (C.c_int, []),
    "pgx_struct_size": (C.c_int, [C.c_int32]),
    "pgx_last_error": (C.c_char_p, []),
    "pgx_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pgx_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "pgx_host_alloc_flags": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64, C.c_int32]),
    "pgx_host_free": (C.c_int, [C.c_void_p]),
    "pgx_device_alloc": (C.c_int, [C.c_int32, C.c_int64, C.POINTER(C.c_void_p)]),
    "pgx_device_free": (C.c_int, [C.c_int32, C.c_void_p]),
    "pgx_device_upload": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]),
    "pgx_device_zero": (C.c_int, [C.c_int32, C.c_void_p, C.c_int64]),
    "pgx_bank_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(BankConfig), _f32p, _i32p]),
    "pgx_bank_destroy": (C.c_int, [C.c_void_p]),
    "pgx_bank_get_info": (C.c_int, [C.c_void_p, C.POINTER(BankInfo)]),
    "pgx_bank_reset": (C.c_int, [C.c_void_p, _i32p, C.c_int32]),
    "pgx_bank_load_filter": (C.c_int, [C.c_void_p, C.c_int32, _f32p]),
    "pgx_bank_set_filter_map": (C.c_int, [C.c_void_p, _i32p]),
    "pgx_bank_set_output_gains": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "pgx_bank_use_filter_map_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pgx_bank_process": (C.c_int, [C.c_void_p, C.c_void_p, Layout, C.c_void_p, Layout, C.c_int32]),
    "pgx_bank_process_mix": (C.c_int, [C.c_void_p, C.c_void_p, Layout, C.c_void_p, Layout, C.c_int32]),
    "pgx_bank_submit": (C.c_int, [C.c_void_p, C.c_void_p, Layout, C.c_void_p, Layout, C.c_int32, C.c_int32,
                                  C.POINTER(C.c_int64)]),
    "pgx_bank_wait": (C.c_int, [C.c_void_p, C.c_int64]),
    "pgx_bank_pull": (C.c_int, [C.c_void_p, C.c_void_p, Layout, C.c_void_p, Layout, C.c_int32, C.c_int32]),
    "pgx_bank_stream": (C.c_void_p, [C.c_void_p]),
    "pgx_osc_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(OscConfig), C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "pgx_osc_destroy": (C.c_int, [C.c_void_p]),
    "pgx_osc_reset": (C.c_int, [C.c_void_p]),
    "pgx_osc_render_device": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                        C.POINTER(C.c_void_p)]),
    "pgx_osc_render": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "pgx_osc_launches": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "pgx_osc_rollback": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pgx_osc_render_modulated": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int32, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]),
    "pgx_bank_process_device": (C.c_int, [C.c_void_p, C.c_void_p, Layout, C.c_void_p, Layout, C.c_int32,
                                          C.c_int32, C.c_void_p]),
    "pgx_bank_synchronize": (C.c_int, [C.c_void_p]),
    "pgx_bank_profile_begin": (C.c_int, [C.c_void_p]),
    "pgx_bank_profile_end": (C.c_int, [C.c_void_p, C.POINTER(Profile)]),
    "pgx_comm_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_void_p]),
    "pgx_comm_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pgx_comm_check": (C.c_int, [C.c_void_p]),
    "pgx_comm_destroy": (C.c_int, [C.c_void_p]),
    "pgx_mix_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "pgx_bank_attach_comm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pgx_nearest_direction": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "pgx_mix_sum": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]),
    "pgx_mix_sum_device": (C.c_int, [C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
}

_lib = None


def lib():
    """Load libpgx.so once.  Loud failure when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA library is not built. Run `python -m pygmu2_b200.build` "
                "(or __graft_entry__.build()). pygmu2_b200 has no CPU fallback."
            )
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        if h.pgx_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libpgx ABI {h.pgx_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
        for which, st in enumerate((Layout, BankConfig, BankInfo, Profile, OscConfig)):
            if h.pgx_struct_size(which) != C.sizeof(st):
                raise RuntimeError(f"libpgx struct {st.__name__}: library {h.pgx_struct_size(which)} bytes, "
                                   f"binding {C.sizeof(st)} bytes")
        _lib = h
    return _lib


def check(rc: int) -> None:
    """Map pgx_status onto the reference's exception conventions."""
    if rc == PGX_OK:
        return
    msg = lib().pgx_last_error().decode("utf-8", "replace")
    if rc == PGX_ERR_INVALID:
        raise ValueError(msg)
    if rc == PGX_ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().pgx_device_count(C.byref(n))
    return int(n.value) if rc == PGX_OK else 0


def require_device() -> None:
    if device_count() < 1:
        raise RuntimeError("pygmu2_b200: no CUDA device visible; this package has no CPU fallback")


def f32_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class PinnedArray:
    """A numpy view (float32, or int16 for PCM staging) over cudaHostAlloc'ed memory (pinned, for async H2D/D2H)."""

    def __init__(self, shape, dtype=np.float32, write_combined: bool = False):
        self.shape = tuple(int(s) for s in shape)
        dt = np.dtype(dtype)
        n = int(np.prod(self.shape)) if self.shape else 1
        self._ptr = C.c_void_p()
        check(lib().pgx_host_alloc_flags(C.byref(self._ptr), max(n, 1) * dt.itemsize, 1 if write_combined else 0))
        buf = (C.c_char * (max(n, 1) * dt.itemsize)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dt, count=n).reshape(self.shape)

    def free(self) -> None:
        if self._ptr is not None and self._ptr.value:
            self.array = None
            lib().pgx_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
