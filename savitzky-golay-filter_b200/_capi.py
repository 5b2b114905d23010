"""ctypes binding of libsavgol_b200.so -- the C ABI declared in include/savgol_b200.h.

The struct layouts below are the reference's public ABI (include/iterative/savgolFilter.h:92-113,
savgol_stream.h:29-37, savgol2d.h:82-103); tests/test_abi_cpu.py checks their sizes/offsets.
There is no fallback of any kind: if the shared library is missing, importing this module fails.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsavgol_b200.so")

MAX_HALF_WINDOW = 32
IPC_HANDLE_BYTES = 64   # SAVGOL_B200_IPC_HANDLE_BYTES
MAX_WINDOW = 65

f32p = C.POINTER(C.c_float)


class SavgolConfig(C.Structure):
    _fields_ = [("half_window", C.c_uint8), ("poly_order", C.c_uint8), ("derivative", C.c_uint8),
                ("time_step", C.c_float), ("boundary", C.c_int)]


class SavgolFilterStruct(C.Structure):
    _fields_ = [("config", SavgolConfig), ("window_size", C.c_int), ("dt_scale", C.c_float),
                ("center_weights", C.c_float * MAX_WINDOW),
                ("edge_weights", (C.c_float * MAX_WINDOW) * MAX_HALF_WINDOW)]


class SavgolStreamStruct(C.Structure):
    _fields_ = [("filter", C.POINTER(SavgolFilterStruct)), ("buffer", C.c_float * MAX_WINDOW),
                ("write_pos", C.c_int), ("samples_received", C.c_size_t), ("samples_output", C.c_size_t),
                ("owns_filter", C.c_bool), ("dt_inv", C.c_float)]


class Savgol2DConfig(C.Structure):
    _fields_ = [("half_window_x", C.c_uint8), ("half_window_y", C.c_uint8), ("poly_order", C.c_uint8),
                ("deriv_x", C.c_uint8), ("deriv_y", C.c_uint8), ("delta_x", C.c_float), ("delta_y", C.c_float)]


class Savgol2DFilterStruct(C.Structure):
    _fields_ = [("config", Savgol2DConfig), ("window_width", C.c_int), ("window_height", C.c_int),
                ("window_area", C.c_int), ("num_terms", C.c_int), ("scale", C.c_float), ("weights", f32p)]


FP = C.POINTER(SavgolFilterStruct)
SP = C.POINTER(SavgolStreamStruct)
F2 = C.POINTER(Savgol2DFilterStruct)

# name -> (restype, argtypes); every symbol include/savgol_b200.h declares
PROTOTYPES = {
    # part 1a
    "savgol_create": (FP, [C.POINTER(SavgolConfig)]),
    "savgol_destroy": (None, [FP]),
    "savgol_apply": (C.c_int, [FP, C.c_void_p, C.c_void_p, C.c_size_t]),
    "savgol_apply_strided": (C.c_int, [FP, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t,
                                       C.c_size_t, C.c_size_t]),
    "savgol_apply_valid": (C.c_size_t, [FP, C.c_void_p, C.c_size_t, C.c_void_p]),
    # part 1b
    "savgol_stream_create": (SP, [C.POINTER(SavgolConfig)]),
    "savgol_stream_init": (C.c_int, [SP, FP]),
    "savgol_stream_destroy": (None, [SP]),
    "savgol_stream_reset": (None, [SP]),
    "savgol_stream_push": (C.c_float, [SP, C.c_float, C.POINTER(C.c_bool)]),
    "savgol_stream_push_full": (C.c_int, [SP, C.c_float, f32p, C.c_int]),
    "savgol_stream_flush": (C.c_int, [SP, f32p, C.c_int]),
    "savgol_stream_flush_leading": (C.c_int, [SP, f32p, C.c_int]),
    "savgol_stream_ready": (C.c_bool, [SP]),
    "savgol_stream_latency": (C.c_size_t, [SP]),
    "savgol_stream_buffered": (C.c_size_t, [SP]),
    "savgol_stream_samples_received": (C.c_size_t, [SP]),
    "savgol_stream_samples_output": (C.c_size_t, [SP]),
    # part 1c
    "savgol2d_create": (F2, [C.POINTER(Savgol2DConfig)]),
    "savgol2d_destroy": (None, [F2]),
    "savgol2d_config_valid": (C.c_bool, [C.POINTER(Savgol2DConfig)]),
    "savgol2d_apply_valid": (C.c_int, [F2, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "savgol2d_apply": (C.c_int, [F2, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "savgol2d_gradient": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int]),
    "savgol2d_hessian": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int]),
    "savgol2d_laplacian": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_float, C.c_float, C.c_int]),
    # part 2
    "savgol_b200_version": (C.c_int, []),
    "savgol_b200_device_ok": (C.c_int, []),
    "savgol_b200_set_stream": (None, [C.c_void_p]),
    "savgol_b200_get_stream": (C.c_void_p, []),
    "savgol_b200_launch_count": (C.c_ulonglong, []),
    "savgol_b200_tma_launch_count": (C.c_ulonglong, []),
    "savgol_b200_plan_1d": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "savgol_b200_staging_chunk": (C.c_size_t, [C.c_size_t, C.c_int]),
    "savgol_b200_host_copy2d": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]),
    "savgol_b200_set_tma": (None, [C.c_int]),
    "savgol_b200_set_exact": (None, [C.c_int]),
    "savgol_b200_set_exact_default": (None, [C.c_int]),
    "savgol_b200_get_exact": (C.c_int, []),
    "savgol_b200_set_inplace_compat": (None, [C.c_int]),
    "savgol_b200_get_inplace_compat": (C.c_int, []),
    "savgol_apply_batch_multi": (C.c_int, [FP, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
                                           C.POINTER(C.c_int), C.c_int]),
    "savgol2d_apply_batch_multi": (C.c_int, [F2, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_int,
                                             C.POINTER(C.c_int), C.c_int]),
    "savgol_apply_slices": (C.c_int, [FP, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.c_int]),
    "savgol_b200_device_count": (C.c_int, []),
    "savgol_b200_alloc": (C.c_void_p, [C.c_int, C.c_size_t]),
    "savgol_b200_free": (None, [C.c_int, C.c_void_p]),
    "savgol_b200_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "savgol_apply_batch": (C.c_int, [FP, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]),
    "savgol_apply_halo": (C.c_int, [FP, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "savgol2d_apply_batch": (C.c_int, [F2, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int,
                                       C.c_size_t, C.c_size_t, C.c_int]),
    "savgol2d_b200_plan": (C.c_int, [F2, C.POINTER(C.c_int), C.POINTER(C.c_float)]),
    "savgol2d_b200_plan_kind": (C.c_int, [F2]),
    "savgol2d_b200_wrapper_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "savgol2d_apply_band": (C.c_int, [F2, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "savgol2d_apply_band_at": (C.c_int, [F2, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "savgol_b200_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]),
    "savgol_b200_ipc_open": (C.c_void_p, [C.c_void_p, C.c_size_t]),
    "savgol_b200_ipc_close": (C.c_int, [C.c_void_p, C.c_size_t]),
    "savgol_mcstream_create": (C.c_void_p, [C.POINTER(SavgolConfig), C.c_size_t]),
    "savgol_mcstream_destroy": (None, [C.c_void_p]),
    "savgol_mcstream_reset": (None, [C.c_void_p]),
    "savgol_mcstream_push": (C.c_longlong, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t]),
    "savgol_mcstream_flush": (C.c_longlong, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "savgol_mcstream_channels": (C.c_size_t, [C.c_void_p]),
    "savgol_mcstream_latency": (C.c_size_t, [C.c_void_p]),
    "savgol_mcstream_samples_received": (C.c_size_t, [C.c_void_p]),
    "savgol_mcstream_samples_output": (C.c_size_t, [C.c_void_p]),
    "savgol_mcstream_state": (f32p, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "savgol_mcstream_checkpoint_size": (C.c_size_t, [C.c_void_p]),
    "savgol_mcstream_save": (C.c_longlong, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "savgol_mcstream_restore": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        path = os.environ.get("SAVGOL_B200_LIB", LIB_PATH)  # override: A/B builds of the same library
        if not os.path.exists(path):
            raise ImportError(
                f"{path} not found: build it with `make -C {os.path.join(HERE, 'csrc')} -j8` "
                "(or __graft_entry__.build()); this package has no CPU fallback")
        lib = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            if "SAVGOL_B200_LIB" in os.environ and not hasattr(lib, name):
                continue             # an older A/B build (tools/ab.sh): symbols added since are simply absent
            fn = getattr(lib, name)  # AttributeError = header/library mismatch, fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
