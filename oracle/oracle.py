"""ctypes front-end for the CPU checker libraries (TEST INFRASTRUCTURE ONLY).

Two libraries live under oracle/:

* ``liboracle.so``  -- oracle/savgol_oracle.c, this repo's own plain-C restatement
  of the reference arithmetic ("port"), plus the pthread row driver
  (oracle/cpu_harness.c).
* ``_ref/libsavgol_ref.so`` -- the unmodified reference compiled from
  /root/reference by oracle/Makefile ("reference").  Present whenever it was
  built in the authoring container; it travels to the GPU box with the snapshot.

Nothing in the product package imports this module.  Allowed importers:
tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MAX_WS = 65
MAX_N = 32

f32p = C.POINTER(C.c_float)


def _fp(a: np.ndarray):
    assert a.dtype == np.float32
    return a.ctypes.data_as(f32p)


def build(quiet: bool = True) -> None:
    """Compile liboracle.so (and _ref when /root/reference exists)."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.sgo_config_ok.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float]
        L.sgo_weights_1d.argtypes = [C.c_int, C.c_int, C.c_int, f32p, f32p]
        L.sgo_dt_inv.argtypes = [C.c_float, C.c_int]
        L.sgo_dt_inv.restype = C.c_float
        L.sgo_apply.argtypes = [C.c_int, f32p, f32p, C.c_float, C.c_int, f32p, f32p, C.c_size_t]
        L.sgo_apply_valid.argtypes = [C.c_int, f32p, C.c_float, f32p, C.c_size_t, f32p]
        L.sgo_apply_valid.restype = C.c_size_t
        L.sgo_apply_strided.argtypes = [C.c_int, f32p, f32p, C.c_float, C.c_void_p, C.c_size_t,
                                        C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
        L.sgo_apply_batch.argtypes = [C.c_int, f32p, f32p, C.c_float, C.c_int, f32p, f32p,
                                      C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]
        L.sgo_stream_run.argtypes = [C.c_int, f32p, f32p, C.c_float, f32p, C.c_size_t, f32p, C.c_int]
        L.sgo_stream_run.restype = C.c_size_t
        L.sgo2d_config_ok.argtypes = [C.c_int] * 5 + [C.c_float] * 2
        L.sgo2d_weights.argtypes = [C.c_int] * 5 + [f32p]
        L.sgo2d_scale.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float]
        L.sgo2d_scale.restype = C.c_float
        L.sgo2d_apply_valid.argtypes = [C.c_int, C.c_int, f32p, C.c_float, f32p, C.c_int, C.c_int,
                                        C.c_int, f32p, C.c_int]
        L.sgo2d_apply.argtypes = [C.c_int, C.c_int, f32p, C.c_float, f32p, C.c_int, C.c_int,
                                  C.c_int, f32p, C.c_int, C.c_int]
        L.sgh_apply_rows.argtypes = [C.c_void_p, C.c_void_p, f32p, f32p, C.c_size_t, C.c_size_t,
                                     C.c_size_t, C.c_size_t, C.c_int]
        L.sgh_apply2d_images.argtypes = [C.c_void_p, C.c_void_p, f32p, f32p, C.c_size_t, C.c_int,
                                         C.c_int, C.c_int, C.c_int]
        L.sgh_stream_channels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, f32p, f32p,
                                          C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int,
                                          C.c_int]
        L.sgh_valid_chunks.argtypes = [C.c_void_p, C.c_void_p, f32p, f32p, C.c_size_t, C.c_int, C.c_size_t, C.c_int]
        _lib = L
    return _lib


# --------------------------------------------------------------------------
# "port" oracle: numpy-level helpers over liboracle.so
# --------------------------------------------------------------------------
BOUNDARY = {"polynomial": 0, "reflect": 1, "periodic": 2, "constant": 3}


class Filter1D:
    """Oracle-side filter: weights + dt_inv (ref: src/savgolFilter.c:688-718)."""

    def __init__(self, n, m, d=0, dt=1.0, boundary="polynomial"):
        L = lib()
        if L.sgo_config_ok(n, m, d, dt) != 0:
            raise ValueError("invalid config")
        self.n, self.m, self.d, self.dt = n, m, d, float(dt)
        self.mode = BOUNDARY[boundary] if isinstance(boundary, str) else int(boundary)
        self.center = np.zeros(MAX_WS, np.float32)
        self.edge = np.zeros((MAX_N, MAX_WS), np.float32)
        rc = L.sgo_weights_1d(n, m, d, _fp(self.center), _fp(self.edge))
        assert rc == 0
        self.dt_inv = float(L.sgo_dt_inv(dt, d))

    def apply(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        if x.ndim == 1:
            y = np.empty_like(x)
            rc = lib().sgo_apply(self.n, _fp(self.center), _fp(self.edge), self.dt_inv, self.mode,
                                 _fp(x), _fp(y), x.size)
            if rc != 0:
                raise ValueError("apply failed")
            return y
        y = np.empty_like(x)
        rows, L_ = x.shape
        rc = lib().sgo_apply_batch(self.n, _fp(self.center), _fp(self.edge), self.dt_inv, self.mode,
                                   _fp(x), _fp(y), rows, L_, L_, L_)
        if rc != 0:
            raise ValueError("apply failed")
        return y

    def apply_valid(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        y = np.empty(max(x.size - 2 * self.n, 0), np.float32)
        k = lib().sgo_apply_valid(self.n, _fp(self.center), self.dt_inv, _fp(x), x.size, _fp(y))
        return y[:k]

    def apply_strided(self, inbuf: np.ndarray, istride, ioff, outbuf: np.ndarray, ostride, ooff, count):
        return lib().sgo_apply_strided(self.n, _fp(self.center), _fp(self.edge), self.dt_inv,
                                       inbuf.ctypes.data, istride, ioff,
                                       outbuf.ctypes.data, ostride, ooff, count)

    def stream_run(self, x: np.ndarray, with_leading=True) -> np.ndarray:
        x = np.ascontiguousarray(x, np.float32)
        y = np.empty(x.size + 4, np.float32)
        k = lib().sgo_stream_run(self.n, _fp(self.center), _fp(self.edge), self.dt_inv,
                                 _fp(x), x.size, _fp(y), 1 if with_leading else 0)
        return y[:k].copy()


BOUNDARY2D = {"valid": 0, "constant": 1, "reflect": 2}


class Filter2D:
    """ref: src/savgol2d.c:304-342"""

    def __init__(self, nx, ny, order, dx=0, dy=0, hx=1.0, hy=1.0):
        L = lib()
        if L.sgo2d_config_ok(nx, ny, order, dx, dy, hx, hy) != 0:
            raise ValueError("invalid 2d config")
        self.nx, self.ny, self.order, self.dx, self.dy = nx, ny, order, dx, dy
        self.W = np.zeros((2 * ny + 1, 2 * nx + 1), np.float32)
        assert L.sgo2d_weights(nx, ny, order, dx, dy, _fp(self.W)) == 0
        self.scale = float(L.sgo2d_scale(dx, dy, hx, hy))

    def apply(self, img: np.ndarray, boundary="constant", out: np.ndarray | None = None) -> np.ndarray:
        img = np.ascontiguousarray(img, np.float32)
        rows, cols = img.shape
        if out is None:
            out = np.zeros_like(img)
        b = BOUNDARY2D[boundary] if isinstance(boundary, str) else int(boundary)
        rc = lib().sgo2d_apply(self.nx, self.ny, _fp(self.W), self.scale, _fp(img), rows, cols, cols,
                               _fp(out), cols, b)
        if rc != 0:
            raise ValueError("apply2d failed")
        return out

    def apply_valid(self, img: np.ndarray) -> np.ndarray:
        img = np.ascontiguousarray(img, np.float32)
        rows, cols = img.shape
        out = np.zeros((rows - 2 * self.ny, cols - 2 * self.nx), np.float32)
        rc = lib().sgo2d_apply_valid(self.nx, self.ny, _fp(self.W), self.scale, _fp(img), rows, cols,
                                     cols, _fp(out), out.shape[1])
        if rc != 0:
            raise ValueError("apply2d_valid failed")
        return out


# --------------------------------------------------------------------------
# "reference": the unmodified reference .so, with its own ABI
# --------------------------------------------------------------------------
class SavgolConfig(C.Structure):
    """ref: include/iterative/savgolFilter.h:92-98 (12 bytes)"""
    _fields_ = [("half_window", C.c_uint8), ("poly_order", C.c_uint8), ("derivative", C.c_uint8),
                ("time_step", C.c_float), ("boundary", C.c_int)]


class SavgolFilter(C.Structure):
    """ref: include/iterative/savgolFilter.h:107-113 (8600 bytes)"""
    _fields_ = [("config", SavgolConfig), ("window_size", C.c_int), ("dt_scale", C.c_float),
                ("center_weights", C.c_float * MAX_WS), ("edge_weights", (C.c_float * MAX_WS) * MAX_N)]


class SavgolStream(C.Structure):
    """ref: include/iterative/savgol_stream.h:29-37 (296 bytes)"""
    _fields_ = [("filter", C.POINTER(SavgolFilter)), ("buffer", C.c_float * MAX_WS),
                ("write_pos", C.c_int), ("samples_received", C.c_size_t),
                ("samples_output", C.c_size_t), ("owns_filter", C.c_bool), ("dt_inv", C.c_float)]


class Savgol2DConfig(C.Structure):
    """ref: include/iterative/savgol2d.h:82-90 (16 bytes)"""
    _fields_ = [("half_window_x", C.c_uint8), ("half_window_y", C.c_uint8), ("poly_order", C.c_uint8),
                ("deriv_x", C.c_uint8), ("deriv_y", C.c_uint8), ("delta_x", C.c_float),
                ("delta_y", C.c_float)]


class Savgol2DFilter(C.Structure):
    """ref: include/iterative/savgol2d.h:95-103 (48 bytes)"""
    _fields_ = [("config", Savgol2DConfig), ("window_width", C.c_int), ("window_height", C.c_int),
                ("window_area", C.c_int), ("num_terms", C.c_int), ("scale", C.c_float),
                ("weights", f32p)]


def bind_reference_abi(L: C.CDLL) -> C.CDLL:
    """Attach the reference's prototypes to a library exporting its symbols.

    Used both for oracle/_ref/libsavgol_ref.so and (in tests) for the product
    libsavgol_b200.so, which exports the same symbols."""
    FP = C.POINTER(SavgolFilter)
    L.savgol_create.argtypes = [C.POINTER(SavgolConfig)]
    L.savgol_create.restype = FP
    L.savgol_destroy.argtypes = [FP]
    L.savgol_destroy.restype = None
    L.savgol_apply.argtypes = [FP, C.c_void_p, C.c_void_p, C.c_size_t]
    L.savgol_apply.restype = C.c_int
    L.savgol_apply_valid.argtypes = [FP, C.c_void_p, C.c_size_t, C.c_void_p]
    L.savgol_apply_valid.restype = C.c_size_t
    L.savgol_apply_strided.argtypes = [FP, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                       C.c_size_t, C.c_size_t, C.c_size_t]
    L.savgol_apply_strided.restype = C.c_int
    SP = C.POINTER(SavgolStream)
    L.savgol_stream_create.argtypes = [C.POINTER(SavgolConfig)]
    L.savgol_stream_create.restype = SP
    L.savgol_stream_init.argtypes = [SP, FP]
    L.savgol_stream_init.restype = C.c_int
    L.savgol_stream_destroy.argtypes = [SP]
    L.savgol_stream_destroy.restype = None
    L.savgol_stream_reset.argtypes = [SP]
    L.savgol_stream_reset.restype = None
    L.savgol_stream_push.argtypes = [SP, C.c_float, C.POINTER(C.c_bool)]
    L.savgol_stream_push.restype = C.c_float
    L.savgol_stream_push_full.argtypes = [SP, C.c_float, f32p, C.c_int]
    L.savgol_stream_push_full.restype = C.c_int
    L.savgol_stream_flush.argtypes = [SP, f32p, C.c_int]
    L.savgol_stream_flush.restype = C.c_int
    L.savgol_stream_flush_leading.argtypes = [SP, f32p, C.c_int]
    L.savgol_stream_flush_leading.restype = C.c_int
    for name, rt in (("ready", C.c_bool), ("latency", C.c_size_t), ("buffered", C.c_size_t),
                     ("samples_received", C.c_size_t), ("samples_output", C.c_size_t)):
        fn = getattr(L, "savgol_stream_" + name)
        fn.argtypes = [SP]
        fn.restype = rt
    F2 = C.POINTER(Savgol2DFilter)
    L.savgol2d_create.argtypes = [C.POINTER(Savgol2DConfig)]
    L.savgol2d_create.restype = F2
    L.savgol2d_destroy.argtypes = [F2]
    L.savgol2d_destroy.restype = None
    L.savgol2d_config_valid.argtypes = [C.POINTER(Savgol2DConfig)]
    L.savgol2d_config_valid.restype = C.c_bool
    L.savgol2d_apply_valid.argtypes = [F2, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.savgol2d_apply_valid.restype = C.c_int
    L.savgol2d_apply.argtypes = [F2, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.savgol2d_apply.restype = C.c_int
    L.savgol2d_gradient.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int]
    L.savgol2d_gradient.restype = C.c_int
    L.savgol2d_hessian.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int]
    L.savgol2d_hessian.restype = C.c_int
    L.savgol2d_laplacian.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_float, C.c_float, C.c_int]
    L.savgol2d_laplacian.restype = C.c_int
    return L


_ref = None


def ref_path() -> str:
    return os.path.join(HERE, "_ref", "libsavgol_ref.so")


def have_ref() -> bool:
    return os.path.exists(ref_path())


def ref() -> C.CDLL:
    """The unmodified reference, compiled by oracle/Makefile."""
    global _ref
    if _ref is None:
        if not have_ref():
            build()
        _ref = bind_reference_abi(C.CDLL(ref_path()))
    return _ref


def make_config(n, m, d=0, dt=1.0, boundary=0) -> SavgolConfig:
    b = BOUNDARY[boundary] if isinstance(boundary, str) else int(boundary)
    return SavgolConfig(n, m, d, dt, b)


def fnptr(L: C.CDLL, name: str) -> int:
    return C.cast(getattr(L, name), C.c_void_p).value
