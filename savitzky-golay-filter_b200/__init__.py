"""savgol_b200 -- Python host mirror of the reference's filter API over the CUDA C-ABI library.

Everything numerical happens in ``libsavgol_b200.so`` (hand-written sm_100a kernels); this
package only marshals pointers.  Inputs may be CUDA ``torch`` tensors (zero copy, asynchronous on
the current torch stream) or host ``numpy`` arrays / CPU tensors (staged through the GPU by the
library).  There is no CPU fallback: without the shared library the import fails, without a
B200-class device every apply raises.

The directory name ``savitzky-golay-filter_b200`` is not an importable identifier; use
``import savgol_b200`` (a shim at the repository root) or ``importlib.import_module``.

Class / method names follow the reference's C API:
``SavgolFilter``  <- savgol_create / savgol_apply* (include/iterative/savgolFilter.h:130-203)
``SavgolStream``  <- savgol_stream_*               (include/iterative/savgol_stream.h:49-126)
``Savgol2DFilter``<- savgol2d_*                    (include/iterative/savgol2d.h:126-269)
``SavgolMCStream``: multi-channel chunked stream (extension).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import (SavgolConfig, Savgol2DConfig, SavgolFilterStruct, SavgolStreamStruct,  # noqa: F401
                    Savgol2DFilterStruct, MAX_HALF_WINDOW, MAX_WINDOW)

BOUNDARY = {"polynomial": 0, "reflect": 1, "periodic": 2, "constant": 3}
BOUNDARY_2D = {"valid": 0, "constant": 1, "reflect": 2}

__all__ = ["SavgolFilter", "SavgolStream", "Savgol2DFilter", "SavgolMCStream", "SavgolConfig", "Savgol2DConfig",
           "lib", "device_ok", "set_exact", "launch_count", "gradient", "hessian", "laplacian"]


def lib() -> C.CDLL:
    return _capi.load()


def device_ok() -> bool:
    return bool(lib().savgol_b200_device_ok())


def set_exact(flag: bool) -> None:
    """True: reference summation order with unfused multiply/add (bit-identical, slower)."""
    lib().savgol_b200_set_exact(1 if flag else 0)


def launch_count() -> int:
    return int(lib().savgol_b200_launch_count())


# ----------------------------------------------------------------------------------------------
def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


def _prep(x, name="input"):
    """-> (pointer, keepalive, is_cuda).  float32, last-dim contiguous."""
    if _is_torch(x):
        import torch
        if x.dtype != torch.float32:
            raise TypeError(f"{name} must be float32")
        if x.dim() > 0 and x.stride(-1) != 1:
            raise ValueError(f"{name}: last dimension must be contiguous")
        if x.is_cuda:
            lib().savgol_b200_set_stream(C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
        return x.data_ptr(), x, x.is_cuda
    a = np.asarray(x)
    if a.dtype != np.float32:
        raise TypeError(f"{name} must be float32")
    if a.ndim > 0 and a.strides[-1] != 4:
        raise ValueError(f"{name}: last dimension must be contiguous")
    return a.ctypes.data, a, False


def _empty_like(x, shape=None):
    if _is_torch(x):
        import torch
        return torch.empty(tuple(shape) if shape is not None else x.shape, dtype=x.dtype, device=x.device)
    return np.empty(shape if shape is not None else np.asarray(x).shape, np.float32)


def _row_pitch(x) -> int:
    if _is_torch(x):
        return int(x.stride(0))
    return int(np.asarray(x).strides[0] // 4)


class SavgolFilter:
    """1D Savitzky-Golay filter (ref: savgol_create, src/savgolFilter.c:688-718)."""

    def __init__(self, half_window, poly_order, derivative=0, time_step=1.0, boundary="polynomial"):
        b = BOUNDARY[boundary] if isinstance(boundary, str) else int(boundary)
        self.config = SavgolConfig(int(half_window), int(poly_order), int(derivative), float(time_step), b)
        self._h = lib().savgol_create(C.byref(self.config))
        if not self._h:
            raise ValueError("savgol_create: invalid configuration")

    # -- public struct fields (ref: include/iterative/savgolFilter.h:107-113)
    @property
    def half_window(self) -> int:
        return int(self._h.contents.config.half_window)

    @property
    def window_size(self) -> int:
        return int(self._h.contents.window_size)

    @property
    def dt_scale(self) -> float:
        return float(self._h.contents.dt_scale)

    @property
    def center_weights(self) -> np.ndarray:
        return np.ctypeslib.as_array(self._h.contents.center_weights)[: self.window_size].copy()

    @property
    def edge_weights(self) -> np.ndarray:
        return np.ctypeslib.as_array(self._h.contents.edge_weights).reshape(MAX_HALF_WINDOW, MAX_WINDOW).copy()

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None):
            lib().savgol_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- apply family
    def apply(self, x, out=None):
        """savgol_apply for a 1D signal, savgol_apply_batch for a [signals, length] array."""
        xp, _kx, _ = _prep(x)
        if out is None:
            out = _empty_like(x)
        op, _ko, _ = _prep(out, "output")
        if x.ndim == 1:
            rc = lib().savgol_apply(self._h, xp, op, x.shape[0])
        elif x.ndim == 2:
            rc = lib().savgol_apply_batch(self._h, xp, op, x.shape[0], x.shape[1], _row_pitch(x), _row_pitch(out))
        else:
            raise ValueError("apply expects a 1D signal or a 2D [signals, length] batch")
        if rc != 0:
            raise RuntimeError("savgol_apply failed (see stderr)")
        return out

    def apply_valid(self, x, out=None):
        xp, _kx, _ = _prep(x)
        L = x.shape[0]
        n_out = max(L - 2 * self.half_window, 0)
        if out is None:
            out = _empty_like(x, (n_out,))
        op, _ko, _ = _prep(out, "output")
        k = lib().savgol_apply_valid(self._h, xp, L, op)
        return out[:k]

    def apply_strided(self, in_ptr, in_stride, in_offset, out_ptr, out_stride, out_offset, count) -> int:
        """Raw savgol_apply_strided (byte strides / offsets); returns the C return code."""
        return int(lib().savgol_apply_strided(self._h, in_ptr, in_stride, in_offset, out_ptr, out_stride, out_offset, count))

    def apply_halo(self, x, left=None, right=None, out=None):
        """One slice of a partitioned signal with explicit n-sample halos (device tensors)."""
        xp, _kx, _ = _prep(x)
        if out is None:
            out = _empty_like(x)
        op, _ko, _ = _prep(out, "output")
        lp = _prep(left, "left halo")[0] if left is not None else None
        rp = _prep(right, "right halo")[0] if right is not None else None
        n = self.half_window
        if left is not None and left.shape[0] != n or right is not None and right.shape[0] != n:
            raise ValueError("halos must hold exactly half_window samples")
        rc = lib().savgol_apply_halo(self._h, xp, op, x.shape[0], lp, rp)
        if rc != 0:
            raise RuntimeError("savgol_apply_halo failed (see stderr)")
        return out


class SavgolStream:
    """Single-channel sample-at-a-time stream (ref: src/savgol_stream.c:80-315); host arithmetic."""

    def __init__(self, half_window, poly_order, derivative=0, time_step=1.0):
        cfg = SavgolConfig(int(half_window), int(poly_order), int(derivative), float(time_step), 0)
        self._h = lib().savgol_stream_create(C.byref(cfg))
        if not self._h:
            raise ValueError("savgol_stream_create: invalid configuration")
        self._buf = (C.c_float * (MAX_HALF_WINDOW + 1))()

    def push(self, sample: float):
        valid = C.c_bool(False)
        v = lib().savgol_stream_push(self._h, float(sample), C.byref(valid))
        return (float(v), bool(valid.value))

    def push_full(self, sample: float):
        k = lib().savgol_stream_push_full(self._h, float(sample), self._buf, MAX_HALF_WINDOW + 1)
        return [self._buf[i] for i in range(k)]

    def flush(self):
        k = lib().savgol_stream_flush(self._h, self._buf, MAX_HALF_WINDOW + 1)
        return [self._buf[i] for i in range(max(k, 0))]

    def flush_leading(self):
        k = lib().savgol_stream_flush_leading(self._h, self._buf, MAX_HALF_WINDOW + 1)
        return [self._buf[i] for i in range(k)]

    def reset(self):
        lib().savgol_stream_reset(self._h)

    @property
    def ready(self):
        return bool(lib().savgol_stream_ready(self._h))

    @property
    def latency(self):
        return int(lib().savgol_stream_latency(self._h))

    @property
    def buffered(self):
        return int(lib().savgol_stream_buffered(self._h))

    @property
    def samples_received(self):
        return int(lib().savgol_stream_samples_received(self._h))

    @property
    def samples_output(self):
        return int(lib().savgol_stream_samples_output(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib().savgol_stream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SavgolMCStream:
    """Multi-channel chunked stream: ``channels`` lock-stepped streams, fixed latency half_window."""

    def __init__(self, channels, half_window, poly_order, derivative=0, time_step=1.0):
        cfg = SavgolConfig(int(half_window), int(poly_order), int(derivative), float(time_step), 0)
        self.n = int(half_window)
        self.channels = int(channels)
        self._h = lib().savgol_mcstream_create(C.byref(cfg), self.channels)
        if not self._h:
            raise RuntimeError("savgol_mcstream_create failed")

    def push(self, chunk, out=None):
        """chunk: [channels, K].  Returns (out, produced): out[:, :produced] are the new outputs."""
        xp, _k, _ = _prep(chunk)
        C_, K = chunk.shape
        if C_ != self.channels:
            raise ValueError("chunk must be [channels, K]")
        if out is None:
            out = _empty_like(chunk, (C_, K + self.n))
        op, _ko, _ = _prep(out, "output")
        k = lib().savgol_mcstream_push(self._h, xp, _row_pitch(chunk), K, op, _row_pitch(out))
        if k < 0:
            raise RuntimeError("savgol_mcstream_push failed (see stderr)")
        return out, int(k)

    def flush(self, like, out=None):
        if out is None:
            out = _empty_like(like, (self.channels, self.n))
        op, _ko, _ = _prep(out, "output")
        k = lib().savgol_mcstream_flush(self._h, op, _row_pitch(out))
        if k < 0:
            raise RuntimeError("savgol_mcstream_flush failed")
        return out, int(k)

    def reset(self):
        lib().savgol_mcstream_reset(self._h)

    def save(self) -> bytes:
        """Checkpoint (counters + per-channel carry state) as a host blob."""
        n = int(lib().savgol_mcstream_checkpoint_size(self._h))
        buf = C.create_string_buffer(n)
        if lib().savgol_mcstream_save(self._h, buf, n) != n:
            raise RuntimeError("savgol_mcstream_save failed")
        return buf.raw

    def restore(self, blob: bytes):
        """Resume from a blob written by save() of a stream with the same configuration."""
        if lib().savgol_mcstream_restore(self._h, blob, len(blob)) != 0:
            raise ValueError("savgol_mcstream_restore: checkpoint does not match this stream")

    @property
    def latency(self):
        return int(lib().savgol_mcstream_latency(self._h))

    @property
    def samples_received(self):
        return int(lib().savgol_mcstream_samples_received(self._h))

    @property
    def samples_output(self):
        return int(lib().savgol_mcstream_samples_output(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib().savgol_mcstream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Savgol2DFilter:
    """2D Savitzky-Golay filter (ref: savgol2d_create, src/savgol2d.c:304-342)."""

    def __init__(self, half_window_x, half_window_y, poly_order, deriv_x=0, deriv_y=0, delta_x=1.0, delta_y=1.0):
        self.config = Savgol2DConfig(int(half_window_x), int(half_window_y), int(poly_order), int(deriv_x), int(deriv_y),
                                     float(delta_x), float(delta_y))
        self._h = lib().savgol2d_create(C.byref(self.config))
        if not self._h:
            raise ValueError("savgol2d_create: invalid configuration")

    @property
    def weights(self) -> np.ndarray:
        f = self._h.contents
        return np.ctypeslib.as_array(f.weights, shape=(f.window_height, f.window_width)).copy()

    @property
    def scale(self) -> float:
        return float(self._h.contents.scale)

    @property
    def handle(self):
        return self._h

    def apply(self, img, boundary="constant", out=None):
        """savgol2d_apply for [rows, cols], savgol2d_apply_batch for [images, rows, cols]."""
        b = BOUNDARY_2D[boundary] if isinstance(boundary, str) else int(boundary)
        xp, _k, _ = _prep(img)
        if out is None:
            out = _empty_like(img)
            if b == 0:
                out[...] = 0
        op, _ko, _ = _prep(out, "output")
        if img.ndim == 2:
            rows, cols = img.shape
            rc = lib().savgol2d_apply(self._h, xp, rows, cols, _row_pitch(img), op, _row_pitch(out), b)
        elif img.ndim == 3:
            n, rows, cols = img.shape
            if _is_torch(img):
                ip, opi, istr, ostr = img.stride(0), out.stride(0), img.stride(1), out.stride(1)
            else:
                ip, opi, istr, ostr = img.strides[0] // 4, out.strides[0] // 4, img.strides[1] // 4, out.strides[1] // 4
            rc = lib().savgol2d_apply_batch(self._h, xp, rows, cols, istr, ip, op, ostr, opi, n, b)
        else:
            raise ValueError("apply expects [rows, cols] or [images, rows, cols]")
        if rc != 0:
            raise RuntimeError("savgol2d_apply failed (see stderr)")
        return out

    def apply_valid(self, img, out=None):
        xp, _k, _ = _prep(img)
        rows, cols = img.shape
        cfg = self.config
        if out is None:
            out = _empty_like(img, (rows - 2 * cfg.half_window_y, cols - 2 * cfg.half_window_x))
        op, _ko, _ = _prep(out, "output")
        rc = lib().savgol2d_apply_valid(self._h, xp, rows, cols, _row_pitch(img), op, _row_pitch(out))
        if rc != 0:
            raise RuntimeError("savgol2d_apply_valid failed")
        return out

    def apply_band(self, buf, top_halo, bottom_halo, boundary="constant", out=None, image_row0=0):
        """One horizontal band of a larger image: ``buf`` = [top_halo rows | band | bottom_halo rows] (device
        tensor); a halo of 0 rows marks an image border; ``image_row0`` = image row of buf[0] (makes the band
        round exactly like the whole-image call).  Returns the filtered band rows."""
        xp, _k, _ = _prep(buf)
        rows, cols = buf.shape
        if out is None:
            out = _empty_like(buf, (rows - top_halo - bottom_halo, cols))
        op, _ko, _ = _prep(out, "output")
        rc = lib().savgol2d_apply_band_at(self._h, xp, rows, cols, _row_pitch(buf), op, _row_pitch(out), BOUNDARY_2D[boundary],
                                          int(top_halo), int(bottom_halo), int(image_row0))
        if rc != 0:
            raise RuntimeError("savgol2d_apply_band failed (see stderr)")
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().savgol2d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gradient(img, half_win_x, half_win_y, poly_order, delta_x=1.0, delta_y=1.0, boundary="constant"):
    """savgol2d_gradient (ref: src/savgol2d.c:462-499) -> (grad_x, grad_y)."""
    b = BOUNDARY_2D[boundary]
    xp, _k, _ = _prep(img)
    gx, gy = _empty_like(img), _empty_like(img)
    rows, cols = img.shape
    rc = lib().savgol2d_gradient(half_win_x, half_win_y, poly_order, xp, rows, cols, _row_pitch(img),
                                 _prep(gx)[0], _prep(gy)[0], delta_x, delta_y, b)
    if rc != 0:
        raise RuntimeError("savgol2d_gradient failed")
    return gx, gy


def hessian(img, half_win_x, half_win_y, poly_order, delta_x=1.0, delta_y=1.0, boundary="constant"):
    """savgol2d_hessian (ref: src/savgol2d.c:501-558) -> (hxx, hxy, hyy)."""
    b = BOUNDARY_2D[boundary]
    xp, _k, _ = _prep(img)
    hxx, hxy, hyy = _empty_like(img), _empty_like(img), _empty_like(img)
    rows, cols = img.shape
    rc = lib().savgol2d_hessian(half_win_x, half_win_y, poly_order, xp, rows, cols, _row_pitch(img),
                                _prep(hxx)[0], _prep(hxy)[0], _prep(hyy)[0], delta_x, delta_y, b)
    if rc != 0:
        raise RuntimeError("savgol2d_hessian failed")
    return hxx, hxy, hyy


def laplacian(img, half_win_x, half_win_y, poly_order, delta_x=1.0, delta_y=1.0, boundary="constant"):
    """savgol2d_laplacian (ref: src/savgol2d.c:560-618)."""
    b = BOUNDARY_2D[boundary]
    xp, _k, _ = _prep(img)
    out = _empty_like(img)
    rows, cols = img.shape
    rc = lib().savgol2d_laplacian(half_win_x, half_win_y, poly_order, xp, rows, cols, _row_pitch(img),
                                  _prep(out)[0], delta_x, delta_y, b)
    if rc != 0:
        raise RuntimeError("savgol2d_laplacian failed")
    return out
