"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes front-end of oracle/msda_core.c (built by oracle/Makefile)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libmsda_oracle.so"
_lib = None


def build() -> Path:
    subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not _SO.exists():
            build()
        _lib = C.CDLL(str(_SO))
    return _lib


def msda_core(value: np.ndarray, shapes, loc: np.ndarray, weights: np.ndarray) -> np.ndarray:
    """value [B,Lv,H,D], loc [B,Q,H,L,P,2], weights [B,Q,H,L,P] (float32 or float64) -> [B,Q,H*D]."""
    dt = value.dtype
    assert dt in (np.float32, np.float64) and loc.dtype == dt and weights.dtype == dt
    value, loc, weights = map(np.ascontiguousarray, (value, loc, weights))
    B, Lv, H, D = value.shape
    Q, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    hw = np.asarray(shapes, dtype=np.int32).reshape(-1)
    out = np.empty((B, Q, H * D), dtype=dt)
    fn = _load().msda_core_f32 if dt == np.float32 else _load().msda_core_f64
    fn.restype = None
    fn(value.ctypes.data_as(C.c_void_p), hw.ctypes.data_as(C.c_void_p), C.c_int(L), C.c_int(B), C.c_int64(Lv),
       C.c_int(H), C.c_int(D), loc.ctypes.data_as(C.c_void_p), weights.ctypes.data_as(C.c_void_p), C.c_int(Q),
       C.c_int(P), out.ctypes.data_as(C.c_void_p))
    return out


def msda_core_backward(value: np.ndarray, shapes, loc: np.ndarray, weights: np.ndarray, grad_out: np.ndarray):
    """Gradients of msda_core w.r.t. (value, loc, weights); grad_out [B,Q,H*D]."""
    dt = value.dtype
    assert dt in (np.float32, np.float64) and loc.dtype == dt and weights.dtype == dt and grad_out.dtype == dt
    value, loc, weights, grad_out = map(np.ascontiguousarray, (value, loc, weights, grad_out))
    B, Lv, H, D = value.shape
    Q, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    hw = np.asarray(shapes, dtype=np.int32).reshape(-1)
    gv, gl, gw = np.zeros_like(value), np.zeros_like(loc), np.zeros_like(weights)
    fn = _load().msda_core_backward_f32 if dt == np.float32 else _load().msda_core_backward_f64
    fn.restype = None
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    fn(p(value), p(hw), C.c_int(L), C.c_int(B), C.c_int64(Lv), C.c_int(H), C.c_int(D), p(loc), p(weights), C.c_int(Q),
       C.c_int(P), p(grad_out), p(gv), p(gl), p(gw))
    return gv, gl, gw
