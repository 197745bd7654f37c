"""ORACLE — TEST INFRASTRUCTURE ONLY. ctypes front-end of oracle/_ref/libmsda_ref_cuda.so: the
reference's own CUDA forward kernel (MOTR/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299,923-954)
compiled for sm_100a by oracle/Makefile. Never imported by the product package."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_SO = Path(__file__).resolve().parent / "_ref" / "libmsda_ref_cuda.so"
_lib = None
_shape_cache = {}


def available() -> bool:
    return _SO.exists()


def msda_im2col(value: torch.Tensor, shapes, loc: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """value [B,S,H,D] fp32 cuda, loc [B,Q,H,L,P,2], weights [B,Q,H,L,P] -> [B,Q,H*D] (legacy op layout)."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(_SO))
    B, S, H, D = value.shape
    Q, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    key = (tuple(tuple(int(x) for x in hw) for hw in shapes), str(value.device))
    if key not in _shape_cache:  # built once per pyramid so the call is CUDA-graph capturable
        sh_ = torch.as_tensor(shapes, dtype=torch.int64, device=value.device)
        _shape_cache[key] = (sh_, torch.cat((sh_.new_zeros((1,)), sh_.prod(1).cumsum(0)[:-1])))
    sh, lsi = _shape_cache[key]
    out = torch.zeros(B, Q, H * D, dtype=torch.float32, device=value.device)
    rc = _lib.ref_msda_im2col_f32(C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_void_p(value.data_ptr()),
                                  C.c_void_p(sh.data_ptr()), C.c_void_p(lsi.data_ptr()), C.c_void_p(loc.data_ptr()),
                                  C.c_void_p(weights.data_ptr()), B, S, H, D, L, Q, P, C.c_void_p(out.data_ptr()))
    if rc != 0:
        raise RuntimeError(f"reference kernel launch failed: cudaError {rc}")
    return out


def msda_col2im(value: torch.Tensor, shapes, loc: torch.Tensor, weights: torch.Tensor, grad_out: torch.Tensor):
    """The reference's backward kernel: -> (grad_value, grad_loc, grad_weights), all fp32."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(_SO))
    B, S, H, D = value.shape
    Q, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    key = (tuple(tuple(int(x) for x in hw) for hw in shapes), str(value.device))
    if key not in _shape_cache:
        sh_ = torch.as_tensor(shapes, dtype=torch.int64, device=value.device)
        _shape_cache[key] = (sh_, torch.cat((sh_.new_zeros((1,)), sh_.prod(1).cumsum(0)[:-1])))
    sh, lsi = _shape_cache[key]
    gv, gl, gw = torch.zeros_like(value), torch.zeros_like(loc), torch.zeros_like(weights)
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    rc = _lib.ref_msda_col2im_f32(C.c_void_p(torch.cuda.current_stream().cuda_stream), p(grad_out.contiguous()), p(value),
                                  p(sh), p(lsi), p(loc), p(weights), B, S, H, D, L, Q, P, p(gv), p(gl), p(gw))
    if rc != 0:
        raise RuntimeError(f"reference kernel launch failed: cudaError {rc}")
    return gv, gl, gw
