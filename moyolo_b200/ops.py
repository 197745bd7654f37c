"""Tensor-level wrappers over the C ABI. PyTorch is used for device memory and streams only.

Every function requires CUDA tensors and enqueues on `torch.cuda.current_stream()`; outputs are
allocated here (the C library never allocates), see include/moyolo_b200.h.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import BF16, F32, F64

_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.float64: F64}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise RuntimeError(f"moyolo_b200: unsupported dtype {t.dtype}") from None


def _cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            # same wording as the reference extension (MOTR/models/ops/src/ms_deform_attn.h:36)
            raise RuntimeError("Not implemented on the CPU")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# offsets|logits projection inside the gather kernel (MOYOLO_PROJ_FUSED=0 disables; row bound see proj_fused_supported)
import os as _os  # noqa: E402
PROJ_FUSED = _os.environ.get("MOYOLO_PROJ_FUSED", "1") != "0"
PROJ_FUSED_MAX_ROWS = int(_os.environ.get("MOYOLO_PROJ_FUSED_MAX_ROWS", "1024"))


def launch_count() -> int:
    """Kernels launched (or recorded into a graph being captured) through the C ABI by this host thread so far:
    a monotonic thread-local counter kept by the library (moyolo_launch_count); callers take differences."""
    return int(_lib.lib().moyolo_launch_count())


class GatherProbe:
    """Instrumentation handed to a gather call by its caller (bench.py's roofline leg): `pre(B, Lv, C, R, H, L, P,
    value_bytes)` / `post()` run around the launch and the (idempotent) launch is repeated `repeat` times between
    them, so the ~5 us of a CUDA-event node pair inside a graph is amortised."""
    __slots__ = ("pre", "post", "repeat")

    def __init__(self, pre, post, repeat: int = 1):
        self.pre, self.post, self.repeat = pre, post, repeat


def _shapes_arr(shapes: Sequence[Sequence[int]]):
    flat = [int(v) for hw in shapes for v in hw]
    return (C.c_int32 * len(flat))(*flat), len(flat) // 2


def msda_sampled(value: torch.Tensor, shapes, loc: torch.Tensor, weights: torch.Tensor,
                 row_offsets: Optional[torch.Tensor] = None) -> torch.Tensor:
    """multi_scale_deformable_attn_pytorch (utils.py:41-78) / legacy ms_deform_attn_forward.

    value [B, Lv, H, Dh] (last two dims contiguous), loc [B, Q, H, L, P, 2], weights [B, Q, H, L, P]
    -> [B, Q, H*Dh].
    """
    _cuda(value, loc, weights)
    B, Lv, H, Dh = value.shape
    if value.stride(3) != 1 or value.stride(2) != Dh:
        value = value.contiguous()
    loc, weights = loc.contiguous(), weights.contiguous()
    _, Q, _, L, P, _ = loc.shape
    arr, n_levels = _shapes_arr(shapes)
    if n_levels != L:
        raise ValueError(f"value_shapes has {n_levels} levels but sampling locations have {L}")
    out = torch.empty(B, Q, H * Dh, dtype=value.dtype, device=value.device)
    if B * Q == 0:
        return out
    _lib.check(_lib.lib().moyolo_msda_sampled_forward(
        value.data_ptr(), _dt(value), value.stride(0), value.stride(1), arr, L, B, Lv, H, Dh, P,
        loc.data_ptr(), weights.data_ptr(), _dt(loc), B * Q, _ptr(row_offsets), out.data_ptr(), H * Dh,
        _stream()))
    return out


def msda_sampled_backward(value: torch.Tensor, shapes, loc: torch.Tensor, weights: torch.Tensor,
                          grad_out: torch.Tensor, row_offsets: Optional[torch.Tensor] = None):
    """Gradients of msda_sampled w.r.t. (value, loc, weights) -- the legacy ms_deform_attn_backward
    (MOTR/models/ops/src/cuda/ms_deform_attn_cuda.cu:83-153). grad_out [B, Q, H*Dh].
    Returns (grad_value [B, Lv, H, Dh], grad_loc like loc, grad_weights like weights); fp32 for fp32 / bf16
    value, fp64 for fp64 value."""
    _cuda(value, loc, weights, grad_out)
    B, Lv, H, Dh = value.shape
    if value.stride(3) != 1 or value.stride(2) != Dh:
        value = value.contiguous()
    gdt = torch.float64 if value.dtype == torch.float64 else torch.float32
    loc, weights = loc.contiguous().to(gdt), weights.contiguous().to(gdt)
    _, Q, _, L, P, _ = loc.shape
    arr, n_levels = _shapes_arr(shapes)
    if n_levels != L:
        raise ValueError(f"value_shapes has {n_levels} levels but sampling locations have {L}")
    grad_out = grad_out.reshape(B * Q, H * Dh).to(gdt)
    if grad_out.stride(1) != 1:
        grad_out = grad_out.contiguous()
    grad_value = torch.zeros(B, Lv, H, Dh, dtype=gdt, device=value.device)
    grad_loc = torch.zeros_like(loc)
    grad_w = torch.zeros_like(weights)
    if B * Q == 0:
        return grad_value, grad_loc, grad_w
    _lib.check(_lib.lib().moyolo_msda_sampled_backward(
        value.data_ptr(), _dt(value), value.stride(0), value.stride(1), arr, L, B, Lv, H, Dh, P, loc.data_ptr(),
        weights.data_ptr(), _dt(loc), grad_out.data_ptr(), grad_out.stride(0), B * Q, _ptr(row_offsets),
        grad_value.data_ptr(), grad_loc.data_ptr(), grad_w.data_ptr(), _stream()))
    return grad_value, grad_loc, grad_w


def msda_fused_headmajor(value_hm: torch.Tensor, shapes, offsets: torch.Tensor, logits: torch.Tensor, refer: torch.Tensor,
                         n_points: int, softmax_mode: int = _lib.SOFTMAX, row_offsets: Optional[torch.Tensor] = None,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`msda_fused` on a head-major value tensor [B, H, Lv, Dh] (experimental layout: the x0 / x1 corners of a
    sampling point are contiguous; benchmarks/headmajor_probe.py). bf16, 8 heads x 32, 3 levels x 4 points."""
    _cuda(value_hm, offsets, logits, refer)
    B, H, Lv, Dh = value_hm.shape
    if value_hm.stride(3) != 1:
        raise ValueError("value must be [B, H, Lv, Dh] with contiguous channels")
    R = offsets.shape[0]
    arr, L = _shapes_arr(shapes)
    refer = refer.contiguous()
    if out is None:
        out = torch.empty(R, H * Dh, dtype=value_hm.dtype, device=value_hm.device)
    _lib.check(_lib.lib().moyolo_msda_fused_forward_headmajor(
        value_hm.data_ptr(), _dt(value_hm), value_hm.stride(0), value_hm.stride(1), value_hm.stride(2), arr, L, B, Lv, H, Dh,
        n_points, offsets.data_ptr(), offsets.stride(0), logits.data_ptr(), logits.stride(0), refer.data_ptr(),
        refer.shape[1], refer.shape[2], softmax_mode, R, _ptr(row_offsets), out.data_ptr(), out.stride(0), _stream()))
    return out


def msda_fused(value: torch.Tensor, shapes, offsets: torch.Tensor, logits: torch.Tensor, refer: torch.Tensor,
               n_heads: int, n_points: int, batch: int, softmax_mode: int = _lib.SOFTMAX,
               row_offsets: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
               probe: Optional[GatherProbe] = None) -> torch.Tensor:
    """Fused softmax + location + bilinear gather (transformer.py:268-285).

    value  [B, Lv, C] view (last dim contiguous; arbitrary batch/position strides)
    offsets [R, H*L*P*2] fp32 view, logits [R, H*L*P] fp32 view (row-strided slices allowed)
    refer  [R, ref_levels, 2|4] fp32 contiguous
    -> [R, C] of value.dtype
    """
    _cuda(value, offsets, logits, refer)
    B, Lv, Cc = value.shape
    if B != batch or value.stride(2) != 1:
        raise ValueError("value must be [B, Lv, C] with contiguous channels")
    Dh = Cc // n_heads
    R = offsets.shape[0]
    arr, L = _shapes_arr(shapes)
    if offsets.stride(1) != 1 or logits.stride(1) != 1 or offsets.dtype != torch.float32 or \
            logits.dtype != torch.float32:
        raise ValueError("offsets/logits must be fp32 with contiguous columns")
    refer = refer.contiguous()
    if refer.dtype != torch.float32:
        raise ValueError("refer must be fp32")
    if out is None:
        out = torch.empty(R, Cc, dtype=value.dtype, device=value.device)
    if R == 0:
        return out
    if probe is not None:
        probe.pre(B, Lv, Cc, R, n_heads, L, n_points, value.element_size())
    for _ in range(probe.repeat if probe is not None else 1):
        _lib.check(_lib.lib().moyolo_msda_fused_forward(
            value.data_ptr(), _dt(value), value.stride(0), value.stride(1), arr, L, B, Lv, n_heads, Dh, n_points,
            offsets.data_ptr(), offsets.stride(0), logits.data_ptr(), logits.stride(0), refer.data_ptr(),
            refer.shape[1], refer.shape[2], softmax_mode, R, _ptr(row_offsets), out.data_ptr(), out.stride(0),
            _stream()))
    if probe is not None:
        probe.post()
    return out


def msda_proj_fused(value: torch.Tensor, shapes, xq: torch.Tensor, w_offlog: torch.Tensor, b_offlog: torch.Tensor,
                    refer: torch.Tensor, n_heads: int, n_points: int, batch: int,
                    softmax_mode: int = _lib.SOFTMAX, row_offsets: Optional[torch.Tensor] = None,
                    out: Optional[torch.Tensor] = None, probe: Optional[GatherProbe] = None) -> torch.Tensor:
    """msda_fused with the offsets|logits projection (xq . w_offlog^T + b_offlog) inside the gather kernel:
    transformer.py:268-285 in one launch. bf16, 8 heads x 32, 3 levels x 4 points."""
    _cuda(value, xq, w_offlog, b_offlog, refer)
    B, Lv, Cc = value.shape
    if B != batch or value.stride(2) != 1:
        raise ValueError("value must be [B, Lv, C] with contiguous channels")
    R = xq.shape[0]
    arr, L = _shapes_arr(shapes)
    if xq.stride(1) != 1 or xq.dtype != torch.bfloat16 or w_offlog.dtype != torch.bfloat16 or not w_offlog.is_contiguous():
        raise ValueError("xq / w_offlog must be bf16 with contiguous columns")
    refer = refer.contiguous()
    if out is None:
        out = torch.empty(R, Cc, dtype=value.dtype, device=value.device)
    if R == 0:
        return out
    if probe is not None:
        probe.pre(B, Lv, Cc, R, n_heads, L, n_points, value.element_size())
    for _ in range(probe.repeat if probe is not None else 1):
        _lib.check(_lib.lib().moyolo_msda_proj_fused_forward(
            value.data_ptr(), _dt(value), value.stride(0), value.stride(1), arr, L, B, Lv, n_heads, Cc // n_heads,
            n_points, xq.data_ptr(), xq.stride(0), w_offlog.data_ptr(), b_offlog.data_ptr(), refer.data_ptr(),
            refer.shape[1], refer.shape[2], softmax_mode, R, _ptr(row_offsets), out.data_ptr(), out.stride(0),
            _stream()))
    if probe is not None:
        probe.post()
    return out


def proj_fused_supported(value_dtype, n_heads: int, head_dim: int, n_levels: int, n_points: int, rows: int) -> bool:
    """Configurations served by msda_proj_fused; beyond ~1k rows the per-CTA re-read of the projection weights
    costs more than the saved launch (the tcgen05 GEMM + pair gather then win)."""
    return (PROJ_FUSED and value_dtype == torch.bfloat16 and n_heads == 8 and head_dim == 32 and n_levels == 3
            and n_points == 4 and rows <= PROJ_FUSED_MAX_ROWS)


# fp32 Linear layers on the tensor cores (MOYOLO_FP32_TENSOR=0: the CUDA-core kernel linear_simt instead): both
# operands are expanded into three bf16 terms laid out along K (moyolo_split_bf16x3) and ONE bf16 tcgen05 GEMM over
# K' = 6K with fp32 accumulation gives the fp32 product to ~3e-8 relative to rms.
FP32_TENSOR = _os.environ.get("MOYOLO_FP32_TENSOR", "1") != "0"


def split_bf16x3(x: torch.Tensor, role: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[M, K] fp32 -> [M, 6K] bf16 three-term expansion in activation (role 0) or weight (role 1) block order."""
    M, K = x.shape
    if out is None:
        out = torch.empty(M, 6 * K, dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().moyolo_split_bf16x3(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), M, K, role, _stream()))
    return out


def _fp32_tensor_ok(x, w, M, N, K, engine) -> bool:
    return (FP32_TENSOR and x.dtype == torch.float32 and engine != _lib.GEMM_SIMT and M > 0 and K % 64 == 0 and
            N % 32 == 0 and x.stride(0) % 4 == 0)


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], out_dtype: torch.dtype = None,
           relu: bool = False, zero_rows: Optional[torch.Tensor] = None, engine: int = _lib.GEMM_AUTO,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = act(x . w^T + b); x [M, K] (row stride free), w [N, K] contiguous, b fp32 [N] or None."""
    _cuda(x, w, b)
    M, K = x.shape
    N = w.shape[0]
    if x.stride(1) != 1:
        x = x.contiguous()
    if not w.is_contiguous() or w.shape[1] != K or w.dtype != x.dtype:
        raise ValueError("linear: w must be contiguous [N, K] of x.dtype")
    if b is not None and (b.dtype != torch.float32 or not b.is_contiguous()):
        raise ValueError("linear: bias must be contiguous fp32")
    out_dtype = out_dtype or x.dtype
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=x.device)
    if _fp32_tensor_ok(x, w, M, N, K, engine):
        # the weight's expansion is cached ON the tensor object (it lives and dies with it; re-made after an in-place
        # update): packed weights (executor.LinearPack) are persistent objects, so this runs once per weight
        cached = getattr(w, "_moyolo_split", None)
        if cached is None or cached[0] != w._version:
            # The GEMM kernels fetch WEIGHTS before their programmatic-dependency wait (weights are immutable while
            # frames run, common.cuh), so a freshly written expansion must be complete before the first GEMM that
            # reads it is even launched: one host synchronisation per weight, at first use (never in steady state).
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("moyolo_b200: an fp32 weight expansion would be created inside a CUDA graph capture; "
                                   "run the captured code once eagerly first")
            cached = (w._version, split_bf16x3(w, 1))
            torch.cuda.current_stream().synchronize()
            w._moyolo_split = cached
        w6 = cached[1]
        x6 = split_bf16x3(x, 0)
        _lib.check(_lib.lib().moyolo_linear(
            x6.data_ptr(), x6.stride(0), w6.data_ptr(), _ptr(b), out.data_ptr(), out.stride(0), M, N, 6 * K, BF16,
            _dt(out), _lib.EPI_RELU if relu else _lib.EPI_NONE, _ptr(zero_rows), _lib.GEMM_TCGEN05, _stream()))
        return out
    _lib.check(_lib.lib().moyolo_linear(
        x.data_ptr(), x.stride(0), w.data_ptr(), _ptr(b), out.data_ptr(), out.stride(0), M, N, K, _dt(x),
        _dt(out), _lib.EPI_RELU if relu else _lib.EPI_NONE, _ptr(zero_rows), engine, _stream()))
    return out


def linear_tall(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], out: torch.Tensor,
                zero_rows: Optional[torch.Tensor] = None, max_ctas: int = 0) -> torch.Tensor:
    """y = x . w^T + b for tall x [M, 256] (bf16) through the persistent weight-resident kernel, on at most
    `max_ctas` SMs (0 = all)."""
    _cuda(x, w, b, out)
    M, K = x.shape
    N = w.shape[0]
    if x.stride(1) != 1 or out.stride(1) != 1 or not w.is_contiguous() or x.dtype != torch.bfloat16 or \
            out.dtype != torch.bfloat16:
        raise ValueError("linear_tall: bf16 x / w / out with contiguous columns")
    _lib.check(_lib.lib().moyolo_linear_tall(x.data_ptr(), x.stride(0), w.data_ptr(), _ptr(b), out.data_ptr(),
                                             out.stride(0), M, N, K, _ptr(zero_rows), int(max_ctas), _stream()))
    return out


def linear_dual(x1: torch.Tensor, x2: torch.Tensor, n_split: int, w: torch.Tensor, b: Optional[torch.Tensor],
                out: torch.Tensor) -> torch.Tensor:
    """out[:, :n_split] = x1 . w[:n_split]^T, out[:, n_split:] = x2 . w[n_split:]^T (+ b) in one launch
    (bf16 operands, tcgen05). The MHA in-projection with q = k = x + pos, v = x (transformer.py:637-638)."""
    _cuda(x1, x2, w, b, out)
    M, K = x1.shape
    N = w.shape[0]
    if x1.stride(1) != 1 or x2.stride(1) != 1 or not w.is_contiguous() or x2.shape != x1.shape:
        raise ValueError("linear_dual: x1/x2 must be [M, K] with contiguous columns, w contiguous [N, K]")
    _lib.check(_lib.lib().moyolo_linear_dual(
        x1.data_ptr(), x1.stride(0), x2.data_ptr(), x2.stride(0), int(n_split), w.data_ptr(), _ptr(b), out.data_ptr(),
        out.stride(0), M, N, K, _dt(out), _stream()))
    return out


def linear_add_layernorm(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], residual: Optional[torch.Tensor],
                         gamma: torch.Tensor, beta: torch.Tensor, eps: float, out_f32: Optional[torch.Tensor] = None,
                         out_lp: Optional[torch.Tensor] = None, pos: Optional[torch.Tensor] = None,
                         out_pos: Optional[torch.Tensor] = None) -> None:
    """LayerNorm(x . w^T + b + residual) with the GEMM, the residual add and the LayerNorm in ONE launch
    (N == 256, bf16 operands; outputs written through the given buffers, any of which may be None)."""
    _cuda(x, w, b, residual, gamma, beta, out_f32, out_lp, pos, out_pos)
    M, K = x.shape
    N = w.shape[0]
    if x.stride(1) != 1 or not w.is_contiguous():
        raise ValueError("linear_add_layernorm: x must have contiguous columns, w contiguous [N, K]")
    for t in (residual, out_f32, out_lp, pos, out_pos):
        if t is not None and (not t.is_contiguous() or t.shape != (M, N)):
            raise ValueError("linear_add_layernorm: row buffers must be contiguous [M, N]")
    _lib.check(_lib.lib().moyolo_linear_add_layernorm(
        x.data_ptr(), x.stride(0), w.data_ptr(), _ptr(b), _ptr(residual), gamma.data_ptr(), beta.data_ptr(),
        float(eps), M, N, K, _ptr(out_f32), _ptr(out_lp), _ptr(pos), _ptr(out_pos), _stream()))


# One-launch FFN block (moyolo_ffn_add_layernorm). Correct and bit-identical to the two launches it replaces, but
# MEASURED SLOWER on B200 (382 rows: 12.6 us vs 10.7 us for hidden 1024, 8.8 vs 7.8 us for hidden 256): publishing the
# hidden slab through L2 behind a cluster barrier costs more than the programmatic-dependent-launch boundary it
# removes. Off by default; MOYOLO_FFN_FUSED=1 selects it.
FFN_FUSED = _os.environ.get("MOYOLO_FFN_FUSED", "0") == "1"


def ffn_fused_supported(dt, C: int, F: int) -> bool:
    return FFN_FUSED and dt == torch.bfloat16 and C == 256 and F in (256, 1024)


def ffn_add_layernorm(x: torch.Tensor, w1: torch.Tensor, b1, w2: torch.Tensor, b2, h: torch.Tensor,
                      residual: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor, eps: float,
                      out_f32: Optional[torch.Tensor] = None, out_lp: Optional[torch.Tensor] = None,
                      pos: Optional[torch.Tensor] = None, out_pos: Optional[torch.Tensor] = None) -> None:
    """LayerNorm(residual + relu(x . w1^T + b1) . w2^T + b2) in ONE launch (d_model 256, hidden 256 / 1024, bf16
    operands); h is the bf16 [M, F] scratch for the hidden activations."""
    _cuda(x, w1, b1, w2, b2, h, residual, gamma, beta, out_f32, out_lp, pos, out_pos)
    M, Cc = x.shape
    F = w1.shape[0]
    if x.stride(1) != 1 or not w1.is_contiguous() or not w2.is_contiguous() or not h.is_contiguous() or \
            h.shape[0] < M or h.shape[1] != F or w2.shape != (Cc, F):
        raise ValueError("ffn_add_layernorm: x [M,C] contiguous columns, w1 [F,C], w2 [C,F], h [>=M,F] contiguous")
    for t in (residual, out_f32, out_lp, pos, out_pos):
        if t is not None and (not t.is_contiguous() or t.shape != (M, Cc)):
            raise ValueError("ffn_add_layernorm: row buffers must be contiguous [M, C]")
    _lib.check(_lib.lib().moyolo_ffn_add_layernorm(
        x.data_ptr(), x.stride(0), w1.data_ptr(), _ptr(b1), w2.data_ptr(), _ptr(b2), h.data_ptr(), F, _ptr(residual),
        gamma.data_ptr(), beta.data_ptr(), float(eps), M, Cc, _ptr(out_f32), _ptr(out_lp), _ptr(pos), _ptr(out_pos),
        _stream()))


SCORE_FUSED = _os.environ.get("MOYOLO_SCORE_FUSED", "1") != "0"


def linear_add_layernorm_scores(x: torch.Tensor, w: torch.Tensor, b, residual, gamma, beta, eps: float,
                                score_w: torch.Tensor, score_b: torch.Tensor, out_f32=None, out_lp=None, logits=None,
                                scores=None, labels=None) -> None:
    """linear_add_layernorm + the class-score head (logits, sigmoid(max), argmax) of the normalised rows in ONE
    launch; score_w fp32 [nc, 256], nc <= 8."""
    _cuda(x, w, b, residual, gamma, beta, score_w, score_b, out_f32, out_lp, logits, scores, labels)
    M, K = x.shape
    N = w.shape[0]
    nc = score_w.shape[0]
    if x.stride(1) != 1 or not w.is_contiguous() or not score_w.is_contiguous() or score_w.shape[1] != N:
        raise ValueError("linear_add_layernorm_scores: x columns / w / score_w must be contiguous, score_w [nc, N]")
    _lib.check(_lib.lib().moyolo_linear_add_layernorm_scores(
        x.data_ptr(), x.stride(0), w.data_ptr(), _ptr(b), _ptr(residual), gamma.data_ptr(), beta.data_ptr(), float(eps),
        M, N, K, _ptr(out_f32), _ptr(out_lp), score_w.data_ptr(), score_b.data_ptr(), nc, _ptr(logits), _ptr(scores),
        _ptr(labels), _stream()))


def self_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, row_offsets: torch.Tensor,
                   row_offsets_host: Sequence[int], n_heads: int, attn_mask: Optional[torch.Tensor] = None,
                   out: Optional[torch.Tensor] = None, seg_len: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(q k^T / sqrt(Dh)) v per (sequence, head); q/k/v are [R, C] column slices."""
    _cuda(q, k, v, row_offsets, attn_mask)
    R, Cc = q.shape
    if out is None:
        out = torch.empty(R, Cc, dtype=q.dtype, device=q.device)
    host = (C.c_int32 * len(row_offsets_host))(*[int(x) for x in row_offsets_host])
    _lib.check(_lib.lib().moyolo_self_attention(
        q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), out.data_ptr(),
        out.stride(0), _dt(q), len(row_offsets_host) - 1, row_offsets.data_ptr(), host, _ptr(seg_len), n_heads,
        Cc // n_heads, _ptr(attn_mask), _stream()))
    return out


def add_layernorm(x: torch.Tensor, residual: Optional[torch.Tensor], gamma: torch.Tensor, beta: torch.Tensor,
                  eps: float, want_f32: bool = True, want_lp: bool = False,
                  lp_dtype: Optional[torch.dtype] = None, pos: Optional[torch.Tensor] = None,
                  out_f32: Optional[torch.Tensor] = None, out_lp: Optional[torch.Tensor] = None,
                  out_pos: Optional[torch.Tensor] = None):
    """LayerNorm(x + residual); returns (out_f32 | None, out_lp | None, out_plus_pos_lp | None).
    Pre-allocated outputs may be passed (out_f32 may alias `residual`: every row is read before written)."""
    _cuda(x, residual, gamma, beta, pos)
    R, Cc = x.shape
    lp_dtype = lp_dtype or (out_lp.dtype if out_lp is not None else (out_pos.dtype if out_pos is not None
                                                                      else torch.float32))
    if out_f32 is None and want_f32:
        out_f32 = torch.empty(R, Cc, dtype=torch.float32, device=x.device)
    if out_lp is None and want_lp:
        out_lp = torch.empty(R, Cc, dtype=lp_dtype, device=x.device)
    if out_pos is None and pos is not None:
        out_pos = torch.empty(R, Cc, dtype=lp_dtype, device=x.device)
    _lib.check(_lib.lib().moyolo_add_layernorm(
        x.data_ptr(), _ptr(residual), gamma.data_ptr(), beta.data_ptr(), float(eps), R, Cc, _ptr(out_f32),
        _ptr(out_lp), _ptr(pos), _ptr(out_pos), _DT[lp_dtype], _stream()))
    return out_f32, out_lp, out_pos


def add_cast(a: torch.Tensor, b: Optional[torch.Tensor], dtype: torch.dtype,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(a, b)
    if out is None:
        out = torch.empty(a.shape, dtype=dtype, device=a.device)
    _lib.check(_lib.lib().moyolo_add_cast(a.data_ptr(), _ptr(b), out.data_ptr(), _DT[dtype], a.numel(), _stream()))
    return out


def box_refine(h: torch.Tensor, w3: torch.Tensor, b3: torch.Tensor, ref: torch.Tensor,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(h, w3, b3, ref, out)
    R, K = h.shape
    if out is None:
        out = torch.empty(R, 4, dtype=torch.float32, device=h.device)
    elif not out.is_contiguous() or out.dtype != torch.float32 or out.numel() != R * 4:
        raise ValueError("box_refine: out must be contiguous fp32 with R*4 elements")
    _lib.check(_lib.lib().moyolo_box_refine(h.data_ptr(), h.stride(0), _dt(h), w3.data_ptr(), b3.data_ptr(),
                                            ref.data_ptr(), out.data_ptr(), R, K, _stream()))
    return out


def score_head(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, want_scores: bool = True, out=None,
               max_logit: Optional[torch.Tensor] = None):
    """out: optional (logits [R,nc] f32, scores [R] f32, labels [R] i32) pre-allocated; max_logit [R] f32 optional."""
    _cuda(x, w, b)
    R, K = x.shape
    nc = w.shape[0]
    if out is not None:
        logits, scores, labels = out
    else:
        logits = torch.empty(R, nc, dtype=torch.float32, device=x.device)
        scores = torch.empty(R, dtype=torch.float32, device=x.device) if want_scores else None
        labels = torch.empty(R, dtype=torch.int32, device=x.device) if want_scores else None
    _lib.check(_lib.lib().moyolo_score_head(x.data_ptr(), x.stride(0), _dt(x), w.data_ptr(), b.data_ptr(),
                                            logits.data_ptr(), _ptr(scores), _ptr(labels), R, K, nc, _ptr(max_logit),
                                            _stream()))
    return logits, scores, labels


def sigmoid(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(x)
    x = x.contiguous()
    y = torch.empty_like(x) if out is None else out
    _lib.check(_lib.lib().moyolo_sigmoid(x.data_ptr(), y.data_ptr(), x.numel(), _stream()))
    return y


def inverse_sigmoid(x: torch.Tensor) -> torch.Tensor:
    _cuda(x)
    x = x.contiguous()
    y = torch.empty_like(x)
    _lib.check(_lib.lib().moyolo_inverse_sigmoid(x.data_ptr(), y.data_ptr(), x.numel(), _stream()))
    return y


def pos2posemb(pos: torch.Tensor, num_pos_feats: int = 64, temperature: float = 10000.0,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(pos)
    pos = pos.contiguous()
    n_coord = pos.shape[-1]
    rows = pos.numel() // n_coord
    emb = out if out is not None else torch.empty(*pos.shape[:-1], n_coord * num_pos_feats, dtype=torch.float32,
                                                  device=pos.device)
    _lib.check(_lib.lib().moyolo_pos2posemb(pos.data_ptr(), emb.data_ptr(), rows, n_coord, num_pos_feats,
                                            float(temperature), _stream()))
    return emb


def linear_k4_relu(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, out_dtype: torch.dtype) -> torch.Tensor:
    _cuda(x, w, b)
    R = x.shape[0]
    N = w.shape[0]
    y = torch.empty(R, N, dtype=out_dtype, device=x.device)
    _lib.check(_lib.lib().moyolo_linear_k4_relu(x.data_ptr(), w.data_ptr(), _ptr(b), y.data_ptr(), _DT[out_dtype],
                                                R, N, _stream()))
    return y


def track_workspace_bytes(n: int) -> int:
    return int(_lib.lib().moyolo_track_workspace_bytes(int(n)))


def track_assign(scores: torch.Tensor, boxes: torch.Tensor, obj_idxes: torch.Tensor, disappear_time: torch.Tensor,
                 counters: torch.Tensor, workspace: torch.Tensor, score_thresh: float = 0.4,
                 filter_thresh: float = 0.5, miss_tolerance: int = 5, iou_thresh: float = 0.8) -> None:
    """In-place RuntimeTrackerBase.update (head.py:1201-1283) on device."""
    _cuda(scores, boxes, obj_idxes, disappear_time, counters, workspace)
    n = scores.shape[0]
    assert obj_idxes.dtype == torch.int64 and disappear_time.dtype == torch.int64 and counters.dtype == torch.int64
    assert workspace.numel() * workspace.element_size() >= track_workspace_bytes(n)
    _lib.check(_lib.lib().moyolo_track_assign(
        scores.data_ptr(), boxes.data_ptr(), obj_idxes.data_ptr(), disappear_time.data_ptr(), counters.data_ptr(), n,
        float(score_thresh), float(filter_thresh), int(miss_tolerance), float(iou_thresh), workspace.data_ptr(),
        _stream()))


def track_compact(obj_idxes: torch.Tensor, fields: Sequence[torch.Tensor], outs: Sequence[torch.Tensor],
                  n_active: torch.Tensor, active_index: torch.Tensor) -> None:
    """Select rows with obj_idxes >= 0 (order preserved) from every field into `outs`."""
    _cuda(obj_idxes, n_active, active_index, *fields, *outs)
    n = obj_idxes.shape[0]
    nf = len(fields)
    src = (C.c_void_p * nf)(*[f.data_ptr() for f in fields])
    dst = (C.c_void_p * nf)(*[o.data_ptr() for o in outs])
    rb = (C.c_int64 * nf)(*[f.stride(0) * f.element_size() if f.dim() > 1 else f.element_size() for f in fields])
    _lib.check(_lib.lib().moyolo_track_compact(obj_idxes.data_ptr(), n, n_active.data_ptr(), active_index.data_ptr(),
                                               src, dst, rb, nf, _stream()))


def track_assign_batched(scores, boxes, ids, dis, counters, row_offsets, n_seq: int, max_rows_per_seq: int,
                         workspace, score_thresh=0.4, filter_thresh=0.5, miss_tolerance=5, iou_thresh=0.8,
                         ctrl=None) -> None:
    """RuntimeTrackerBase.update for every lock-step sequence in one launch (one CTA per sequence)."""
    _cuda(scores, boxes, ids, dis, counters, row_offsets, workspace)
    assert workspace.numel() * workspace.element_size() >= n_seq * track_workspace_bytes(max_rows_per_seq)
    _lib.check(_lib.lib().moyolo_track_assign_batched(
        scores.data_ptr(), boxes.data_ptr(), ids.data_ptr(), dis.data_ptr(), counters.data_ptr(),
        row_offsets.data_ptr(), n_seq, max_rows_per_seq, float(score_thresh), float(filter_thresh),
        int(miss_tolerance), float(iou_thresh), workspace.data_ptr(), _ptr(ctrl), _stream()))


def frame_assemble(n_seq, n_detect, C, cap, n_tracks, t_ref, t_qpos, t_label, t_ids, t_dis, class_embed, det_embed,
                   det_refer, x, refer_logit, pos, ids, dis, row_offsets, rows_pad, num_pos_feats=64,
                   temperature=10000.0, ctrl=None, refer_sig=None, x_lp=None, xq_lp=None) -> None:
    """Optional fused outputs: refer_sig = sigmoid(refer_logit), x_lp = x and xq_lp = x + pos as GEMM operands."""
    lp = x_lp if x_lp is not None else xq_lp
    _lib.check(_lib.lib().moyolo_frame_assemble(
        n_seq, n_detect, C, cap, n_tracks.data_ptr(), t_ref.data_ptr(), t_qpos.data_ptr(), t_label.data_ptr(),
        t_ids.data_ptr(), t_dis.data_ptr(), class_embed.data_ptr(), det_embed.data_ptr(), det_refer.data_ptr(),
        x.data_ptr(), refer_logit.data_ptr(), pos.data_ptr(), ids.data_ptr(), dis.data_ptr(), row_offsets.data_ptr(),
        rows_pad, num_pos_feats, float(temperature), _ptr(ctrl), _ptr(refer_sig), _ptr(x_lp), _ptr(xq_lp),
        _dt(lp) if lp is not None else F32, _stream()))


def frame_compact(n_seq, C, cap, row_offsets, ids, dis, labels, refer_logit, pos, hs, boxes, n_active, active_index,
                  c_ref, c_pos, c_hs, c_box, t_label, t_ids, t_dis, ctrl=None, q_qk_lp=None, q_tgt_lp=None,
                  num_pos_feats=64, temperature=10000.0) -> None:
    """Optional fused QIM operands: q_qk_lp = c_hs + pos2posemb(c_ref), q_tgt_lp = c_hs (qim.py:255,271)."""
    lp = q_qk_lp if q_qk_lp is not None else q_tgt_lp
    _lib.check(_lib.lib().moyolo_frame_compact(
        n_seq, C, cap, row_offsets.data_ptr(), ids.data_ptr(), dis.data_ptr(), labels.data_ptr(),
        refer_logit.data_ptr(), pos.data_ptr(), hs.data_ptr(), boxes.data_ptr(), n_active.data_ptr(),
        active_index.data_ptr(), c_ref.data_ptr(), c_pos.data_ptr(), c_hs.data_ptr(), c_box.data_ptr(),
        t_label.data_ptr(), t_ids.data_ptr(), t_dis.data_ptr(), _ptr(ctrl), _ptr(q_qk_lp), _ptr(q_tgt_lp),
        _dt(lp) if lp is not None else F32, num_pos_feats, float(temperature), _stream()))


def frame_assign_compact(n_seq, C, cap, rows_pad, row_offsets, scores, ids_in, dis_in, counters, ids_out, dis_out,
                         labels, refer_logit, pos, hs, boxes, n_active, active_index, c_ref, c_pos, c_hs, c_box,
                         t_label, t_ids, t_dis, score_thresh=0.4, filter_thresh=0.5, miss_tolerance=5, ctrl=None,
                         q_qk_lp=None, q_tgt_lp=None, num_pos_feats=64, temperature=10000.0) -> None:
    """ID assignment (head.py:1232-1243) + active-track compaction in one launch; `counters` is only read
    (follow with track_suppress_batched)."""
    lp = q_qk_lp if q_qk_lp is not None else q_tgt_lp
    _lib.check(_lib.lib().moyolo_frame_assign_compact(
        n_seq, C, cap, rows_pad, row_offsets.data_ptr(), scores.data_ptr(), ids_in.data_ptr(), dis_in.data_ptr(),
        counters.data_ptr(), float(score_thresh), float(filter_thresh), int(miss_tolerance), ids_out.data_ptr(),
        dis_out.data_ptr(), labels.data_ptr(), refer_logit.data_ptr(), pos.data_ptr(), hs.data_ptr(), _ptr(boxes),
        n_active.data_ptr(), active_index.data_ptr(), c_ref.data_ptr(), c_pos.data_ptr(), c_hs.data_ptr(),
        _ptr(c_box), t_label.data_ptr(), t_ids.data_ptr(), t_dis.data_ptr(), _ptr(ctrl), _ptr(q_qk_lp),
        _ptr(q_tgt_lp), _dt(lp) if lp is not None else F32, num_pos_feats, float(temperature), _stream()))


def track_suppress_batched(boxes, ids, counters, row_offsets, n_seq: int, max_rows_per_seq: int, workspace,
                           iou_thresh=0.8, ctrl=None) -> None:
    """Duplicate filter + renumbering side effects on `counters` (head.py:1155-1196, 1268-1282) for ids that
    frame_assign_compact already updated."""
    _cuda(boxes, ids, counters, row_offsets, workspace)
    assert workspace.numel() * workspace.element_size() >= n_seq * track_workspace_bytes(max_rows_per_seq)
    _lib.check(_lib.lib().moyolo_track_suppress_batched(
        boxes.data_ptr(), ids.data_ptr(), counters.data_ptr(), row_offsets.data_ptr(), n_seq, max_rows_per_seq,
        float(iou_thresh), workspace.data_ptr(), _ptr(ctrl), _stream()))


def frame_writeback(n_seq, C, cap, row_offsets, n_active, new_qpos, c_box, t_qpos, t_ref, n_tracks, ctrl=None,
                    info=None, boxes=None, active_index=None) -> None:
    """c_box None: the boxes of the active rows are gathered from `boxes` through `active_index`."""
    _lib.check(_lib.lib().moyolo_frame_writeback(
        n_seq, C, cap, row_offsets.data_ptr(), n_active.data_ptr(), new_qpos.data_ptr(), _ptr(c_box),
        t_qpos.data_ptr(), t_ref.data_ptr(), n_tracks.data_ptr(), _ptr(ctrl), _ptr(info), _ptr(boxes),
        _ptr(active_index), _stream()))


def frame_emit(n_seq, rows_pad, row_offsets, ids, boxes, scores, labels, n_active, active_index, seq_ids, frame_rows,
               table, ctrl) -> None:
    """Per-frame packed result rows + append of the tracked objects to the device track table."""
    _lib.check(_lib.lib().moyolo_frame_emit(
        n_seq, rows_pad, row_offsets.data_ptr(), ids.data_ptr(), boxes.data_ptr(), scores.data_ptr(),
        labels.data_ptr(), n_active.data_ptr(), active_index.data_ptr(), seq_ids.data_ptr(), frame_rows.data_ptr(),
        table.data_ptr(), table.shape[0], ctrl.data_ptr(), _stream()))


# ---- encoder-side query selection (head.py:993-1113) ----------------------------------------------------------
def enc_output_scores(x: torch.Tensor, w: torch.Tensor, b, gamma, beta, eps: float, zero_in_rows, score_w, score_b,
                      out_f32=None, out_lp=None, logits=None, max_logit=None) -> None:
    """LayerNorm(Linear(mask * x)) + class-score head over all rows in one tcgen05 kernel (bf16 operands)."""
    _cuda(x, w, b, gamma, beta, zero_in_rows, score_w, score_b, out_f32, out_lp, logits, max_logit)
    M = x.shape[0]
    nc = 0 if score_w is None else score_w.shape[0]
    _lib.check(_lib.lib().moyolo_enc_output_scores(
        x.data_ptr(), x.stride(0), w.data_ptr(), _ptr(b), gamma.data_ptr(), beta.data_ptr(), float(eps),
        _ptr(zero_in_rows), _ptr(score_w), _ptr(score_b), nc, M, _ptr(out_f32), _ptr(out_lp), _ptr(logits),
        _ptr(max_logit), _stream()))


def topk(scores: torch.Tensor, k: int, out: Optional[torch.Tensor] = None, vals: Optional[torch.Tensor] = None):
    """Indices (int32 [batch, k]) of the k largest entries of every row of scores [batch, n] fp32, ordered as
    torch.topk(sorted=True) (descending; ties by ascending index)."""
    _cuda(scores, out, vals)
    B, n = scores.shape
    if scores.stride(1) != 1 or scores.dtype != torch.float32:
        raise ValueError("topk: scores must be fp32 with contiguous columns")
    if out is None:
        out = torch.empty(B, k, dtype=torch.int32, device=scores.device)
    _lib.check(_lib.lib().moyolo_topk(scores.data_ptr(), scores.stride(0), n, B, int(k), out.data_ptr(), _ptr(vals),
                                      _stream()))
    return out


def select_gather(features: torch.Tensor, logits, idx: torch.Tensor, embed: torch.Tensor, embed_lp=None,
                  enc_scores=None) -> None:
    _cuda(features, logits, idx, embed, embed_lp, enc_scores)
    B, Lv, Cc = features.shape
    k = idx.shape[1]
    nc = 0 if logits is None else logits.shape[-1]
    _lib.check(_lib.lib().moyolo_select_gather(
        features.data_ptr(), _ptr(logits), idx.data_ptr(), B, k, Lv, Cc, nc, embed.data_ptr(), _ptr(embed_lp),
        _dt(embed_lp) if embed_lp is not None else F32, _ptr(enc_scores), _stream()))


def anchor_box(h: torch.Tensor, w3: torch.Tensor, b3: torch.Tensor, idx: torch.Tensor, shapes, len_v: int,
               out: torch.Tensor, grid_size: float = 0.05, eps: float = 1e-2) -> torch.Tensor:
    _cuda(h, w3, b3, idx, out)
    R, K = h.shape
    arr, L = _shapes_arr(shapes)
    _lib.check(_lib.lib().moyolo_anchor_box(h.data_ptr(), h.stride(0), _dt(h), w3.data_ptr(), b3.data_ptr(),
                                            idx.data_ptr(), arr, L, len_v, float(grid_size), float(eps), out.data_ptr(),
                                            R, K, _stream()))
    return out


def anchor_invalid(shapes, len_v: int, device, grid_size: float = 0.05, eps: float = 1e-2) -> torch.Tensor:
    """uint8 [len_v]: 1 where the anchor of a pyramid position is masked out (head.py:1006)."""
    arr, L = _shapes_arr(shapes)
    out = torch.empty(len_v, dtype=torch.uint8, device=device)
    _lib.check(_lib.lib().moyolo_anchor_invalid(arr, L, len_v, float(grid_size), float(eps), out.data_ptr(), _stream()))
    return out


def mask_rows(x: torch.Tensor, zero_rows: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[r] = 0 where zero_rows[r % len(zero_rows)] else x[r]; x fp32 [R, C] contiguous."""
    _cuda(x, zero_rows, out)
    R, Cc = x.shape
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.lib().moyolo_mask_rows(x.data_ptr(), zero_rows.data_ptr(), zero_rows.numel(), out.data_ptr(), R, Cc,
                                           _stream()))
    return out
