"""Stand-in for the reference's pybind extension `MultiScaleDeformableAttention`.

The reference cannot even be imported without that module (ultralytics/utils/ops.py:886 ->
MOTR/models/ops/functions/ms_deform_attn_func.py:21 `import MultiScaleDeformableAttention as MSDA`).
`install()` registers this module under that name so the unmodified reference imports and its
`MSDeformAttnFunction.forward` (ms_deform_attn_func.py:24-31) lands in libmoyolo_b200.

Signatures follow MOTR/models/ops/src/vision.cpp:13-16 and ms_deform_attn_cuda.cu:20-80.
"""
from __future__ import annotations

import sys

import torch

from . import ops


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    """value [N,S,M,D], spatial_shapes [L,2] int64, level_start_index [L] int64,
    sampling_loc [N,Lq,M,L,P,2], attn_weight [N,Lq,M,L,P] -> [N,Lq,M*D]."""
    for name, t in (("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                    ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)):
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")  # ms_deform_attn_cuda.cu:28-32
        if not t.is_cuda:
            raise RuntimeError("Not implemented on the CPU" if name == "value" else
                               f"{name} must be a CUDA tensor")  # ms_deform_attn.h:36, .cu:34-38
    batch = value.shape[0]
    step = min(batch, int(im2col_step))
    if step <= 0 or batch % step != 0:
        raise RuntimeError(f"batch({batch}) must divide im2col_step({step})")  # .cu:50-52
    shapes = spatial_shapes.tolist()  # legacy FFI carries shapes on device: one small D2H here
    starts = level_start_index.tolist()
    acc = 0
    for (h, w), s in zip(shapes, starts):
        if s != acc:
            raise ValueError("level_start_index is not the exclusive cumsum of H*W")
        acc += h * w
    return ops.msda_sampled(value, shapes, sampling_loc, attn_weight)


def _check_legacy(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, extra=()):
    for name, t in (("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
                    ("sampling_loc", sampling_loc), ("attn_weight", attn_weight), *extra):
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")  # ms_deform_attn_cuda.cu:28-32, 92-97
        if not t.is_cuda:
            raise RuntimeError("Not implemented on the CPU" if name == "value" else
                               f"{name} must be a CUDA tensor")  # ms_deform_attn.h:36,57, .cu:34-38, 99-104


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step):
    """-> [grad_value, grad_sampling_loc, grad_attn_weight], each of its input's shape and dtype
    (MOTR/models/ops/src/cuda/ms_deform_attn_cuda.cu:83-153)."""
    _check_legacy(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, (("grad_output", grad_output),))
    batch = value.shape[0]
    step = min(batch, int(im2col_step))
    if step <= 0 or batch % step != 0:
        raise RuntimeError(f"batch({batch}) must divide im2col_step({step})")  # .cu:116-118
    gv, gl, gw = ops.msda_sampled_backward(value, spatial_shapes.tolist(), sampling_loc, attn_weight, grad_output)
    return [gv.to(value.dtype), gl.to(sampling_loc.dtype), gw.to(attn_weight.dtype)]


class MSDeformAttnFunction(torch.autograd.Function):
    """Same call signature as the reference's autograd function
    (MOTR/models/ops/functions/ms_deform_attn_func.py:24-41), forward and backward in libmoyolo_b200."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        ctx.im2col_step = im2col_step
        output = ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                        attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        value, shapes, start, loc, aw = ctx.saved_tensors
        gv, gl, gw = ms_deform_attn_backward(value, shapes, start, loc, aw, grad_output.contiguous(), ctx.im2col_step)
        return gv, None, None, gl, gw, None


def install(name: str = "MultiScaleDeformableAttention") -> None:
    """Register this module as `MultiScaleDeformableAttention` for the reference's import."""
    sys.modules[name] = sys.modules[__name__]
