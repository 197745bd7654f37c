"""moyolo_b200 — B200 (sm_100a) decoder hot path of DecoderTracker / MO-YOLO.

Public surface: the reference's operator API (`MSDeformAttn`, decoder layers/decoders, `MLP`,
`pos2posemb`), the legacy extension module (`msda_ext`), the device-side tracker (`tracker`) and
the sequence-sharding launcher (`sharding`). All compute goes through libmoyolo_b200.so.
"""
from .executor import get_default_precision, set_default_precision, set_gemm_engine  # noqa: F401
from .modules import (MLP, DeformableTransformerDecoder, DeformableTransformerDecoderLayer,  # noqa: F401
                      MOTRDecoderLayer, MOTRMSDeformAttn, MOTRTransformerDecoder, MSDeformAttn,
                      inverse_sigmoid, multi_scale_deformable_attn, pos2posemb)

__version__ = "0.1.0"
