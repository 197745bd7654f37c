"""CPU: the C-ABI shared library builds/loads without a GPU and exports every declared symbol."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def built_lib():
    from moyolo_b200 import build
    return build.build()


def declared_symbols():
    text = (ROOT / "include" / "moyolo_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(moyolo_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/moyolo_b200.h but not exported"


def test_ctypes_signatures_cover_header(built_lib):
    from moyolo_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.lib()
    assert lib.moyolo_version() == 100
    assert isinstance(_lib.last_error(), str)


def test_no_cpu_fallback():
    """Ops refuse CPU tensors with the reference extension's wording (ms_deform_attn.h:36)."""
    import torch
    import moyolo_b200 as m
    attn = m.MSDeformAttn(256, 3, 8, 4)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        attn(torch.zeros(1, 3, 256), torch.zeros(1, 3, 1, 4), torch.zeros(1, 21, 256), [[4, 4], [2, 2], [1, 1]])
    with pytest.raises(ValueError, match="must be 2 or 4"):
        attn(torch.zeros(1, 3, 256), torch.zeros(1, 3, 1, 3), torch.zeros(1, 21, 256), [[4, 4], [2, 2], [1, 1]])
    with pytest.raises(AssertionError):
        attn(torch.zeros(1, 3, 256), torch.zeros(1, 3, 1, 4), torch.zeros(1, 20, 256), [[4, 4], [2, 2], [1, 1]])
    with pytest.raises(ValueError, match="divisible"):
        m.MSDeformAttn(250, 3, 8, 4)


def test_state_dict_keys_match_reference_contract():
    """Checkpoint-key contract of SURVEY.md §8(b)."""
    import torch
    import moyolo_b200 as m
    layer = m.MOTRDecoderLayer(256, 8, 1024, 0., torch.nn.ReLU(), 3, 4)
    dec = m.MOTRTransformerDecoder(256, layer, 6)
    keys = dict(dec.state_dict())
    expect = {"self_attn.in_proj_weight": (768, 256), "self_attn.in_proj_bias": (768,),
              "self_attn.out_proj.weight": (256, 256), "norm1.weight": (256,),
              "cross_attn.sampling_offsets.weight": (192, 256), "cross_attn.sampling_offsets.bias": (192,),
              "cross_attn.attention_weights.weight": (96, 256), "cross_attn.value_proj.weight": (256, 256),
              "cross_attn.output_proj.weight": (256, 256), "norm2.bias": (256,), "linear1.weight": (1024, 256),
              "linear2.weight": (256, 1024), "norm3.weight": (256,)}
    for i in range(6):
        for k, shp in expect.items():
            assert tuple(keys[f"layers.{i}.{k}"].shape) == shp
    from moyolo_b200 import synthetic as syn
    sd = syn.make_decoder_state(syn.DecoderSpec(), 0)
    dec.load_state_dict({k: v for k, v in sd.items() if k.startswith("layers.")})  # strict
