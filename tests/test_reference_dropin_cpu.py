"""CPU-box check of the drop-in boundary against the LIVE reference tree (skipped where /root/reference does not
exist, i.e. on the GPU box): INTEGRATION.md 2a / 2b.

  * 2a: with `moyolo_b200.msda_ext.install()` the unmodified reference imports (`import ultralytics` needs the
    pybind module `MultiScaleDeformableAttention`, MOTR/models/ops/functions/ms_deform_attn_func.py:21) and its
    `MSDeformAttnFunction` is bound to this library's entry points.
  * 2b: the reference's own `MYDecoder` (ultralytics/nn/modules/head.py:807-863) built with the drop-in classes
    swapped in for `.transformer`'s (head.py:843-845) has EXACTLY the reference's state_dict (keys, shapes,
    dtypes), loads a reference checkpoint strictly, and calls the drop-in forward with the reference's arguments.
Runs in a subprocess: importing the full reference registers stub packages in sys.modules.
"""
import subprocess
import sys
import textwrap
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.skipif(not Path("/root/reference/ultralytics/nn/modules/head.py").exists(),
                                reason="reference tree not present (GPU box)")

SCRIPT = textwrap.dedent('''
    import sys, importlib
    sys.path.insert(0, %r)
    import torch
    import moyolo_b200 as m
    from moyolo_b200 import msda_ext
    from oracle import ref_loader

    # ---- 2a: the reference imports with this library registered as its native extension
    msda_ext.install()
    head, qim, structures = ref_loader.load_full_reference(msda_module=sys.modules["MultiScaleDeformableAttention"])
    f = importlib.import_module("MOTR.models.ops.functions.ms_deform_attn_func")
    assert f.MSDA is msda_ext, "the reference did not bind to moyolo_b200.msda_ext"
    assert f.MSDA.ms_deform_attn_forward is msda_ext.ms_deform_attn_forward
    assert f.MSDA.ms_deform_attn_backward is msda_ext.ms_deform_attn_backward
    import ultralytics
    assert "ultralytics.nn.modules.head" in sys.modules
    # the reference's CPU guard is reproduced by the stand-in (MOTR/models/ops/src/ms_deform_attn.h:36)
    v = torch.zeros(1, 30, 2, 2)
    shapes = torch.tensor([[6, 4], [3, 2]])
    try:
        f.MSDeformAttnFunction.apply(v, shapes, torch.tensor([0, 24]), torch.zeros(1, 2, 2, 2, 2, 2), torch.zeros(1, 2, 2, 2, 2), 2)
        raise SystemExit("CPU call did not raise")
    except RuntimeError as e:
        assert "Not implemented on the CPU" in str(e), e

    # ---- 2b: MYDecoder with the drop-in classes has the reference's checkpoint layout
    torch.manual_seed(0)
    ref_dec = head.MYDecoder(nc=5, ch=(256, 512, 512))
    ref_sd = ref_dec.state_dict()
    T = importlib.import_module("ultralytics.nn.modules.transformer")
    names = ("MLP", "DeformableTransformerDecoder", "DeformableTransformerDecoderLayer", "MSDeformAttn",
             "MOTRDecoderLayer", "MOTRMSDeformAttn", "MOTRTransformerDecoder")
    saved = {n: getattr(head, n) for n in names if hasattr(head, n)}
    for n in saved:
        setattr(head, n, getattr(m, n))          # the one-line import change of INTEGRATION.md 2b
    try:
        new_dec = head.MYDecoder(nc=5, ch=(256, 512, 512))
    finally:
        for n, c in saved.items():
            setattr(head, n, c)
    assert type(new_dec.decoder).__module__.startswith("moyolo_b200"), type(new_dec.decoder)
    assert type(new_dec.decoder.layers[0]).__module__.startswith("moyolo_b200")
    assert type(new_dec.decoder.layers[0].cross_attn).__module__.startswith("moyolo_b200")
    new_sd = new_dec.state_dict()
    assert list(new_sd.keys()) == list(ref_sd.keys()), set(new_sd) ^ set(ref_sd)
    for k in ref_sd:
        assert new_sd[k].shape == ref_sd[k].shape and new_sd[k].dtype == ref_sd[k].dtype, k
    missing, unexpected = new_dec.load_state_dict(ref_sd, strict=True)
    assert not missing and not unexpected
    for k in ref_sd:
        assert torch.equal(new_dec.state_dict()[k], ref_sd[k]), k
    # same initialisation scheme for the attention module (transformer.py:221-237)
    a, b = T.MSDeformAttn(256, 3, 8, 4), m.MSDeformAttn(256, 3, 8, 4)
    assert torch.equal(a.sampling_offsets.bias, b.sampling_offsets.bias)
    assert float(b.sampling_offsets.weight.abs().max()) == 0 and float(b.attention_weights.weight.abs().max()) == 0
    # MYDecoder.forward reaches the drop-in decoder with the reference's call (head.py:938): on a CPU box the library
    # refuses CPU tensors with the reference extension's wording instead of silently computing elsewhere
    new_dec.eval()
    x = [torch.randn(1, c, h, w) for c, (h, w) in zip((256, 512, 512), ((8, 8), (4, 4), (2, 2)))]
    try:
        with torch.no_grad():
            new_dec(x)
        raise SystemExit("drop-in decoder ran on CPU tensors")
    except RuntimeError as e:
        assert "Not implemented on the CPU" in str(e), e
    print("DROPIN_OK", len(ref_sd))
''')


@pytest.mark.timeout(600)
def test_reference_imports_and_loads_with_dropin_classes():
    r = subprocess.run([sys.executable, "-c", SCRIPT % str(ROOT)], capture_output=True, text=True, cwd=str(ROOT), timeout=580)
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
