"""GPU: encoder-side query selection against the goldens minted from the reference MYDecoder
(head.py:993-1113) and the top-k kernel against torch.topk."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_rms
from moyolo_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("B,n,k", [(1, 13566, 300), (4, 8400, 300), (2, 252, 48), (3, 22050, 1024), (1, 300, 300),
                                   (2, 5000, 1)])
def test_topk_matches_torch(dev, B, n, k):
    from moyolo_b200 import ops
    g = torch.Generator().manual_seed(n + k)
    s = torch.randn(B, n, generator=g)
    s[:, ::7] = s[:, 3:4]            # many exact ties, some of them straddling the k-th value
    s[0, :5] = torch.tensor([float("inf"), -float("inf"), 0.0, -0.0, 1e-30])[: min(5, n)]
    sd = s.to(dev)
    vals = torch.empty(B, k, device=dev)
    idx = ops.topk(sd, k, vals=vals).cpu().long()
    ref = torch.sort(s, dim=1, descending=True, stable=True)   # ties: ascending index == our contract
    assert torch.equal(idx, ref.indices[:, :k])
    assert torch.equal(vals.cpu(), ref.values[:, :k])
    # same multiset of values as torch.topk (whose tie order is unspecified)
    tk = torch.topk(s, k, dim=1)
    assert torch.equal(torch.sort(vals.cpu(), dim=1).values, torch.sort(tk.values, dim=1).values)


@pytest.mark.parametrize("name", ["select_tiny", "select_c1", "select_kitti_nc5"])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_query_selection_vs_reference_golden(dev, name, precision, tol):
    """fp32: max|a-b| <= 1e-4*rms per tensor and the SAME selected positions in the same order, except where two
    reference scores are closer than 1e-5 (a swap there is within the fp32 tolerance). bf16 (bf16 GEMM operands):
    2e-2*rms on feats and on embeddings/boxes of the positions both selections share; >= 90 % shared positions."""
    from moyolo_b200 import ops
    from moyolo_b200.selector import QuerySelector
    from oracle import torch_port as tp
    meta, g = load_golden(name)
    spec = syn.DecoderSpec(nc=meta["nc"])
    shapes, B, nq = meta["shapes"], meta["B"], meta["nq"]
    ch = (256, 512, 512)
    sd = syn.make_selector_state(spec, ch, meta["weight_seed"])
    maps = syn.make_pyramid_maps(meta["seed"], B, shapes, ch)
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    sel = QuerySelector(sd, spec, shapes, ch, dev, precision, nq, B)
    maps_cl = [m.permute(0, 2, 3, 1).contiguous().to(dev).to(dt) for m in maps]
    Lv = syn.level_sizes(shapes)
    feats = torch.zeros(B, Lv, 256, dtype=dt, device=dev)
    embed = torch.zeros(B, nq, 256, device=dev)
    refer = torch.zeros(B, nq, 4, device=dev)
    sel.run(maps_cl, feats, embed, refer)
    torch.cuda.synchronize()
    step = meta["feats_row_step"]
    # feats is STORED in bf16 in bf16 mode: output rounding alone is 2^-9*|x| with |x| up to ~5 rms over 10^6 elements,
    # on top of the operand rounding -> 3e-2*rms for this tensor (fp32-stored tensors below keep 2e-2)
    assert rel_rms(feats[:, ::step].float().cpu().numpy(), g["feats"]) < (tol if precision == "fp32" else 3e-2)
    # full oracle (pinned to the same golden by tests/test_oracle_vs_golden.py) for the per-position quantities
    with torch.no_grad():
        ofeats, _ = tp.encoder_input(sd, maps)
        o = tp.query_selection(sd, ofeats, shapes, nq)
    idx = sel.idx.cpu().long()
    ref_idx = o["topk"]
    ref_scores = o["scores"].max(-1).values
    if precision == "fp32":
        for b in range(B):
            rs = ref_scores[b, ref_idx[b]]
            gap = torch.minimum(torch.cat([rs[:1] * 0 + 1, (rs[:-1] - rs[1:]).abs()]),
                                torch.cat([(rs[:-1] - rs[1:]).abs(), rs[:1] * 0 + 1]))
            clear = gap > 1e-5
            assert torch.equal(idx[b][clear], ref_idx[b][clear]), f"selection order differs in batch {b}"
    else:
        for b in range(B):
            shared = len(set(idx[b].tolist()) & set(ref_idx[b].tolist()))
            assert shared >= 0.9 * nq, f"only {shared}/{nq} shared positions"
    # embeddings / logits / boxes at the positions OUR selection picked, against the oracle's per-position values
    bi = torch.arange(B).unsqueeze(-1).repeat(1, nq)
    assert rel_rms(embed.cpu().numpy(), o["features"][bi, idx].numpy()) < tol
    assert rel_rms(sel.enc_scores.cpu().numpy(), o["scores"][bi, idx].numpy()) < tol
    anchors, _ = tp.generate_anchors(shapes)
    obox = (tp.mlp_forward(sd, o["features"], 3, "enc_bbox_head.") + anchors)[bi, idx].numpy()
    got = refer.cpu().numpy()
    fin = np.isfinite(obox)
    assert np.array_equal(fin, np.isfinite(got))
    assert rel_rms(got[fin], obox[fin]) < tol


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_frames_from_neck_maps_equal_frames_from_selected_queries(dev, precision):
    """TrackEngine with the query selection inside the frame graph (inputs = neck maps) must produce exactly the
    tracks of the decoder-only engine fed with the standalone selector's outputs (same kernels, same order)."""
    from moyolo_b200.selector import QuerySelector
    from moyolo_b200.tracker import TrackEngine
    spec = syn.DecoderSpec()
    shapes = [list(s) for s in syn.PYRAMIDS["tiny"]]
    ch, nd, S, n_frames = (256, 512, 512), 48, 2, 5
    sd = dict(syn.make_decoder_state(spec, 7))
    sd.update(syn.make_selector_state(spec, ch, 0))
    sd[f"dec_score_head.{spec.n_layers - 1}.bias"] = sd[f"dec_score_head.{spec.n_layers - 1}.bias"] + 3.0  # some tracks are born
    dt = torch.bfloat16 if precision == "bf16" else torch.float32
    Lv = syn.level_sizes(shapes)
    sel_a = QuerySelector(sd, spec, shapes, ch, dev, precision, nd, S)
    sel_b = QuerySelector(sd, spec, shapes, ch, dev, precision, nd, S)
    eng_maps = TrackEngine(sd, spec, shapes, dev, precision, nd, S, selector=sel_a)
    eng_dec = TrackEngine(sd, spec, shapes, dev, precision, nd, S)
    base = [m.permute(0, 2, 3, 1).contiguous() for m in syn.make_pyramid_maps(5, S, shapes, ch)]
    g = torch.Generator().manual_seed(11)
    born = 0
    for t in range(n_frames):
        maps = [(m + 0.05 * t * torch.randn(m.shape, generator=g)).to(dev).to(dt).contiguous() for m in base]
        feats = torch.zeros(S, Lv, 256, dtype=dt, device=dev)
        de, dr = torch.zeros(S, nd, 256, device=dev), torch.zeros(S, nd, 4, device=dev)
        sel_b.run(maps, feats, de, dr)
        a = eng_maps.step(*maps)
        a = [{k: v.clone() for k, v in o.items()} for o in a]
        b = eng_dec.step(feats, de, dr)
        for s in range(S):
            for k in ("ids", "boxes", "scores", "labels"):
                assert torch.equal(a[s][k], b[s][k]), (t, s, k)
            born += int((a[s]["ids"] >= 0).sum())
    assert born > 0, "no object was ever tracked"
    assert torch.equal(eng_maps.track_table(), eng_dec.track_table())


def test_selection_ahead_equals_in_graph_pipelined(dev):
    """Selection-ahead mode (input projection + value projection + query selection as their own graph on a side
    stream, overlapping the previous frame) against the in-graph mode over a pipelined submit/collect sequence:
    bit-identical rows and track table."""
    from moyolo_b200.selector import QuerySelector
    from moyolo_b200.tracker import TrackEngine
    spec = syn.DecoderSpec()
    shapes = [list(s) for s in syn.PYRAMIDS["tiny"]]
    ch, nd, S, n_frames = (256, 512, 512), 48, 2, 8
    sd = dict(syn.make_decoder_state(spec, 7))
    sd.update(syn.make_selector_state(spec, ch, 0))
    sd[f"dec_score_head.{spec.n_layers - 1}.bias"] = sd[f"dec_score_head.{spec.n_layers - 1}.bias"] + 3.0
    base = [m.permute(0, 2, 3, 1).contiguous() for m in syn.make_pyramid_maps(5, S, shapes, ch)]
    g = torch.Generator().manual_seed(11)
    frames = [[(m + 0.05 * t * torch.randn(m.shape, generator=g)).to(dev).to(torch.bfloat16).contiguous() for m in base]
              for t in range(n_frames)]
    out = {}
    for ahead in (False, True, "abort"):   # "abort": ahead mode with mis-speculated padded sizes (device-side abort + re-launch)
        sel = QuerySelector(sd, spec, shapes, ch, dev, "bf16", nd, S)
        kw = dict(margin=0, bucket=8) if ahead == "abort" else {}
        eng = TrackEngine(sd, spec, shapes, dev, "bf16", nd, S, selector=sel, value_ahead=bool(ahead), **kw)
        assert eng._sel_ahead == bool(ahead)
        got = {}
        for t in range(n_frames):
            eng.submit(*frames[t], want_rows=True)
            if t > 0:
                got[t - 1] = [{k: v.clone() for k, v in o.items()} for o in eng.collect(t - 1)]
        got[n_frames - 1] = [{k: v.clone() for k, v in o.items()} for o in eng.collect(n_frames - 1)]
        out[ahead] = (got, eng.track_table().clone().cpu(), eng.n_tracks_host())
        if ahead == "abort":
            assert eng.aborts > 0, "the abort / re-launch path was not exercised"
    a = out[False]
    for key in (True, "abort"):
        b = out[key]
        assert a[2] == b[2] and max(a[2]) > 0
        assert torch.equal(a[1], b[1]), key
        for t in range(n_frames):
            for s in range(S):
                for k in ("ids", "boxes", "scores", "labels"):
                    assert torch.equal(a[0][t][s][k], b[0][t][s][k]), (key, t, s, k)
