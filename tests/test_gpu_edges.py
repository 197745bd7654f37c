"""GPU: edge cases of the hot path through the C ABI — empty and ragged inputs, the largest sweep configuration,
points entirely outside the maps, error behaviour (same exception types as the reference raises), per-frame reset
(the reference's SHIPPED semantics, `is_first` every frame -> IDs restart at 0) and the result
writers fed from the engine's device track table."""
import numpy as np
import pytest
import torch

from conftest import rel_rms

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _fused_inputs(syn, R, H, L, P, B, Lv, C, seed, dev, dt=torch.float32):
    g = torch.Generator().manual_seed(seed)
    value = torch.randn(B, Lv, C, generator=g).to(dev, dt)
    off = torch.randn(R, H * L * P * 2, generator=g).to(dev)
    lg = torch.randn(R, H * L * P, generator=g).to(dev)
    refer = torch.cat([torch.rand(R, 1, 2, generator=g), torch.rand(R, 1, 2, generator=g) * 0.4 + 0.02], -1).to(dev)
    return value, off, lg, refer


def test_largest_sweep_point_vs_c_oracle(dev):
    """BASELINE.json configs[4] corner: Q=1200, L=4, P=8, B=2 (L*P = 32 points per lane group, 640x640 pyramid +
    a 10x10 level), fp32: pre-normalised gather and its backward against the plain-C restatements."""
    from moyolo_b200 import ops, synthetic as syn
    from oracle import c_core
    shapes = [list(s) for s in syn.PYRAMIDS["C1"]] + [[10, 10]]
    value, loc, w, go = syn.make_core_grad_inputs(77, 2, 1200, 8, 32, shapes, 8)
    out = ops.msda_sampled(value.to(dev), shapes, loc.to(dev), w.to(dev))
    ref = c_core.msda_core(value.numpy(), shapes, loc.numpy(), w.numpy())
    assert rel_rms(out.cpu().numpy(), ref) < FP32_TOL
    gv, gl, gw = ops.msda_sampled_backward(value.to(dev), shapes, loc.to(dev), w.to(dev), go.to(dev))
    rv, rl, rw = c_core.msda_core_backward(value.numpy(), shapes, loc.numpy(), w.numpy(), go.numpy())
    for a, r in ((gv, rv), (gl, rl), (gw, rw)):
        assert rel_rms(a.cpu().numpy(), r) < FP32_TOL


def test_empty_and_ragged_rows(dev):
    """R = 0 for every gather entry point; a ragged batch whose middle sequence has no rows equals the dense
    computation of the non-empty sequences."""
    from moyolo_b200 import ops, synthetic as syn
    H, L, P, C = 8, 3, 4, 256
    shapes = [list(s) for s in syn.PYRAMIDS["tiny"]]
    Lv = syn.level_sizes(shapes)
    value, off, lg, refer = _fused_inputs(syn, 0, H, L, P, 3, Lv, C, 1, dev)
    assert ops.msda_fused(value, shapes, off, lg, refer, H, P, 3).shape == (0, C)
    vb = value.to(torch.bfloat16)
    w = torch.zeros(H * L * P * 3, C, dtype=torch.bfloat16, device=dev)
    assert ops.msda_proj_fused(vb, shapes, torch.zeros(0, C, dtype=torch.bfloat16, device=dev), w,
                               torch.zeros(H * L * P * 3, device=dev), refer, H, P, 3).shape == (0, C)
    v4 = value.view(3, Lv, H, 32)
    gv, gl, gw = ops.msda_sampled_backward(v4, shapes, torch.zeros(3, 0, H, L, P, 2, device=dev),
                                           torch.zeros(3, 0, H, L, P, device=dev), torch.zeros(3, 0, C, device=dev))
    assert gv.shape == v4.shape and float(gv.abs().max()) == 0.0 and gl.numel() == 0 and gw.numel() == 0
    # ragged: sequence 1 of 3 is empty
    R = 50
    value, off, lg, refer = _fused_inputs(syn, R, H, L, P, 3, Lv, C, 2, dev)
    ro = torch.tensor([0, 20, 20, R], dtype=torch.int32, device=dev)
    got = ops.msda_fused(value, shapes, off, lg, refer, H, P, 3, row_offsets=ro)
    a = ops.msda_fused(value[0:1], shapes, off[:20], lg[:20], refer[:20], H, P, 1)
    b = ops.msda_fused(value[2:3], shapes, off[20:], lg[20:], refer[20:], H, P, 1)
    assert torch.equal(got, torch.cat([a, b]))


def test_points_outside_contribute_zero(dev):
    """grid_sample zero padding (utils.py:65): every sampling point at least one pixel outside the maps -> exact zeros,
    forward and all three gradients."""
    from moyolo_b200 import ops, synthetic as syn
    shapes = [list(s) for s in syn.PYRAMIDS["tiny"]]
    value, loc, w, go = syn.make_core_grad_inputs(5, 1, 16, 8, 32, shapes, 4)
    loc = torch.where(torch.rand_like(loc) < 0.5, loc * 0 - 0.5, loc * 0 + 1.5)
    out = ops.msda_sampled(value.to(dev), shapes, loc.to(dev), w.to(dev))
    assert float(out.abs().max()) == 0.0
    for t in ops.msda_sampled_backward(value.to(dev), shapes, loc.to(dev), w.to(dev), go.to(dev)):
        assert float(t.abs().max()) == 0.0


def test_error_behaviour_matches_reference(dev):
    """transformer.py:201-202, 262, 284 and ms_deform_attn_cuda.cu:28-52: ValueError for bad dims, AssertionError
    when the shapes do not add up to Lv, RuntimeError for CPU / non-contiguous tensors in the legacy FFI."""
    import moyolo_b200 as m
    from moyolo_b200 import msda_ext, ops, synthetic as syn
    with pytest.raises(ValueError):
        m.MSDeformAttn(d_model=250, n_heads=8)
    attn = m.MSDeformAttn(256, 3, 8, 4).to(dev)
    shapes = [list(s) for s in syn.PYRAMIDS["tiny"]]
    Lv = syn.level_sizes(shapes)
    q, val = torch.randn(1, 5, 256, device=dev), torch.randn(1, Lv, 256, device=dev)
    with pytest.raises(ValueError, match="must be 2 or 4"):
        attn(q, torch.rand(1, 5, 1, 3, device=dev), val, shapes)
    with pytest.raises(AssertionError):
        attn(q, torch.rand(1, 5, 1, 4, device=dev), val[:, :-1], shapes)
    with pytest.raises(ValueError):
        ops.msda_sampled(val.view(1, Lv, 8, 32), shapes[:2], torch.rand(1, 5, 8, 3, 4, 2, device=dev),
                         torch.rand(1, 5, 8, 3, 4, device=dev))
    sh = torch.as_tensor(shapes, dtype=torch.long, device=dev)
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    v4 = val.view(1, Lv, 8, 32)
    loc, aw = torch.rand(1, 5, 8, 3, 4, 2, device=dev), torch.rand(1, 5, 8, 3, 4, device=dev)
    with pytest.raises(RuntimeError, match="contiguous"):
        msda_ext.ms_deform_attn_forward(v4.transpose(2, 3).contiguous().transpose(2, 3), sh, lsi, loc, aw, 64)
    with pytest.raises(RuntimeError, match="im2col_step"):
        msda_ext.ms_deform_attn_forward(v4.repeat(3, 1, 1, 1), sh, lsi, loc.repeat(3, 1, 1, 1, 1, 1),
                                        aw.repeat(3, 1, 1, 1, 1), 2)


def test_per_frame_reset_gives_shipped_semantics_and_writers(dev, tmp_path):
    """The reference as shipped never clears `is_first` (head.py:106,115): every frame starts without tracks and
    IDs are the running count, in query order, of score >= 0.4. `TrackEngine.reset()` before
    each frame reproduces that; without it tracks are carried. The device track table then feeds the result
    writers (f3)."""
    from moyolo_b200 import results as R, synthetic as syn
    from moyolo_b200.tracker import TrackEngine
    spec = syn.DecoderSpec()
    sd = syn.make_decoder_state(spec, 7)
    shapes = [list(s) for s in syn.PYRAMIDS["tiny"]]
    nd = 64
    gen = syn.SequenceGenerator(syn.SequenceSpec(name="tiny", n_frames=4, n_detect=nd, seed=3, shapes=shapes),
                                spec.d_model, dev)
    frames = [tuple(t.clone() for t in gen.next_frame()) for _ in range(4)]
    eng = TrackEngine(sd, spec, shapes, dev, "fp32", nd, 1)
    out = eng.step(frames[0][0][None], frames[0][1][None], frames[0][2][None])[0]
    sd = syn.calibrate_score_bias(sd, out["logits"], spec, 0.1)
    eng = TrackEngine(sd, spec, shapes, dev, "fp32", nd, 1)
    for f in frames:  # O2: reset every frame
        eng.reset()
        o = eng.step(f[0][None], f[1][None], f[2][None])[0]
        assert o["ids"].shape[0] == nd
        s, ids = o["scores"].cpu().numpy(), o["ids"].cpu().numpy()
        want = np.where(s >= np.float32(0.4), np.cumsum(s >= np.float32(0.4)) - 1, -1)
        assert np.array_equal(ids, want)
    eng.reset()
    carried = 0
    for f in frames:  # O3: carried tracks
        o = eng.step(f[0][None], f[1][None], f[2][None])[0]
        carried = max(carried, o["ids"].shape[0] - nd)
    assert carried > 0
    table = eng.track_table()
    assert table.shape[1] == 9 and table.shape[0] > 0
    lines = R.mot_challenge_lines(table, 640, 640, seq=0)
    assert len(lines) == int((table[:, 2] >= 0).sum()) and lines[0].count(",") == 9
    n = R.write_mot_challenge(tmp_path / "seq0.txt", table, 640, 640)
    assert n == len(lines) and (tmp_path / "seq0.txt").read_text().splitlines()[0] == lines[0].rstrip("\n")
    per_frame = R.save_txt_lines(table, 640, 640, save_conf=True)
    assert sorted(per_frame) == sorted({int(v) for v in table[:, 1].tolist()})


def _tiny5_batches(dev, S, n_frames, nd):
    from moyolo_b200 import synthetic as syn
    spec, shapes, sd, plant = syn.tracking_workload("tiny5", 7)
    gens = [syn.PlantedSequenceGenerator(syn.SequenceSpec("tiny", n_frames, nd, 1 + s, shapes=shapes), spec, plant)
            for s in range(S)]
    frames = [[tuple(t.clone() for t in g.next_frame()) for _ in range(n_frames)] for g in gens]
    batches = [tuple(torch.stack([frames[s][t][k] for s in range(S)]).to(dev) for k in range(3)) for t in range(n_frames)]
    return spec, shapes, sd, batches


def test_partial_reset_keeps_other_sequences_aligned(dev):
    """reset(seq) in the middle of a pipelined run (a new video starts in one lock-step slot, val.py:288-291): the
    reset slot continues exactly like a fresh engine from that frame on (IDs from 0), the other slot is not
    disturbed, and collect() keeps slicing both sequences' rows at the right offsets (`_T_before`)."""
    from moyolo_b200.tracker import DecoderWeights, TrackEngine
    nd, S, n_frames, cut = 48, 2, 10, 5
    spec, shapes, sd, batches = _tiny5_batches(dev, S, n_frames, nd)
    W = DecoderWeights(sd, spec, dev, "bf16")

    def run(reset_at=None, start=0):
        eng = TrackEngine(sd, spec, shapes, dev, "bf16", nd, S, weights=W)
        rows = {}
        for t in range(start, n_frames):
            if t == reset_at:
                eng.reset(0)
            f = eng.submit(*batches[t], want_rows=True)
            if f > 0:
                rows[t - 1] = [{k: v.clone() for k, v in o.items()} for o in eng.collect(f - 1)]
        rows[n_frames - 1] = [{k: v.clone() for k, v in o.items()} for o in eng.collect(n_frames - 1 - start)]
        return rows

    plain, with_reset, fresh = run(), run(reset_at=cut), run(start=cut)
    assert plain[cut - 1][0]["ids"].shape[0] > nd, "slot 0 carried no track before the reset"
    for t in range(n_frames):
        for k in ("ids", "boxes", "scores", "labels"):
            assert torch.equal(with_reset[t][1][k], plain[t][1][k]), ("untouched slot", t, k)
            if t < cut:
                assert torch.equal(with_reset[t][0][k], plain[t][0][k]), ("before the reset", t, k)
            elif k in ("ids", "labels"):
                assert torch.equal(with_reset[t][0][k], fresh[t][0][k]), ("after the reset", t, k)
            else:   # (the other slot carries tracks here and none in the fresh run: same values, other row offsets)
                assert torch.allclose(with_reset[t][0][k], fresh[t][0][k], atol=1e-2), ("after the reset", t, k)
    assert with_reset[cut][0]["ids"].shape[0] == nd and int(with_reset[cut][0]["ids"].max()) >= 0


def test_track_capacity_overflow_is_reported(dev):
    """More live tracks than `cap` in the dynamic engine: the reference has no such limit, so the engine must say so
    instead of silently dropping identities (kCtrlTrackOverflow)."""
    from moyolo_b200.tracker import TrackEngine
    nd, n_frames = 48, 10
    spec, shapes, sd, batches = _tiny5_batches(dev, 1, n_frames, nd)
    ok = TrackEngine(sd, spec, shapes, dev, "bf16", nd, 1)
    for b in batches:
        ok.submit(*b, want_rows=False)
    assert max(ok.n_tracks_host()) > 8
    eng = TrackEngine(sd, spec, shapes, dev, "bf16", nd, 1, cap=8)
    with pytest.raises(RuntimeError, match="cap=8"):
        for b in batches:
            eng.submit(*b, want_rows=False)
        eng.drain()
