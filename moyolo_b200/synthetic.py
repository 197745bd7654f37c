"""Deterministic synthetic weights, pyramids and sequences (SURVEY.md §8(d) "synthetic inputs").

Everything is generated on the CPU from `torch.Generator` seeds so that the golden-vector script
(run where the reference exists), the CPU tests and the GPU box regenerate bit-identical inputs;
golden files store a checksum of the regenerated inputs to detect RNG drift.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import torch

# feature-pyramid shapes [(H, W)] of the named configs
PYRAMIDS: Dict[str, List[Tuple[int, int]]] = {
    "C1": [(80, 80), (40, 40), (20, 20)],          # 640x640
    "MOT17": [(76, 136), (38, 68), (19, 34)],      # 1088x608
    "DanceTrack": [(100, 168), (50, 84), (25, 42)],  # 1344x800
    "KITTI": [(48, 156), (24, 78), (12, 39)],      # 1248x384
    "tiny": [(12, 16), (6, 8), (3, 4)],
}


@dataclass
class DecoderSpec:
    d_model: int = 256
    n_heads: int = 8
    d_ffn: int = 1024
    n_levels: int = 3
    n_points: int = 4
    n_layers: int = 6
    nc: int = 1
    pos_hidden: int = 512  # query_pos_head = MLP(4, 2*hd, hd, 2)
    qim_hidden: int = 256  # head.py:117-118

    def key(self) -> str:
        return f"d{self.d_model}h{self.n_heads}f{self.d_ffn}L{self.n_levels}P{self.n_points}n{self.n_layers}c{self.nc}"


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def _lin(g, out_f, in_f, w_std=None, b_std=0.02):
    w_std = w_std if w_std is not None else 1.0 / math.sqrt(in_f)
    return torch.randn(out_f, in_f, generator=g) * w_std, torch.randn(out_f, generator=g) * b_std


def _ln(g, d):
    return 1.0 + 0.1 * torch.randn(d, generator=g), 0.05 * torch.randn(d, generator=g)


def make_decoder_state(spec: DecoderSpec, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's state_dict key names.

    Keys: `layers.{i}.*` (decoder layers), `dec_bbox_head.{i}.layers.{j}.*`, `dec_score_head.{i}.*`,
    `query_pos_head.layers.{j}.*`, `denoising_class_embed.weight`, `track_embed.*` (QIM).
    sampling_offsets ~ N(0, 0.02) on top of the directional bias and attention_weights ~ N(0, 0.05)
    so the sampling pattern and softmax are non-degenerate (the module default zeros them).
    """
    g = _gen(seed)
    d, H, L, P = spec.d_model, spec.n_heads, spec.n_levels, spec.n_points
    sd: Dict[str, torch.Tensor] = {}

    def put(prefix, wb):
        sd[prefix + ".weight"], sd[prefix + ".bias"] = wb

    ang = torch.arange(H, dtype=torch.float32) * (2.0 * math.pi / H)
    ring = torch.stack([ang.cos(), ang.sin()], -1)
    ring = (ring / ring.abs().max(-1, keepdim=True)[0]).view(H, 1, 1, 2).repeat(1, L, P, 1)
    ring = ring * torch.arange(1, P + 1, dtype=torch.float32).view(1, 1, -1, 1)
    for i in range(spec.n_layers):
        p = f"layers.{i}."
        w, b = _lin(g, 3 * d, d)
        sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"] = w, b
        put(p + "self_attn.out_proj", _lin(g, d, d))
        sd[p + "norm1.weight"], sd[p + "norm1.bias"] = _ln(g, d)
        w, b = _lin(g, H * L * P * 2, d, w_std=0.02, b_std=0.0)
        sd[p + "cross_attn.sampling_offsets.weight"] = w
        sd[p + "cross_attn.sampling_offsets.bias"] = ring.reshape(-1).clone() + 0.1 * torch.randn(H * L * P * 2, generator=g)
        put(p + "cross_attn.attention_weights", _lin(g, H * L * P, d, w_std=0.05, b_std=0.1))
        put(p + "cross_attn.value_proj", _lin(g, d, d))
        put(p + "cross_attn.output_proj", _lin(g, d, d))
        sd[p + "norm2.weight"], sd[p + "norm2.bias"] = _ln(g, d)
        put(p + "linear1", _lin(g, spec.d_ffn, d))
        put(p + "linear2", _lin(g, d, spec.d_ffn))
        sd[p + "norm3.weight"], sd[p + "norm3.bias"] = _ln(g, d)
    for i in range(spec.n_layers):
        put(f"dec_bbox_head.{i}.layers.0", _lin(g, d, d))
        put(f"dec_bbox_head.{i}.layers.1", _lin(g, d, d))
        put(f"dec_bbox_head.{i}.layers.2", _lin(g, 4, d, w_std=0.2 / math.sqrt(d)))
        w, _ = _lin(g, spec.nc, d, w_std=0.1)
        sd[f"dec_score_head.{i}.weight"] = w
        sd[f"dec_score_head.{i}.bias"] = torch.full((spec.nc,), -3.3) + 0.05 * torch.randn(spec.nc, generator=g)
    put("query_pos_head.layers.0", _lin(g, spec.pos_hidden, 4, w_std=0.5))
    put("query_pos_head.layers.1", _lin(g, d, spec.pos_hidden))
    sd["denoising_class_embed.weight"] = torch.randn(spec.nc, d, generator=g)
    # QIM (MOTR/models/qim.py:85-105): dim_in = d, hidden = qim_hidden
    q = "track_embed."
    w, b = _lin(g, 3 * d, d)
    sd[q + "self_attn.in_proj_weight"], sd[q + "self_attn.in_proj_bias"] = w, b
    put(q + "self_attn.out_proj", _lin(g, d, d))
    put(q + "linear1", _lin(g, spec.qim_hidden, d))
    put(q + "linear2", _lin(g, d, spec.qim_hidden))
    put(q + "linear_feat1", _lin(g, spec.qim_hidden, d))
    put(q + "linear_feat2", _lin(g, d, spec.qim_hidden))
    for n in ("norm_feat", "norm1", "norm2"):
        sd[q + n + ".weight"], sd[q + n + ".bias"] = _ln(g, d)
    return sd


def make_selector_state(spec: DecoderSpec, ch: Sequence[int] = (256, 512, 512), seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init weights of the encoder-side query selection of MYDecoder (head.py:839-863) with the
    reference's key names: `input_proj.{l}.0.weight` (1x1 conv, no bias), `input_proj.{l}.1.*` (BatchNorm2d incl.
    running statistics), `enc_output.{0,1}.*` (Linear + LayerNorm), `enc_score_head.*`, `enc_bbox_head.layers.{j}.*`.
    Separate generator from make_decoder_state so existing goldens keep their weights."""
    g = _gen(7919 + seed)
    d = spec.d_model
    sd: Dict[str, torch.Tensor] = {}
    for l, c in enumerate(ch):
        sd[f"input_proj.{l}.0.weight"] = torch.randn(d, c, 1, 1, generator=g) / math.sqrt(c)
        sd[f"input_proj.{l}.1.weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
        sd[f"input_proj.{l}.1.bias"] = 0.05 * torch.randn(d, generator=g)
        sd[f"input_proj.{l}.1.running_mean"] = 0.1 * torch.randn(d, generator=g)
        sd[f"input_proj.{l}.1.running_var"] = 0.5 + torch.rand(d, generator=g)
    sd["enc_output.0.weight"], sd["enc_output.0.bias"] = _lin(g, d, d)
    sd["enc_output.1.weight"], sd["enc_output.1.bias"] = _ln(g, d)
    w, _ = _lin(g, spec.nc, d, w_std=0.15)
    sd["enc_score_head.weight"], sd["enc_score_head.bias"] = w, torch.full((spec.nc,), -2.0)
    sd["enc_bbox_head.layers.0.weight"], sd["enc_bbox_head.layers.0.bias"] = _lin(g, d, d)
    sd["enc_bbox_head.layers.1.weight"], sd["enc_bbox_head.layers.1.bias"] = _lin(g, d, d)
    sd["enc_bbox_head.layers.2.weight"], sd["enc_bbox_head.layers.2.bias"] = _lin(g, 4, d, w_std=0.3 / math.sqrt(d))
    return sd


def make_pyramid_maps(seed: int, B: int, shapes, ch: Sequence[int] = (256, 512, 512)) -> list:
    """Neck outputs [B, C_l, H_l, W_l] (fp32, NCHW as the reference's Conv2d takes them)."""
    g = _gen(seed)
    return [torch.randn(B, c, int(h), int(w), generator=g) for c, (h, w) in zip(ch, shapes)]


def sub_state(sd: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def level_sizes(shapes: Sequence[Sequence[int]]) -> int:
    return sum(int(h) * int(w) for h, w in shapes)


def make_core_inputs(seed: int, B: int, Q: int, n_heads: int, head_dim: int, shapes, n_points: int,
                     outside_frac: float = 0.1):
    """Inputs of the core op a1 (value, loc, weights); ~10% of points fall outside [0,1] (zero padding)
    and a few hit exact borders / pixel centres."""
    g = _gen(seed)
    Lv, L = level_sizes(shapes), len(shapes)
    value = torch.randn(B, Lv, n_heads, head_dim, generator=g)
    loc = torch.rand(B, Q, n_heads, L, n_points, 2, generator=g)
    far = torch.rand(B, Q, n_heads, L, n_points, 1, generator=g) < outside_frac
    loc = torch.where(far, loc * 1.6 - 0.3, loc)
    flat = loc.view(-1)
    specials = torch.tensor([0.0, 1.0, -0.2, 1.25, 0.5, 0.5 / shapes[0][1], 1.0 - 1e-7, 1e-7])
    n = min(len(specials), flat.numel())
    flat[:n] = specials[:n]
    w = torch.rand(B, Q, n_heads, L, n_points, generator=g) + 1e-5
    w = w / w.sum((-1, -2), keepdim=True)
    return value, loc, w


def make_core_grad_inputs(seed: int, B: int, Q: int, n_heads: int, head_dim: int, shapes, n_points: int,
                          outside_frac: float = 0.1):
    """Inputs of the core op's backward: as make_core_inputs but WITHOUT the exact-border / pixel-centre
    specials (the location gradient is discontinuous where a pixel coordinate is an integer, so two correct fp32
    implementations may legitimately pick different one-sided derivatives there) plus the output gradient."""
    g = _gen(seed)
    Lv, L = level_sizes(shapes), len(shapes)
    value = torch.randn(B, Lv, n_heads, head_dim, generator=g)
    loc = torch.rand(B, Q, n_heads, L, n_points, 2, generator=g)
    far = torch.rand(B, Q, n_heads, L, n_points, 1, generator=g) < outside_frac
    loc = torch.where(far, loc * 1.6 - 0.3, loc)
    w = torch.rand(B, Q, n_heads, L, n_points, generator=g) + 1e-5
    w = w / w.sum((-1, -2), keepdim=True)
    grad_out = torch.randn(B, Q, n_heads * head_dim, generator=g)
    return value, loc, w, grad_out


def make_module_inputs(seed: int, B: int, Q: int, d_model: int, shapes, ref_dim: int = 4, ref_levels: int = 1):
    """Inputs of MSDeformAttn.forward / a decoder layer: query, refer_bbox (sigmoid space), feats, query_pos."""
    g = _gen(seed)
    Lv = level_sizes(shapes)
    query = torch.randn(B, Q, d_model, generator=g)
    cxcy = torch.rand(B, Q, ref_levels, 2, generator=g)
    if ref_dim == 4:
        wh = torch.rand(B, Q, ref_levels, 2, generator=g) * 0.48 + 0.02
        refer = torch.cat([cxcy, wh], -1)
    else:
        refer = cxcy
    feats = torch.randn(B, Lv, d_model, generator=g)
    query_pos = torch.randn(B, Q, d_model, generator=g) * 0.5
    return query, refer, feats, query_pos


def make_decoder_inputs(seed: int, B: int, Q: int, d_model: int, shapes):
    """Inputs of the 6-layer decoders: embed, refer_bbox in LOGIT space, feats, fixed query_pos."""
    g = _gen(seed)
    Lv = level_sizes(shapes)
    embed = torch.randn(B, Q, d_model, generator=g)
    cxcy = torch.rand(B, Q, 2, generator=g) * 0.9 + 0.05
    wh = torch.rand(B, Q, 2, generator=g) * 0.28 + 0.02
    boxes = torch.cat([cxcy, wh], -1)
    refer_logit = torch.log(boxes / (1 - boxes))
    feats = torch.randn(B, Lv, d_model, generator=g)
    query_pos = torch.randn(B, Q, d_model, generator=g) * 0.5
    return embed, refer_logit, feats, query_pos


def make_denoising_masks(seed: int, B: int, Q: int, Lv: int, n_groups: int = 2, group_size: int = 6, pad_frac: float = 0.2):
    """Masks in the shapes the reference's training path produces: the [Q, Q] bool self-attention mask of the
    denoising groups (ultralytics/models/utils/ops.py:363-375: matching queries cannot see the reconstruction
    queries, the groups cannot see each other; True = masked) and a [B, Lv] bool value padding mask (True rows of
    the value tensor are zeroed, transformer.py:265-266)."""
    num_dn = n_groups * group_size
    assert num_dn < Q
    attn = torch.zeros(Q, Q, dtype=torch.bool)
    attn[num_dn:, :num_dn] = True
    for i in range(n_groups):
        attn[group_size * i:group_size * (i + 1), group_size * (i + 1):num_dn] = True
        attn[group_size * i:group_size * (i + 1), :group_size * i] = True
    pad = torch.rand(B, Lv, generator=_gen(seed)) < pad_frac
    return attn, pad


def checksum(*tensors: torch.Tensor) -> float:
    """Order-sensitive float64 checksum used to verify regenerated inputs against a golden file."""
    acc = 0.0
    for i, t in enumerate(tensors):
        t64 = t.detach().double().reshape(-1)
        idx = torch.arange(1, t64.numel() + 1, dtype=torch.float64)
        acc += float((t64 * torch.cos(idx * 0.001 + i)).sum())
    return acc


@dataclass
class SequenceSpec:
    """A synthetic video: temporally coherent pyramid features and detect queries."""
    name: str = "MOT17"
    n_frames: int = 300
    n_detect: int = 300
    seed: int = 0
    drift: float = 0.05
    shapes: List[Tuple[int, int]] = field(default_factory=lambda: list(PYRAMIDS["MOT17"]))


class SequenceGenerator:
    """feats_0 ~ N(0,1), feats_t = feats_{t-1} + drift*N(0,1); detect embeddings/boxes drift likewise.

    Generated on `device` with a device generator (bench) or on the CPU (parity tests): frames of a
    CPU-generated sequence are bit-reproducible anywhere.
    """

    def __init__(self, spec: SequenceSpec, d_model: int = 256, device="cpu", dtype=torch.float32):
        self.spec, self.d, self.device, self.dtype = spec, d_model, torch.device(device), dtype
        self.g = torch.Generator(device=self.device)
        self.g.manual_seed(1000 * spec.seed + 17)
        Lv = level_sizes(spec.shapes)
        self.feats = torch.randn(Lv, d_model, generator=self.g, device=self.device)
        self.det_embed = torch.randn(spec.n_detect, d_model, generator=self.g, device=self.device)
        cxcy = torch.rand(spec.n_detect, 2, generator=self.g, device=self.device) * 0.9 + 0.05
        wh = torch.rand(spec.n_detect, 2, generator=self.g, device=self.device) * 0.28 + 0.02
        self.det_box = torch.cat([cxcy, wh], -1)
        self.t = 0

    def next_frame(self):
        """Returns (feats [Lv, C], detect_embed [nd, C], detect_refer_logit [nd, 4])."""
        if self.t > 0:
            s = self.spec.drift
            self.feats = self.feats + s * torch.randn(self.feats.shape, generator=self.g, device=self.device)
            self.det_embed = self.det_embed + s * torch.randn(self.det_embed.shape, generator=self.g,
                                                              device=self.device)
            jitter = 0.01 * torch.randn(self.det_box.shape, generator=self.g, device=self.device)
            self.det_box = (self.det_box + jitter).clamp(0.02, 0.98)
        self.t += 1
        b = self.det_box
        return self.feats.to(self.dtype), self.det_embed, torch.log(b / (1 - b))


def calibrate_score_bias(sd: Dict[str, torch.Tensor], frame0_logits: torch.Tensor, spec: DecoderSpec,
                         birth_frac: float = 0.05, thresh: float = 0.4) -> Dict[str, torch.Tensor]:
    """Shift the last decoder score-head bias so that `birth_frac` of frame-0 queries score >= thresh
    ("score_head ... so scores straddle 0.4/0.5"). `frame0_logits` are the raw
    logits [N, nc] of frame 0 computed with the un-shifted weights; returns a new state dict."""
    best = torch.sort(frame0_logits.detach().float().cpu().max(-1).values).values
    n = best.numel()
    k = min(max(int(round((1.0 - birth_frac) * n)), 1), n - 1)
    lo, hi = max(k - 3, 1), min(k + 3, n - 1)
    gaps = best[lo:hi + 1] - best[lo - 1:hi]
    j = lo + int(torch.argmax(gaps))
    q = 0.5 * (best[j] + best[j - 1]).item()  # threshold lands mid-gap: no score sits on the threshold
    target = math.log(thresh / (1.0 - thresh))
    out = dict(sd)
    key = f"dec_score_head.{spec.n_layers - 1}.bias"
    out[key] = sd[key] + (target - q)
    return out


# ------------------------------------------------------------------------------------------------
# Planted-margin tracking workload (VERDICT r1 task 1a): births and deaths are planted, every score sits far
# from the 0.4 / 0.5 thresholds of RuntimeTrackerBase (head.py:1146), so free-running ID parity can be checked
# on every frame in fp32 AND bf16.
#
# Mechanism. The first P = 1 + nc channels of the residual stream are "protected": no sub-layer writes to them
# (rows [:P] of out_proj / output_proj / linear2 and their biases are zero), every LayerNorm scales them by one
# common gain and no shift, so their SIGNS and RATIOS survive all 18 post-norm steps of the decoder while the
# other 256 - P channels stay fully random. Channel 0 carries "objectness" (+A: an object, -A: none), channel
# 1 + c the class one-hot. The last score head reads logit_c = k_obj * x[0] + k_cls * x[1 + c] + small random
# terms, so sign(x[0]) decides score >> 0.5 or << 0.4 and argmax_c is the planted class.
#   * births: each frame a random subset of the detect queries carries +A in channel 0 (head.py:1232-1236);
#   * carried tracks restart from denoising_class_embed[label] (head.py:888-900): classes in `persistent` hold
#     +A there (the track keeps scoring high), the others -A (the track scores low on every later frame, its
#     disappear_time runs up and it dies after `miss_tolerance` frames, head.py:1238-1243).
# ------------------------------------------------------------------------------------------------
@dataclass
class PlantSpec:
    births_per_frame: float = 8.0
    persistent: Tuple[int, ...] = ()      # class ids whose tracks never die
    amp: float = 4.0                      # |planted value| in the input embeddings
    ln_gain: float = 1.4                  # LayerNorm weight of the protected channels
    k_obj: float = 0.4
    k_cls: float = 0.25
    bias: float = -2.0
    w_noise: float = 0.012                # std of the random part of the score head
    box_gain: float = 0.25
    persistent_until: int = 32            # persistent classes are only born in frames < this (bounds the track count)                # scale of the last box-head layer


def make_tracking_state(spec: DecoderSpec, seed: int = 0, plant: PlantSpec = PlantSpec()) -> Dict[str, torch.Tensor]:
    """make_decoder_state with the protected objectness / class channels described above."""
    sd = dict(make_decoder_state(spec, seed))
    P, nc, A = 1 + spec.nc, spec.nc, plant.amp
    for i in range(spec.n_layers):
        p = f"layers.{i}."
        for name in ("self_attn.out_proj", "cross_attn.output_proj", "linear2"):
            w, b = sd[p + name + ".weight"].clone(), sd[p + name + ".bias"].clone()
            w[:P] = 0
            b[:P] = 0
            sd[p + name + ".weight"], sd[p + name + ".bias"] = w, b
        for name in ("norm1", "norm2", "norm3"):
            w, b = sd[p + name + ".weight"].clone(), sd[p + name + ".bias"].clone()
            w[:P] = plant.ln_gain
            b[:P] = 0
            sd[p + name + ".weight"], sd[p + name + ".bias"] = w, b
    for i in range(spec.n_layers):   # small refinement steps: errors of carried boxes must not compound across frames
        sd[f"dec_bbox_head.{i}.layers.2.weight"] = sd[f"dec_bbox_head.{i}.layers.2.weight"] * plant.box_gain
    g = _gen(104729 + seed)
    last = spec.n_layers - 1
    w = torch.randn(nc, spec.d_model, generator=g) * plant.w_noise
    w[:, :P] = 0
    w[:, 0] = plant.k_obj
    for c in range(nc):
        w[c, 1 + c] = plant.k_cls
    sd[f"dec_score_head.{last}.weight"] = w
    sd[f"dec_score_head.{last}.bias"] = torch.full((nc,), plant.bias)
    ce = sd["denoising_class_embed.weight"].clone()
    ce[:, :P] = -A
    for c in range(nc):
        ce[c, 0] = A if c in plant.persistent else -A
        ce[c, 1 + c] = A
    sd["denoising_class_embed.weight"] = ce
    return sd


class PlantedSequenceGenerator(SequenceGenerator):
    """SequenceGenerator whose detect embeddings carry the planted objectness / class channels and whose feature
    maps are SPATIALLY smooth (coarse noise, bilinearly upsampled 8x per level, unit variance): a white-noise map
    makes bilinear sampling ill-conditioned in the sampling location (d value / d loc grows with the map width), so
    rounding noise in the offsets would be amplified ~W-fold per layer -- no real backbone produces such maps."""

    SMOOTH = 8  # upsampling factor of the coarse noise

    def __init__(self, spec: SequenceSpec, dspec: DecoderSpec, plant: PlantSpec = PlantSpec(), device="cpu",
                 dtype=torch.float32):
        super().__init__(spec, dspec.d_model, device, dtype)
        self.plant, self.nc = plant, dspec.nc
        self.feats = self._smooth_noise()

    def _smooth_noise(self) -> torch.Tensor:
        import torch.nn.functional as F
        out = []
        for h, w in self.spec.shapes:
            ch, cw = -(-int(h) // self.SMOOTH) + 1, -(-int(w) // self.SMOOTH) + 1
            coarse = torch.randn(1, self.d, ch, cw, generator=self.g, device=self.device)
            m = F.interpolate(coarse, size=(int(h), int(w)), mode="bilinear", align_corners=True)
            out.append(m[0].flatten(1).t())
        x = torch.cat(out, 0)
        return x / x.std()

    def next_frame(self):
        if self.t > 0:
            s = self.spec.drift
            self.feats = self.feats + s * self._smooth_noise()
            self.det_embed = self.det_embed + s * torch.randn(self.det_embed.shape, generator=self.g,
                                                              device=self.device)
            jitter = 0.01 * torch.randn(self.det_box.shape, generator=self.g, device=self.device)
            self.det_box = (self.det_box + jitter).clamp(0.02, 0.98)
        self.t += 1
        b = self.det_box
        feats, de, dr = self.feats.to(self.dtype), self.det_embed, torch.log(b / (1 - b))
        nd, A, nc = de.shape[0], self.plant.amp, self.nc
        fire = torch.rand(nd, generator=self.g, device=self.device) < self.plant.births_per_frame / nd
        cls = torch.randint(0, nc, (nd,), generator=self.g, device=self.device)
        if self.t > self.plant.persistent_until and self.plant.persistent:   # (self.t is already this frame's index + 1)
            is_p = torch.zeros(nd, dtype=torch.bool, device=self.device)
            for c in self.plant.persistent:
                is_p |= cls == c
            fire &= ~is_p
        de = de.clone()
        de[:, 0] = torch.where(fire, A, -A)
        de[:, 1:1 + nc] = -A
        de[torch.arange(nd, device=self.device), 1 + cls] = A
        return feats, de, dr


# named tracking workloads (BASELINE.json configs[1..3]): pyramid, classes, planted dynamics
TRACKING_WORKLOADS: Dict[str, dict] = {
    # MOT17: one class, every track dies miss_tolerance frames after its birth -> constant churn, ~40 carried tracks
    "MOT17": dict(pyramid="MOT17", nc=1, plant=PlantSpec(births_per_frame=8.0, persistent=())),
    # DanceTrack: one class, tracks persist -> the query count grows towards 500 (configs[2])
    "DanceTrack": dict(pyramid="DanceTrack", nc=1, plant=PlantSpec(births_per_frame=5.0, persistent=(0,))),
    # KITTI: KITTI.yaml's 5 classes, classes 0-2 persist, 3-4 die (configs[3])
    "KITTI": dict(pyramid="KITTI", nc=5, plant=PlantSpec(births_per_frame=6.0, persistent=(0, 1, 2))),
    "C1": dict(pyramid="C1", nc=1, plant=PlantSpec(births_per_frame=8.0, persistent=())),
    "tiny": dict(pyramid="tiny", nc=1, plant=PlantSpec(births_per_frame=4.0, persistent=())),
    "tiny5": dict(pyramid="tiny", nc=5, plant=PlantSpec(births_per_frame=4.0, persistent=(0, 1, 2))),
}


def tracking_workload(name: str, seed: int = 0):
    """(DecoderSpec, shapes, state_dict, PlantSpec) of a named planted tracking workload."""
    w = TRACKING_WORKLOADS[name]
    spec = DecoderSpec(nc=w["nc"])
    shapes = [list(s) for s in PYRAMIDS[w["pyramid"]]]
    return spec, shapes, make_tracking_state(spec, seed, w["plant"]), w["plant"]
