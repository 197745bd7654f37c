"""Decoder-layer executor: the kernel schedule behind the drop-in modules.

One decoder layer (ultralytics/nn/modules/transformer.py:627-652 / :431-450) is run as the chain

    in-proj GEMMs (q,k from x+pos; v from x) -> self-attention -> out-proj GEMM -> add+LayerNorm
    -> fused offsets|logits GEMM -> deformable gather -> output-proj GEMM -> add+LayerNorm
    -> FFN GEMM(ReLU) -> FFN GEMM -> add+LayerNorm [-> + pos for the next layer]

with `value_proj` hoisted out of the layer loop: `feats` is the same tensor in every layer
(transformer.py:705 passes it unchanged), so all layers' values come from ONE GEMM
[B*Lv, C] x [C, n_layers*C] and each layer's gather reads its own column slice in place.

Two precisions: "fp32" (CUDA-core GEMMs, reference-grade parity) and "bf16" (bf16 GEMM operands and
value tensor on tcgen05; residual stream, LayerNorm, softmax, sampling locations and accumulation
stay fp32).
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import torch

from . import _lib, ops

_DEFAULT_PRECISION = "fp32"
_GEMM_ENGINE = _lib.GEMM_AUTO
# bf16 frame path: merge the q,k / v in-projections into one dual-operand GEMM and fold residual-add +
# LayerNorm into the epilogue of the GEMM that feeds it (8 launches per decoder layer instead of 12).
FUSE_EPILOGUES = os.environ.get("MOYOLO_FUSE", "1") != "0"


def set_default_precision(p: str) -> None:
    global _DEFAULT_PRECISION
    if p not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    _DEFAULT_PRECISION = p


def get_default_precision() -> str:
    return _DEFAULT_PRECISION


def set_gemm_engine(engine: int) -> None:
    """Force the GEMM engine (GEMM_AUTO / GEMM_SIMT / GEMM_TCGEN05) — used by tests and profiling."""
    global _GEMM_ENGINE
    _GEMM_ENGINE = engine


def fused_epilogues(dt: torch.dtype, C: int) -> bool:
    """True when the dual-operand GEMM and the GEMM+LayerNorm epilogue serve this configuration."""
    return FUSE_EPILOGUES and dt == torch.bfloat16 and C == 256 and _GEMM_ENGINE != _lib.GEMM_SIMT


def lp_dtype(precision: str) -> torch.dtype:
    return torch.bfloat16 if precision == "bf16" else torch.float32


def _w(t: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
    return t.detach().to(dt).contiguous()


def _f(t: torch.Tensor) -> torch.Tensor:
    return t.detach().float().contiguous()


def _versions(module: torch.nn.Module):
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


class LinearPack:
    __slots__ = ("w", "b")

    def __init__(self, weight, bias, dt):
        self.w = _w(weight, dt)
        self.b = _f(bias) if bias is not None else None


class MsdaPack:
    """Weights of one MSDeformAttn (transformer.py:214-217); offsets and logits linears fused."""

    def __init__(self, m, dt):
        self.n_heads, self.n_levels, self.n_points, self.d_model = m.n_heads, m.n_levels, m.n_points, m.d_model
        self.n_off = m.n_heads * m.n_levels * m.n_points * 2
        self.offlog = LinearPack(torch.cat([m.sampling_offsets.weight, m.attention_weights.weight], 0),
                                 torch.cat([m.sampling_offsets.bias, m.attention_weights.bias], 0), dt)
        self.value = LinearPack(m.value_proj.weight, m.value_proj.bias, dt)
        self.out = LinearPack(m.output_proj.weight, m.output_proj.bias, dt)
        self.softmax_mode = _lib.SOFTMAX_PLUS1 if getattr(m, "my_softmax", False) and \
            type(m).__name__ == "MOTRMSDeformAttn" else _lib.SOFTMAX


class LayerPack:
    def __init__(self, layer, dt):
        sa = layer.self_attn
        C = sa.embed_dim
        self.C, self.n_heads = C, sa.num_heads
        self.qkv = LinearPack(sa.in_proj_weight, sa.in_proj_bias, dt)
        self.qk = LinearPack(sa.in_proj_weight[:2 * C], sa.in_proj_bias[:2 * C], dt)
        self.v = LinearPack(sa.in_proj_weight[2 * C:], sa.in_proj_bias[2 * C:], dt)
        self.o = LinearPack(sa.out_proj.weight, sa.out_proj.bias, dt)
        self.msda = MsdaPack(layer.cross_attn, dt)
        self.ffn1 = LinearPack(layer.linear1.weight, layer.linear1.bias, dt)
        self.ffn2 = LinearPack(layer.linear2.weight, layer.linear2.bias, dt)
        self.norms = [(_f(n.weight), _f(n.bias), n.eps) for n in (layer.norm1, layer.norm2, layer.norm3)]
        if not isinstance(layer.act, torch.nn.ReLU):
            raise NotImplementedError("moyolo_b200 decoder layer: only ReLU FFN activation is implemented")


class MlpPack:
    def __init__(self, mlp, dt):
        ls = list(mlp.layers)
        self.hidden = [LinearPack(l.weight, l.bias, dt) for l in ls[:-1]]
        self.last_w, self.last_b = _f(ls[-1].weight), _f(ls[-1].bias)
        self.last_lp = LinearPack(ls[-1].weight, ls[-1].bias, dt)
        self.first_w_f32 = _f(ls[0].weight) if ls[0].in_features == 4 else None


def cached_pack(module, key: str, builder, dt):
    """Per-module weight pack, rebuilt when any parameter is modified in place or re-assigned."""
    cache = module.__dict__.setdefault("_moyolo_packs", {})
    ver = _versions(module)
    hit = cache.get((key, dt))
    if hit is None or hit[0] != ver:
        hit = (ver, builder(module, dt))
        cache[(key, dt)] = hit
    return hit[1]


def dense_row_offsets(B: int, Q: int, device) -> tuple:
    host = [b * Q for b in range(B + 1)]
    return torch.tensor(host, dtype=torch.int32, device=device), host


_ro_cache = {}


def cached_dense_row_offsets(B: int, Q: int, device):
    key = (B, Q, str(device))
    if key not in _ro_cache:
        _ro_cache[key] = dense_row_offsets(B, Q, device)
    return _ro_cache[key]


def project_values(feats_lp: torch.Tensor, msda_packs: Sequence[MsdaPack], zero_rows, dt) -> torch.Tensor:
    """One GEMM for every layer's value_proj: feats_lp [B*Lv, C] -> [B*Lv, n_layers*C]."""
    if len(msda_packs) == 1:
        w, b = msda_packs[0].value.w, msda_packs[0].value.b
    else:
        w = torch.cat([p.value.w for p in msda_packs], 0)
        b = torch.cat([p.value.b for p in msda_packs], 0)
    return ops.linear(feats_lp, w, b, out_dtype=dt, zero_rows=zero_rows, engine=_GEMM_ENGINE)


class ValueProjPack:
    def __init__(self, layers, dt):
        self.w = torch.cat([_w(l.cross_attn.value_proj.weight, dt) for l in layers], 0).contiguous()
        self.b = torch.cat([_f(l.cross_attn.value_proj.bias) for l in layers], 0).contiguous()


def msda_forward(pack: MsdaPack, xq_lp: torch.Tensor, refer: torch.Tensor, value_view: torch.Tensor, shapes,
                 batch: int, row_offsets, dt) -> torch.Tensor:
    """offsets|logits GEMM -> fused gather -> output_proj GEMM. Returns fp32 [R, C]."""
    ol = ops.linear(xq_lp, pack.offlog.w, pack.offlog.b, out_dtype=torch.float32, engine=_GEMM_ENGINE)
    g = ops.msda_fused(value_view, shapes, ol[:, :pack.n_off], ol[:, pack.n_off:], refer, pack.n_heads,
                       pack.n_points, batch, pack.softmax_mode, row_offsets)
    return ops.linear(g, pack.out.w, pack.out.b, out_dtype=torch.float32, engine=_GEMM_ENGINE)


def run_layer(pk: LayerPack, x_f32, x_lp, xq_lp, refer, value_view, shapes, batch, row_offsets, row_offsets_host,
              dense: bool, attn_mask, pos_cur, pos_next, dt):
    """One post-norm decoder layer.

    x_f32 residual stream, x_lp its GEMM-operand copy, xq_lp = (x + pos_cur) operand copy.
    pos_cur / pos_next: fp32 [R, C] positional embeddings of this / the next layer (None = no pos).
    Returns (x_f32, x_lp, xq_lp for the next layer or None when pos_next is None).
    """
    R, C = x_f32.shape
    eng = _GEMM_ENGINE
    qkv = torch.empty(R, 3 * C, dtype=dt, device=x_f32.device)
    ops.linear(xq_lp, pk.qk.w, pk.qk.b, out=qkv[:, :2 * C], engine=eng)
    ops.linear(x_lp, pk.v.w, pk.v.b, out=qkv[:, 2 * C:], engine=eng)
    att = ops.self_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], row_offsets, row_offsets_host,
                             pk.n_heads, attn_mask)
    t = ops.linear(att, pk.o.w, pk.o.b, out_dtype=torch.float32, engine=eng)
    g, b_, e = pk.norms[0]
    # after norm1 only the residual (fp32) and the cross-attention query operand (x + pos) are needed
    x1_f32, x1_lp, x1q_lp = ops.add_layernorm(t, x_f32, g, b_, e, want_f32=True, want_lp=pos_cur is None,
                                              lp_dtype=dt, pos=pos_cur)
    t2 = msda_forward(pk.msda, x1q_lp if pos_cur is not None else x1_lp, refer, value_view, shapes, batch,
                      None if dense else row_offsets, dt)
    g, b_, e = pk.norms[1]
    x2_f32, x2_lp, _ = ops.add_layernorm(t2, x1_f32, g, b_, e, want_f32=True, want_lp=True, lp_dtype=dt)
    h = ops.linear(x2_lp, pk.ffn1.w, pk.ffn1.b, relu=True, engine=eng)
    t3 = ops.linear(h, pk.ffn2.w, pk.ffn2.b, out_dtype=torch.float32, engine=eng)
    g, b_, e = pk.norms[2]
    x3_f32, x3_lp, x3q_lp = ops.add_layernorm(t3, x2_f32, g, b_, e, want_f32=True, want_lp=True, lp_dtype=dt,
                                              pos=pos_next)
    return x3_f32, x3_lp, x3q_lp


def bbox_head(mp: MlpPack, x_lp, refer, out=None):
    """sigmoid(MLP(x) + inverse_sigmoid(refer)) (transformer.py:709); last MLP layer fused with the refine."""
    h = x_lp
    for lin in mp.hidden:
        h = ops.linear(h, lin.w, lin.b, relu=True, engine=_GEMM_ENGINE)
    return ops.box_refine(h, mp.last_w, mp.last_b, refer, out=out)


def pos_mlp_forward(mp: MlpPack, refer, dt):
    """pos_mlp(refer) for DeformableTransformerDecoder (transformer.py:491): MLP(4, 2*hd, hd, 2)."""
    if len(mp.hidden) != 1 or mp.hidden[0].w.shape[1] != 4:
        raise NotImplementedError("pos_mlp must be MLP(4, hidden, out, num_layers=2)")
    first = mp.hidden[0]
    h = ops.linear_k4_relu(refer, mp.first_w_f32, first.b, dt)
    return ops.linear(h, mp.last_lp.w, mp.last_lp.b, out_dtype=torch.float32, engine=_GEMM_ENGINE)


# ------------------------------------------------------------------------------------------------
# Workspace variants used by the frame engine (moyolo_b200.tracker): identical kernel chain, but every
# output goes to a pre-allocated FrameWorkspace buffer so a frame is allocation-free and its box head /
# value projection can run on side branches of the frame graph.
# ------------------------------------------------------------------------------------------------
def offlog_width(spec) -> int:
    """Columns of the fused sampling_offsets|attention_weights GEMM: H*L*P*2 + H*L*P."""
    return spec.n_heads * spec.n_levels * spec.n_points * 3


def qkv_proj(xq_lp, x_lp, w, b, out, C: int, eng) -> None:
    """nn.MultiheadAttention in-projection with q = k = x + pos, v = x (transformer.py:637-638):
    out[:, :2C] = xq . [Wq;Wk]^T, out[:, 2C:] = x . Wv^T."""
    if fused_epilogues(xq_lp.dtype, C):
        ops.linear_dual(xq_lp, x_lp, 2 * C, w, b, out=out)
        return
    # (the row-block views are made once per weight and kept on it: ops.linear caches per-weight data on the tensor
    # OBJECT it is handed, so the same objects must come back every call)
    parts = getattr(w, "_moyolo_qkv_parts", None)
    if parts is None:
        parts = (w[:2 * C], b[:2 * C], w[2 * C:], b[2 * C:])
        w._moyolo_qkv_parts = parts
    ops.linear(xq_lp, parts[0], parts[1], out=out[:, :2 * C], engine=eng)
    ops.linear(x_lp, parts[2], parts[3], out=out[:, 2 * C:], engine=eng)


def run_layer_ws(pk: LayerPack, ws, refer, value_view, shapes, batch, row_offsets, row_offsets_host, pos_cur,
                 pos_next, dt, before_gather=None, score=None, gather_probe=None) -> None:
    """One post-norm decoder layer on workspace buffers. In: ws.x (fp32 residual), ws.x_lp, ws.xq_lp
    (= x + pos operand). Out: the same three for the next layer (xq only when pos_next is given)."""
    C = pk.C
    eng = _GEMM_ENGINE
    f32 = dt == torch.float32
    fuse = fused_epilogues(dt, C)
    pkm = pk.msda
    if fuse:
        ops.linear_dual(ws.xq_lp, ws.x_lp, 2 * C, pk.qkv.w, pk.qkv.b, out=ws.qkv)
        ops.self_attention(ws.qkv[:, :C], ws.qkv[:, C:2 * C], ws.qkv[:, 2 * C:], row_offsets, row_offsets_host,
                           pk.n_heads, None, out=ws.att)
        g, b_, e = pk.norms[0]
        if pos_cur is not None:
            ops.linear_add_layernorm(ws.att, pk.o.w, pk.o.b, ws.x, g, b_, e, out_f32=ws.x1, pos=pos_cur,
                                     out_pos=ws.x1q_lp)
        else:
            ops.linear_add_layernorm(ws.att, pk.o.w, pk.o.b, ws.x, g, b_, e, out_f32=ws.x1, out_lp=ws.x1q_lp)
        if ops.proj_fused_supported(dt, pkm.n_heads, C // pkm.n_heads, pkm.n_levels, pkm.n_points, ws.R):
            # offsets|logits projection inside the gather kernel: one launch for transformer.py:268-285
            if before_gather is not None:
                before_gather()
            ops.msda_proj_fused(value_view, shapes, ws.x1q_lp, pkm.offlog.w, pkm.offlog.b, refer, pkm.n_heads,
                                pkm.n_points, batch, pkm.softmax_mode, row_offsets, out=ws.g, probe=gather_probe)
        else:
            ops.linear(ws.x1q_lp, pkm.offlog.w, pkm.offlog.b, out=ws.ol, engine=eng)
            if before_gather is not None:
                before_gather()
            ops.msda_fused(value_view, shapes, ws.ol[:, :pkm.n_off], ws.ol[:, pkm.n_off:], refer, pkm.n_heads,
                           pkm.n_points, batch, pkm.softmax_mode, row_offsets, out=ws.g, probe=gather_probe)
        g, b_, e = pk.norms[1]
        ops.linear_add_layernorm(ws.g, pkm.out.w, pkm.out.b, ws.x1, g, b_, e, out_f32=ws.x2, out_lp=ws.x2_lp)
        g, b_, e = pk.norms[2]
        if ops.ffn_fused_supported(dt, C, pk.ffn1.w.shape[0]):   # FFN1 + FFN2 + residual + LayerNorm in one launch
            ops.ffn_add_layernorm(ws.x2_lp, pk.ffn1.w, pk.ffn1.b, pk.ffn2.w, pk.ffn2.b, ws.h, ws.x2, g, b_, e,
                                  out_f32=ws.x, out_lp=ws.x_lp, pos=pos_next,
                                  out_pos=ws.xq_lp if pos_next is not None else None)
            return
        ops.linear(ws.x2_lp, pk.ffn1.w, pk.ffn1.b, relu=True, out=ws.h, engine=eng)
        if score is not None and pos_next is None:   # last layer: class-score head fused behind the LayerNorm
            sw, sb, logits, scores, labels = score
            ops.linear_add_layernorm_scores(ws.h, pk.ffn2.w, pk.ffn2.b, ws.x2, g, b_, e, sw, sb, out_f32=ws.x,
                                            out_lp=ws.x_lp, logits=logits, scores=scores, labels=labels)
            return
        ops.linear_add_layernorm(ws.h, pk.ffn2.w, pk.ffn2.b, ws.x2, g, b_, e, out_f32=ws.x, out_lp=ws.x_lp,
                                 pos=pos_next, out_pos=ws.xq_lp if pos_next is not None else None)
        return
    ops.linear(ws.xq_lp, pk.qk.w, pk.qk.b, out=ws.qkv[:, :2 * C], engine=eng)
    ops.linear(ws.x_lp, pk.v.w, pk.v.b, out=ws.qkv[:, 2 * C:], engine=eng)
    ops.self_attention(ws.qkv[:, :C], ws.qkv[:, C:2 * C], ws.qkv[:, 2 * C:], row_offsets, row_offsets_host,
                       pk.n_heads, None, out=ws.att)
    ops.linear(ws.att, pk.o.w, pk.o.b, out=ws.t, engine=eng)
    g, b_, e = pk.norms[0]
    ops.add_layernorm(ws.t, ws.x, g, b_, e, out_f32=ws.x1, pos=pos_cur, out_pos=ws.x1q_lp)
    ops.linear(ws.x1q_lp, pkm.offlog.w, pkm.offlog.b, out=ws.ol, engine=eng)
    if before_gather is not None:
        before_gather()
    ops.msda_fused(value_view, shapes, ws.ol[:, :pkm.n_off], ws.ol[:, pkm.n_off:], refer, pkm.n_heads, pkm.n_points,
                   batch, pkm.softmax_mode, row_offsets, out=ws.g, probe=gather_probe)
    ops.linear(ws.g, pkm.out.w, pkm.out.b, out=ws.t, engine=eng)
    g, b_, e = pk.norms[1]
    ops.add_layernorm(ws.t, ws.x1, g, b_, e, out_f32=ws.x2, out_lp=None if f32 else ws.x2_lp)
    ops.linear(ws.x2_lp, pk.ffn1.w, pk.ffn1.b, relu=True, out=ws.h, engine=eng)
    ops.linear(ws.h, pk.ffn2.w, pk.ffn2.b, out=ws.t, engine=eng)
    g, b_, e = pk.norms[2]
    ops.add_layernorm(ws.t, ws.x2, g, b_, e, out_f32=ws.x, out_lp=None if f32 else ws.x_lp, pos=pos_next,
                      out_pos=ws.xq_lp if pos_next is not None else None)


def bbox_head_ws(mp: MlpPack, ws, refer_in, refer_out) -> None:
    """sigmoid(MLP3(x) + inverse_sigmoid(refer)) (transformer.py:709) from ws.x_lp into refer_out."""
    ops.linear(ws.x_lp, mp.hidden[0].w, mp.hidden[0].b, relu=True, out=ws.bh1, engine=_GEMM_ENGINE)
    ops.linear(ws.bh1, mp.hidden[1].w, mp.hidden[1].b, relu=True, out=ws.bh2, engine=_GEMM_ENGINE)
    ops.box_refine(ws.bh2, mp.last_w, mp.last_b, refer_in, out=refer_out)


# ------------------------------------------------------------------------------------------------
# Whole decoder in one launch (csrc/decoder_cluster.cu): the row-tile-persistent cluster kernel.
# ------------------------------------------------------------------------------------------------
# Opt-in (MOYOLO_CLUSTER_DECODER=1 or TrackEngine(cluster_decoder=True)): measured on B200 the one-launch decoder
# takes 237 us per MOT17 frame against ~200 us for the launch-chained layers it replaces (profiles/README.md,
# "cluster decoder"), so the chain stays the default schedule.
CLUSTER_DECODER = os.environ.get("MOYOLO_CLUSTER_DECODER", "0") == "1"


class ClusterDecoder:
    """Descriptor of moyolo_decoder_cluster_forward for one set of decoder weights: all layers of
    MOTRTransformerDecoder.forward (transformer.py:676-728) incl. box refinement and the last layer's score head in
    ONE kernel. bf16, d_model 256, 8 heads, d_ffn 1024, 3 levels x 4 points, ReLU, nc <= 8."""

    def __init__(self, layer_packs: Sequence[LayerPack], bbox_packs: Sequence[MlpPack], shapes, score_w=None,
                 score_b=None):
        import ctypes as C
        self._C = C
        self.n_layers = len(layer_packs)
        self.shapes = [[int(h), int(w)] for h, w in shapes]
        d = _lib.DecoderCluster()
        pk0 = layer_packs[0]
        d.n_layers, d.d_model, d.n_heads = self.n_layers, pk0.C, pk0.n_heads
        d.d_ffn, d.n_levels, d.n_points = pk0.ffn1.w.shape[0], pk0.msda.n_levels, pk0.msda.n_points
        self._keep = []
        for i, (pk, bp) in enumerate(zip(layer_packs, bbox_packs)):
            L = d.layers[i]
            pairs = dict(wqkv=pk.qkv.w, wo=pk.o.w, woff=pk.msda.offlog.w, wout=pk.msda.out.w, w1=pk.ffn1.w, w2=pk.ffn2.w,
                         wb1=bp.hidden[0].w, wb2=bp.hidden[1].w, bqkv=pk.qkv.b, bo=pk.o.b, boff=pk.msda.offlog.b,
                         bout=pk.msda.out.b, b1=pk.ffn1.b, b2=pk.ffn2.b, bb1=bp.hidden[0].b, bb2=bp.hidden[1].b,
                         wb3=bp.last_w, bb3=bp.last_b, ln1_w=pk.norms[0][0], ln1_b=pk.norms[0][1], ln2_w=pk.norms[1][0],
                         ln2_b=pk.norms[1][1], ln3_w=pk.norms[2][0], ln3_b=pk.norms[2][1])
            for name, t in pairs.items():
                if not t.is_contiguous():
                    raise ValueError(f"ClusterDecoder: {name} must be contiguous")
                setattr(L, name, t.data_ptr())
                self._keep.append(t)
        self.eps = float(layer_packs[0].norms[0][2])
        self.softmax_mode = layer_packs[0].msda.softmax_mode
        flat = [v for hw in self.shapes for v in hw]
        for k, v in enumerate(flat):
            d.value_shapes[k] = v
        d.softmax_mode, d.eps = self.softmax_mode, self.eps
        if score_w is not None:
            d.score_w, d.score_b, d.nc = score_w.data_ptr(), score_b.data_ptr(), score_w.shape[0]
            self._keep += [score_w, score_b]
        self.desc = d
        self._limits = {}

    @staticmethod
    def supports(dt, spec_like) -> bool:
        """Static configuration check (d_model 256, 8 heads, d_ffn 1024, 3 levels x 4 points, bf16, nc <= 8)."""
        return (dt == torch.bfloat16 and spec_like.d_model == 256 and spec_like.n_heads == 8 and
                spec_like.d_ffn == 1024 and spec_like.n_levels == 3 and spec_like.n_points == 4 and
                getattr(spec_like, "nc", 1) <= 8 and _GEMM_ENGINE != _lib.GEMM_SIMT)

    def limits(self, rows_per_tile: int):
        """(co-resident clusters, longest sequence in rows) for the kernel's tile height of 32 rows."""
        if rows_per_tile not in self._limits:
            C = self._C
            mc, kc = C.c_int(0), C.c_int(0)
            _lib.check(_lib.lib().moyolo_decoder_cluster_limits(rows_per_tile, C.byref(mc), C.byref(kc)))
            self._limits[rows_per_tile] = (mc.value, kc.value)
        return self._limits[rows_per_tile]

    def tile_rows(self, rows_pad: int, n_seq: int, max_seq_rows: int) -> int:
        """Tile height (32 rows) if the frame fits the device (every row tile needs its own co-resident cluster and
        a sequence's keys must fit the key / value staging), else 0."""
        m = 32
        max_clusters, kv_cap = self.limits(m)
        if (rows_pad + m - 1) // m + n_seq - 1 <= max_clusters and max_seq_rows <= kv_cap:
            return m
        return 0

    def run(self, x_in, pos, refer0, values, row_offsets, n_seq: int, rows_pad: int, rows_per_tile: int, x_out,
            kv, grid_barrier, refer_out: Sequence[Optional[torch.Tensor]], x_lp_out=None, logits=None, scores=None,
            labels=None, status=None, reset_barrier: bool = True, profile=None) -> None:
        """values: bf16 [B, Lv, n_layers*256] (all layers' value projections); kv: bf16 scratch [2, rows_pad, 512];
        grid_barrier: uint32/int32 [1]; refer_out[i]: fp32 [rows_pad, 4] or None."""
        d = self.desc
        for t in (x_in, pos, refer0, x_out):
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("ClusterDecoder.run: x_in / pos / refer0 / x_out must be contiguous fp32")
        if values.dtype != torch.bfloat16 or values.stride(2) != 1 or kv.numel() < 2 * rows_pad * 512:
            raise ValueError("ClusterDecoder.run: values must be bf16 [B, Lv, n_layers*256], kv bf16 [2, rows_pad, 512]")
        d.x_in, d.pos, d.refer0, d.x_out = x_in.data_ptr(), pos.data_ptr(), refer0.data_ptr(), x_out.data_ptr()
        d.x_lp_out = None if x_lp_out is None else x_lp_out.data_ptr()
        for i in range(8):
            t = refer_out[i] if i < len(refer_out) else None
            d.refer_out[i] = None if t is None else t.data_ptr()
        d.kv, d.values = kv.data_ptr(), values.data_ptr()
        d.value_batch_stride, d.value_pos_stride = values.stride(0), values.stride(1)
        d.row_offsets, d.n_seq, d.rows_pad, d.rows_per_tile = row_offsets.data_ptr(), n_seq, rows_pad, rows_per_tile
        d.grid_barrier, d.reset_barrier = grid_barrier.data_ptr(), 1 if reset_barrier else 0
        d.status = None if status is None else status.data_ptr()
        d.logits = None if logits is None else logits.data_ptr()
        d.scores = None if scores is None else scores.data_ptr()
        d.labels = None if labels is None else labels.data_ptr()
        d.profile = None if profile is None else profile.data_ptr()
        _lib.check(_lib.lib().moyolo_decoder_cluster_forward(self._C.byref(d), torch.cuda.current_stream().cuda_stream))
