#!/usr/bin/env python
"""Warm in-graph latency of every kernel of one decoder frame at the MOT17 shapes (R = 357 query rows,
Lv = 13566, C = 256): each op is captured N times back to back in one CUDA graph (stream-ordered, so
launch i+1 waits for launch i exactly as inside the frame graph) and the replay is timed with CUDA
events. Profiling script, not product code. Writes gpurun_out/op_latency.json."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from moyolo_b200 import _lib, ops, synthetic as syn  # noqa: E402

dev = torch.device("cuda:0")
N = 40


def timed(fn, n=N, reps=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n)
    return round(best * 1e3, 2)


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 357
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    C, H = 256, 8
    shapes = [list(s) for s in syn.PYRAMIDS["MOT17"]]
    Lv = syn.level_sizes(shapes)
    bf = torch.bfloat16
    g = torch.Generator().manual_seed(0)
    rn = lambda *s: torch.randn(*s, generator=g).to(dev)  # noqa: E731
    x32, pos32, res32 = rn(R, C), rn(R, C), rn(R, C)
    xb = x32.to(bf)
    res = {}
    # GEMMs of the layer (M = R)
    for name, Nn, K, odt, relu in (("qk_proj", 512, 256, bf, False), ("v_proj", 256, 256, bf, False),
                                   ("o_proj/out_proj", 256, 256, torch.float32, False),
                                   ("offlog", 288, 256, torch.float32, False), ("ffn1", 1024, 256, bf, True),
                                   ("ffn2", 256, 1024, torch.float32, False), ("bbox_h", 256, 256, bf, True)):
        w = (rn(Nn, K) / K ** 0.5).to(bf)
        b = rn(Nn)
        xin = rn(R, K).to(bf)
        out = torch.empty(R, Nn, dtype=odt, device=dev)
        res[f"gemm_{name}_{R}x{Nn}x{K}"] = timed(lambda: ops.linear(xin, w, b, relu=relu, out=out))
    # value projection for all layers
    feats = rn(S * Lv, C).to(bf)
    wv = (rn(6 * C, C) / 16).to(bf)
    bv = rn(6 * C)
    vout = torch.empty(S * Lv, 6 * C, dtype=bf, device=dev)
    res[f"gemm_value_proj_{S * Lv}x1536x256"] = timed(lambda: ops.linear(feats, wv, bv, out=vout), n=10)
    # fused variants
    wq = (rn(768, 256) / 16).to(bf); bq = rn(768)
    xa, xb2 = rn(R, 256).to(bf), rn(R, 256).to(bf)
    oq = torch.empty(R, 768, dtype=bf, device=dev)
    res["gemm_qkv_dual"] = timed(lambda: ops.linear_dual(xa, xb2, 512, wq, bq, out=oq))
    for K in (256, 1024):
        wl = (rn(256, K) / K ** 0.5).to(bf); bl = rn(256)
        xin = rn(R, K).to(bf)
        rs, ps = rn(R, 256), rn(R, 256)
        gam, bet = rn(256), rn(256)
        o32 = torch.empty(R, 256, device=dev); olp = torch.empty(R, 256, dtype=bf, device=dev); opl = torch.empty(R, 256, dtype=bf, device=dev)
        res[f"gemm_ln_K{K}"] = timed(lambda: ops.linear_add_layernorm(xin, wl, bl, rs, gam, bet, 1e-5, out_f32=o32, out_lp=olp, pos=ps, out_pos=opl))
    # FFN block: two launches vs the one-launch cluster kernel
    for F in (1024, 256):
        w1 = (rn(F, 256) / 16).to(bf); b1 = rn(F)
        w2 = (rn(256, F) / F ** 0.5).to(bf); b2 = rn(256)
        xin = rn(R, 256).to(bf)
        hbuf = torch.empty(R, F, dtype=bf, device=dev)
        rs, ps = rn(R, 256), rn(R, 256)
        gam, bet = rn(256), rn(256)
        o32 = torch.empty(R, 256, device=dev); olp = torch.empty(R, 256, dtype=bf, device=dev); opl = torch.empty(R, 256, dtype=bf, device=dev)

        def two():
            ops.linear(xin, w1, b1, relu=True, out=hbuf)
            ops.linear_add_layernorm(hbuf, w2, b2, rs, gam, bet, 1e-5, out_f32=o32, out_lp=olp, pos=ps, out_pos=opl)
        res[f"ffn_F{F} (2 launches)"] = timed(two)
        res[f"ffn_F{F} fused (1 launch)"] = timed(lambda: ops.ffn_add_layernorm(xin, w1, b1, w2, b2, hbuf, rs, gam, bet, 1e-5,
                                                                                out_f32=o32, out_lp=olp, pos=ps, out_pos=opl))
    # attention
    qkv = rn(R, 3 * C).to(bf)
    per = R // S
    offs = [i * per for i in range(S)] + [R]
    ro = torch.tensor(offs, dtype=torch.int32, device=dev)
    att = torch.empty(R, C, dtype=bf, device=dev)
    res["self_attention"] = timed(lambda: ops.self_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], ro, offs, H, out=att))
    # layernorm
    ga, be = rn(C), rn(C)
    res["add_layernorm(f32+lp+pos)"] = timed(lambda: ops.add_layernorm(x32, res32, ga, be, 1e-5, True, True, bf, pos32))
    # gather
    values = vout.view(S, Lv, 6 * C)
    ol = rn(R, 288)
    refer = torch.rand(R, 1, 4, generator=g).to(dev)
    gout = torch.empty(R, C, dtype=bf, device=dev)
    res["msda_fused"] = timed(lambda: ops.msda_fused(values[:, :, :C], shapes, ol[:, :192], ol[:, 192:], refer, H, 4, S,
                                                     row_offsets=ro if S > 1 else None, out=gout))
    wol = (rn(288, 256) / 16).to(bf); bol = rn(288)
    xq = rn(R, 256).to(bf)
    olo = torch.empty(R, 288, device=dev)

    def two_launch():
        ops.linear(xq, wol, bol, out=olo)
        ops.msda_fused(values[:, :, :C], shapes, olo[:, :192], olo[:, 192:], refer, H, 4, S,
                       row_offsets=ro if S > 1 else None, out=gout)
    res["offlog_gemm+msda_fused (2 launches)"] = timed(two_launch)
    if ops.proj_fused_supported(bf, H, 32, 3, 4, R):
        res["msda_proj_fused (1 launch)"] = timed(lambda: ops.msda_proj_fused(
            values[:, :, :C], shapes, xq, wol, bol, refer, H, 4, S, row_offsets=ro if S > 1 else None, out=gout))
    # heads
    w3, b3 = rn(4, C) * 0.05, rn(4)
    refb = torch.rand(R, 4, generator=g).to(dev)
    res["box_refine"] = timed(lambda: ops.box_refine(xb, w3, b3, refb))
    ws, bs = rn(1, C) * 0.1, rn(1)
    res["score_head"] = timed(lambda: ops.score_head(xb, ws, bs))
    res["add_cast"] = timed(lambda: ops.add_cast(x32, pos32, bf))
    res["sigmoid"] = timed(lambda: ops.sigmoid(refb))
    res["pos2posemb"] = timed(lambda: ops.pos2posemb(refb))
    # empty-ish kernel: launch floor of a dependent chain inside a graph
    z = torch.zeros(32, device=dev)
    res["floor(sigmoid 32 elems)"] = timed(lambda: ops.sigmoid(z))
    out = {"rows": R, "seqs": S, "us_per_launch": res}
    print(json.dumps(out, indent=1))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"op_latency_R{R}_S{S}.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
