#!/usr/bin/env python
"""bench.py — decoder frames/sec of the DecoderTracker hot path on B200 (contract: see DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU path (oracle port)

A "step" is one frame of the hot path for every lock-step sequence on the GPU: query assembly ->
6-layer deformable decoder -> box/score heads -> track-query update -> result rows. The N=1 workload is
BASELINE.json configs[1]: a MOT17-shaped synthetic sequence (1088x608 -> pyramid (76,136),(38,68),
(19,34)), 300 detect queries + carried track queries, bf16. At N>1 every rank tracks its own
sequence(s) (weak scaling, no collective in the frame loop; one NCCL all_gather of the track rows at
the end, inside the timed region).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "decoder_frames_per_sec"
UNIT = "frames/s"

# stdout carries exactly ONE line (the JSON result). Libraries write banners to fd 1 (e.g. "NCCL version ..." when
# NCCL_DEBUG=VERSION): everything else is sent to stderr and the real stdout is restored for the result line.
_REAL_STDOUT = None


def _quiet_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict) -> None:
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="moyolo", choices=["moyolo", "reference"])
    ap.add_argument("--workload", default="MOT17", choices=["MOT17", "DanceTrack", "KITTI", "C1", "tiny"])
    ap.add_argument("--seqs-per-gpu", type=int, default=1, help="lock-step sequences per GPU")
    ap.add_argument("--n-detect", type=int, default=300)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-frames", type=int, default=12, help="frames of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--max-tracks", type=int, default=160, help="pre-capture frame graphs up to this many tracks/seq")
    ap.add_argument("--host-lag", type=int, default=2, help="frames the enqueueing host may run ahead (TrackEngine host_lag)")
    ap.add_argument("--no-selection", action="store_true", help="skip the query-selection (f1) leg")
    ap.add_argument("--check-table", action="store_true",
                    help="rank 0 re-runs every sequence of the job on one GPU and compares it with the gathered table")
    ap.add_argument("--table-rows-per-frame", type=int, default=256,
                    help="bound on tracked objects per frame and sequence (fixed capacity of the final gather)")
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed `value` leg (ncu --profile-from-start off)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms DURING the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def wait_first(self, timeout: float = 15.0):
        """Block until the first sample has arrived: nvidia-smi's start-up enumerates every GPU of the box through the
        driver, which stalls running work on ALL of them for a few hundred ms (seen as 5-25 % slower first timed
        regions on random ranks at N=8); the periodic samples afterwards do not."""
        t0 = time.time()
        while self.proc is not None and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.05)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx = max(mx, float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ workload
def workload_config(args, shapes) -> dict:
    """The `config` object: identical in both arms (the driver compares them for equality); everything that
    describes a particular run goes to `run_info`."""
    return {"workload": f"{args.workload}-shaped synthetic sequence, pyramid {shapes}, {args.n_detect} detect queries + "
                        f"carried track queries, 6-layer decoder d=256 h=8 L=3 P=4, track update; planted births / "
                        f"deaths (moyolo_b200.synthetic.TRACKING_WORKLOADS)",
            "baseline_config": "BASELINE.json configs[1]" if args.workload == "MOT17" else
                               ("BASELINE.json configs[2]" if args.workload == "DanceTrack" else
                                ("BASELINE.json configs[3]" if args.workload == "KITTI" else "parity-test case"))}


def make_frames(args, syn, spec, plant, shapes, device, seq_seed, n_frames, lp):
    g = syn.PlantedSequenceGenerator(syn.SequenceSpec(args.workload, n_frames, args.n_detect, seq_seed, shapes=shapes),
                                     spec, plant, device)
    frames = []
    for _ in range(n_frames):
        f, de, dr = g.next_frame()
        frames.append((f.to(lp).contiguous(), de.contiguous(), dr.contiguous()))
    return frames


def frame_floor(spec, Lv, rows, n_detect, peaks, ms_per_frame):
    """Speed-of-light time of one frame at S=1: max(flops / measured bf16 peak, compulsory bytes /
    measured HBM bandwidth) against the measured frame time."""
    C, F, nl, H = spec.d_model, spec.d_ffn, spec.n_layers, spec.n_heads
    LP3 = spec.n_heads * spec.n_levels * spec.n_points * 3
    per_row_layer = 2 * C * 3 * C + 4 * rows * C + 2 * C * C + 2 * C * LP3 + 2 * C * C + 4 * C * F + 2 * (2 * C * C + 4 * C)
    gather = 2 * 4 * spec.n_levels * spec.n_points * C      # bilinear corners x weighted sum per row and layer
    flops = 2 * Lv * C * nl * C + nl * rows * (per_row_layer + gather)
    t_act = max(rows - n_detect, 0)
    flops += t_act * (2 * C * 3 * C + 4 * t_act * C + 2 * C * C + 4 * 2 * C * spec.qim_hidden)   # QIM
    w_bytes = 2 * (nl * (3 * C * C + C * C + LP3 * C + C * C + 2 * C * F + 2 * C * C) + nl * C * C +
                   3 * C * C + C * C + 4 * C * spec.qim_hidden)
    byts = Lv * C * 2 + 2 * Lv * nl * C * 2 + w_bytes + rows * C * 4 * 4
    tf = float(peaks.get("bf16_tflops_sustained", 1400.0)) * 1e12
    bw = float(peaks.get("hbm_gbs", 6650.0)) * 1e9
    t_floor = max(flops / tf, byts / bw)
    return {"flops": int(flops), "bytes": int(byts), "t_floor_us": round(t_floor * 1e6, 2),
            "t_measured_us": round(ms_per_frame * 1e3, 2), "frac": round(t_floor * 1e3 / ms_per_frame, 4),
            "peaks": "MEASURED_PEAKS.json bf16_tflops_sustained / hbm_gbs" if peaks else "fallback 1400 TF/s, 6650 GB/s"}


def gather_bytes(B, Lv, C, R, H, L, P, s_v, proj_fused=False):
    """Compulsory bytes of one deformable-gather launch: value once + raw offsets and
    logits + reference boxes + output. With the offsets|logits projection fused into the kernel the fp32
    [R, H*L*P*3] operand is replaced by the bf16 query rows [R, C] and the projection weight + bias."""
    common = B * Lv * C * s_v + R * 4 * 4 + R * C * s_v
    if proj_fused:
        return common + R * C * 2 + H * L * P * 3 * (C * 2 + 4)
    return common + R * H * L * P * 3 * 4


def batched_gather_roofline(ops, syn, shapes, spec, device, peak, S=16, Q=300):
    """The same gather in the regime where one launch carries enough bytes to be bandwidth-bound: S lock-step
    sequences (what a GPU tracking S videos gathers per layer), value laid out as in the frame (one layer's
    column slice of the all-layers value tensor). warm = back-to-back launches in a CUDA graph with the value
    slice L2-resident where it fits (how it runs right after the value_proj GEMM); cold = a 256 MB L2 flush
    before every launch."""
    H, L, P, C = spec.n_heads, spec.n_levels, spec.n_points, spec.d_model
    Lv = syn.level_sizes(shapes)
    g = torch.Generator().manual_seed(123)
    R = S * Q
    values = torch.randn(S, Lv, 2 * C, generator=g).to(device, torch.bfloat16)   # two layers' slices side by side
    value = values[:, :, :C]
    ol = torch.randn(R, H * L * P * 3, generator=g).to(device)
    refer = torch.cat([torch.rand(R, 1, 2, generator=g), torch.rand(R, 1, 2, generator=g) * 0.3 + 0.02], -1).to(device)
    ro = torch.arange(0, R + 1, Q, dtype=torch.int32, device=device)
    out = torch.empty(R, C, dtype=torch.bfloat16, device=device)
    n_off = H * L * P * 2
    fn = lambda: ops.msda_fused(value, shapes, ol[:, :n_off], ol[:, n_off:], refer, H, P, S, row_offsets=ro, out=out)  # noqa: E731
    fn()
    torch.cuda.synchronize()
    N = 20
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(N):
            fn()
    warm = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
        warm = min(warm, a.elapsed_time(b) / N)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    cold = []
    for _ in range(7):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        cold.append(a.elapsed_time(b))
    cold.sort()
    nbytes = gather_bytes(S, Lv, C, R, H, L, P, 2)
    w_gbs, c_gbs = nbytes / (warm * 1e-3) / 1e9, nbytes / (cold[len(cold) // 2] * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "msda_gather_pair_kernel<bf16,4,3>", "sequences": S, "rows": R,
            "algorithmic_bytes_per_launch": int(nbytes), "warm_us": round(warm * 1e3, 2),
            "cold_us": round(cold[len(cold) // 2] * 1e3, 2), "achieved_warm": round(w_gbs, 1),
            "achieved_cold": round(c_gbs, 1), "peak": peak, "unit": "GB/s", "frac_warm": round(w_gbs / peak, 4),
            "frac_cold": round(c_gbs / peak, 4),
            "note": "same gather, 16 lock-step sequences per launch (bandwidth regime); warm = value slice as left in "
                    "L2 by the preceding launch, cold = L2 flushed before the launch"}


def run_moyolo(args):
    from moyolo_b200 import _lib, ops, synthetic as syn
    from moyolo_b200 import sharding
    from moyolo_b200.tracker import DecoderWeights, TrackEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    host_binding = (sharding.bind_host_to_gpu(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
                    if world > 1 else None)   # own cores per rank, pinned buffers local to the GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert _lib.lib().moyolo_device_supported() == 1, "libmoyolo_b200 targets sm_100a (B200) only"

    spec, shapes, sd, plant = syn.tracking_workload(args.workload, 0)
    lp = torch.bfloat16 if args.precision == "bf16" else torch.float32
    S, K, Wm = args.seqs_per_gpu, args.steps, args.warmup
    weights = DecoderWeights(sd, spec, device, args.precision)
    eng = TrackEngine(sd, spec, shapes, device, args.precision, args.n_detect, S, weights=weights, host_lag=args.host_lag)
    n_graphs = eng.prepare(args.max_tracks)  # frame graphs captured up front, none inside the timed region

    # frames resident in HBM before the timed region: K distinct frames per sequence slot. Every sequence has K
    # frames, so the launcher's LPT assignment deals them round-robin (rank r tracks r, r + world, ...); sequence q is
    # generated from seed 1 + q on whichever rank owns it.
    assign = sharding.lpt_assign([K] * (world * S), world)
    mine = sorted(assign[rank])           # = the order of the launcher's lock-step group
    seqs = [make_frames(args, syn, spec, plant, shapes, device, 1 + q, K, lp) for q in mine]
    warm = make_frames(args, syn, spec, plant, shapes, device, 9999, max(Wm, 1), lp)

    def batch(t, src):
        return (torch.stack([src[s][t][0] for s in range(S)]), torch.stack([src[s][t][1] for s in range(S)]),
                torch.stack([src[s][t][2] for s in range(S)]))

    dev_batches = [batch(t, seqs) for t in range(K)]
    del seqs
    warm_dev = [tuple(torch.stack([w[k]] * S) for k in range(3)) for w in warm]
    host_batches = [tuple(x.cpu().pin_memory() for x in b) for b in dev_batches]
    warm_host = [tuple(x.cpu().pin_memory() for x in b) for b in warm_dev]
    feat_bytes = dev_batches[0][0].numel() * dev_batches[0][0].element_size()
    in_bytes = sum(x.numel() * x.element_size() for x in dev_batches[0])
    gather_cap = K * S * args.table_rows_per_frame   # fixed gather capacity: one collective, no count exchange
    job = [{"n_frames": K} for _ in range(world * S)]
    assert len(mine) == S

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def warmup(batches, want_rows):
        """W untimed frames through the IDENTICAL code path as the leg that follows (same input residency, same
        submit/collect calls, the final gather), then drop all tracks and the track table."""
        eng.reset()
        for t in range(Wm):
            eng.submit(*batches[t % len(batches)], want_rows=want_rows)
            if want_rows and t > 0:
                eng.collect(t - 1)
        sharding.gather_track_rows(eng.track_table().clone(), capacity=gather_cap).count()
        eng.reset()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    # the enqueue loop runs ~10 us ahead of the device per launch: a generation-2 garbage collection in the middle of
    # a 20-frame region is a millisecond stall on one rank (seen as 10-20 % outliers at N=8). Freeze what exists and
    # keep the collector off while timing, as a serving loop would.
    gc.collect()
    gc.freeze()
    gc.disable()

    # ---------------- leg 1: `value` -- inputs resident in HBM, through the sharding launcher ----------------
    # sharding.run_sharded drives the real TrackEngine: the host only enqueues (frame t+1 is submitted while frame t
    # runs), tracked objects are appended to a device-resident table by the frame graph itself and gathered with ONE
    # NCCL all_gather_into_tensor at the end, inside the timed region.
    # With fewer than 100 steps the K-step region is run several times (reset + warm-up in between, like the e2e
    # leg) and the MEDIAN repeat is reported with the spread and every repeat's time: on a shared 8-GPU box one rank
    # in eight sees a 0.5-2 ms stall in roughly every other 5 ms region (profiles/r02_scaling.md).
    value_reps = 1 if K >= 100 else 5
    mid = {}

    def engine_for(n):
        assert n == S
        return eng

    def value_pass():
        warmup(warm_dev, False)
        launches0 = eng.launches
        barrier()
        e0, e_mid, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))

        def mark(local):   # this rank's table is complete here: everything after it is the gather
            e_mid.record()
            mid["local"] = local

        if args.profiler_range:
            torch.cuda.cudart().cudaProfilerStart()
        e0.record()
        gathered = sharding.run_sharded(engine_for, job, rank, world, max_in_flight=S,
                                        batch_fn=lambda grp, t: dev_batches[t], rows_per_frame=args.table_rows_per_frame,
                                        sort=False, sync_inputs=False, before_gather=mark)
        e1.record()
        if args.profiler_range:
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1), e0.elapsed_time(e_mid)], device=device)
        if world > 1:
            allr = torch.zeros(world, 2, device=device)
            dist.all_gather_into_tensor(allr.view(-1), t)
            t = allr
        t = t.view(-1, 2).cpu()
        return {"ms": float(t[:, 0].max()), "gather_ms": e_mid.elapsed_time(e1), "gathered": gathered,
                "launches": eng.launches - launches0, "frame_loop_ms_by_rank": [round(float(v), 3) for v in t[:, 1]]}

    passes = [value_pass() for _ in range(value_reps)]
    order = sorted(range(value_reps), key=lambda i: passes[i]["ms"])
    best = passes[order[value_reps // 2]]          # the median repeat
    value_ms_all = [round(p["ms"], 3) for p in passes]
    ms_total, ms_gather, frames_ms_by_rank = best["ms"], best["gather_ms"], best["frame_loop_ms_by_rank"]
    gathered = passes[-1]["gathered"]               # (the engine's buffers hold the last repeat)
    launches = best["launches"]
    table = gathered.rows(sort=True)
    n_rows_table = int(table.shape[0])
    ab = torch.tensor([float(eng.aborts)], device=device)
    if world > 1:
        dist.all_reduce(ab)
    aborts_value = int(ab.item())      # summed over ranks and repeats
    local_table = mid["local"]()
    per_frame = torch.bincount(local_table[:, 1].long(), minlength=K).float() if local_table.shape[0] else torch.zeros(K)
    tracks_seen = [float(v) for v in per_frame.cpu().tolist()]

    # optional: the gathered table of every sequence of the job equals a single-GPU run of that sequence
    table_check = None
    if args.check_table and rank == 0:
        ok = 0
        chk = TrackEngine(sd, spec, shapes, device, args.precision, args.n_detect, S, weights=weights)
        for r in range(world):
            theirs = sorted(assign[r])
            fr = [make_frames(args, syn, spec, plant, shapes, device, 1 + q, K, lp) for q in theirs]
            chk.reset()
            chk.set_seq_ids(theirs)
            for t in range(K):
                chk.submit(*batch(t, fr), want_rows=False, sync_inputs=True)
            ref_tab = sharding._sort_rows(chk.track_table().clone())
            for q in theirs:
                a, b = ref_tab[ref_tab[:, 0] == q], table[table[:, 0] == q]
                ok += int(a.shape == b.shape and torch.equal(a, b))
        table_check = {"sequences": world * S, "bit_equal_to_single_gpu_run": ok}
        del chk

    # ---------------- leg 2: roofline of the deformable gather (rank 0, instrumented re-run) ---------
    # The same frames are replayed through a second engine whose frame GRAPHS carry a pair of external CUDA-event
    # record nodes around every gather launch (6 per frame): the kernels run inside the captured frame exactly as in
    # the timed legs and the event pairs give the device time of each gather launch.
    roof = roof_batched = None
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    if rank == 0:
        cur_pairs = []

        def pre(B, Lv, C, R, H, L, P, s_v):
            if not torch.cuda.is_current_stream_capturing():
                return
            a = torch.cuda.Event(enable_timing=True, external=True)
            a.record()
            cur_pairs.append([a, None])

        def post():
            if not torch.cuda.is_current_stream_capturing():
                return
            b = torch.cuda.Event(enable_timing=True, external=True)
            b.record()
            cur_pairs[-1][1] = b

        REP = 8  # launches per event pair (identical, idempotent) so the ~5 us of event-node overhead is amortised
        eng2 = TrackEngine(sd, spec, shapes, device, args.precision, args.n_detect, S, weights=weights,
                           gather_probe=ops.GatherProbe(pre, post, REP))
        eng2.prepare(args.max_tracks)
        n_l = spec.n_layers
        pairs_of = {key: cur_pairs[i * n_l:(i + 1) * n_l] for i, key in enumerate(eng2._plans.keys())}
        g_ms, n_pairs, nbytes = 0.0, 0, 0
        n_inst = min(K, 60)
        pf = ops.proj_fused_supported(lp, spec.n_heads, spec.d_model // spec.n_heads, spec.n_levels, spec.n_points,
                                      S * (args.n_detect + args.max_tracks))
        for t in range(n_inst):
            T_in = sum(eng2.n_tracks_host())
            eng2.submit(*dev_batches[t], want_rows=False)
            eng2.drain()
            torch.cuda.synchronize()
            if t < 3:  # first frames warm the instrumented graphs
                continue
            p = eng2._last_plan
            for a, b in pairs_of.get((p.rows_pad, p.slot), []):
                g_ms += a.elapsed_time(b) / REP
                n_pairs += 1
                nbytes += gather_bytes(S, eng2.Lv, spec.d_model, T_in + S * args.n_detect, spec.n_heads, spec.n_levels,
                                       spec.n_points, 2 if args.precision == "bf16" else 4, proj_fused=pf)
        peak = float(peaks.get("hbm_gbs", 6650.0))
        ach = nbytes / (g_ms * 1e-3) / 1e9 if g_ms > 0 else 0.0
        traffic = None
        tf = ROOT / "profiles" / "gather_traffic.json"   # dram bytes per launch from the committed ncu --set full capture
        if tf.exists():
            traffic = json.loads(tf.read_text()).get(f"{args.workload}_S{S}_{args.precision}")
        kname = ("msda_gather_proj_kernel<4,3> (gather with the offsets|logits projection fused in)" if pf else
                 (f"msda_gather_pair_kernel<{args.precision},4,3>" if S * args.n_detect >= 1024 else
                  f"msda_gather_kernel<{args.precision},32,fused>"))
        roof = {"bound": "hbm", "kernel": kname, "achieved": round(ach, 1),
                "peak": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": traffic,
                "launches_timed": n_pairs, "avg_launch_us": round(g_ms * 1e3 / max(n_pairs, 1), 3),
                "algorithmic_bytes_per_launch": int(nbytes / max(n_pairs, 1)),
                "note": "compulsory bytes (value slice once + query rows / projection weights (or raw offsets/logits) + "
                        "refs + out for the frame's real rows, SURVEY.md 8(d)) / device time between external CUDA-event "
                        f"nodes placed around each gather inside the captured frame graph (6 per frame; each event pair "
                        f"brackets {REP} identical back-to-back launches of that gather and the time is divided by it, "
                        "because a pair of event nodes alone costs ~5 us); `roofline_batched` is the same gather where a "
                        "launch carries enough bytes to be bandwidth-bound"}
        if S == 1 and args.precision == "bf16":
            roof_batched = batched_gather_roofline(ops, syn, shapes, spec, device, peak)
        del eng2

    # ---------------- leg 3: `e2e` -- host buffers, H2D + D2H inside the timed region ----------------
    # Every frame: pinned host inputs -> device (copy stream, overlapping the previous frame's compute), the frame,
    # and ONE device->host copy of its packed result rows [rows, 8] which the host then reads. Warm-up goes through
    # the same pinned-host path. With fewer than 100 steps the K-step sequence is repeated (reset in between) and
    # the MEDIAN repeat is reported together with the spread.
    reps = 1 if K >= 100 else min(9, max(3, -(-100 // K)))
    e2e_ms, e2e_local, host_checksum, d2h_bytes = [], [], 0.0, 0
    gpu_rows = []   # host copies of the first cpu_frames frames' result rows (sequence 0) for the parity check
    for rep in range(-1, reps):   # pass -1 is untimed: it keeps host copies of the result rows for the parity check
        warmup(warm_host, True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        host_checksum, d2h_bytes = 0.0, 0
        e0.record()
        for t in range(K):
            eng.submit(*host_batches[t], want_rows=True)
            if t > 0:  # read frame t-1 on the host while frame t runs
                outs = eng.collect(t - 1)
                for o in outs:
                    host_checksum += float(o["scores"].sum()) + float((o["ids"] >= 0).sum())
                if rep == -1 and t - 1 < args.cpu_frames:
                    gpu_rows.append({k: v.clone() for k, v in outs[0].items()})
        outs = eng.collect(K - 1)
        for o in outs:
            host_checksum += float(o["scores"].sum()) + float((o["ids"] >= 0).sum())
            d2h_bytes += o["ids"].shape[0] * 8 * 4
        d2h_bytes += (S + 8) * 4
        e1.record()
        if rep == -1 and K - 1 < args.cpu_frames:
            gpu_rows.append({k: v.clone() for k, v in outs[0].items()})
        barrier()
        ms2 = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        if rep >= 0:
            e2e_ms.append(float(ms2.item()))
            e2e_local.append(e0.elapsed_time(e1))
    e2e_sorted = sorted(e2e_ms)
    by_rank = torch.tensor([sorted(e2e_local)[len(e2e_local) // 2]], device=device)
    if world > 1:
        allr = torch.zeros(world, device=device)
        dist.all_gather_into_tensor(allr, by_rank)
        by_rank = allr
    e2e_by_rank = [round(float(v), 3) for v in by_rank.cpu().tolist()]
    ms_e2e = e2e_sorted[len(e2e_sorted) // 2]

    gc.enable()
    clocks = sampler.stop() if rank == 0 else None
    frames_total = K * S * world
    line = None
    if rank == 0:
        rows_mean = args.n_detect + sum(tracks_seen) / max(len(tracks_seen), 1) / S
        line = {
            "metric": METRIC, "value": round(frames_total / (ms_total * 1e-3), 2), "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": round(ms_total / K, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic",
            "config": workload_config(args, shapes),
            "run_info": {"sequences_per_gpu": S, "frames_per_sequence": K, "sequences_total": world * S,
                         "queries_per_frame_mean": round(rows_mean, 1),
                         "tracks_carried_max": max(tracks_seen) if tracks_seen else 0, "track_rows_gathered": n_rows_table,
                         "host_binding": host_binding, "e2e_median_ms_by_rank": e2e_by_rank,
                         "frame_loop_ms_by_rank": frames_ms_by_rank,
                         "value_repeats": value_reps, "value_ms_per_repeat": value_ms_all,
                         "frame_loop_ms_by_rank_per_repeat": [p["frame_loop_ms_by_rank"] for p in passes],
                         "value_spread": round((max(value_ms_all) - min(value_ms_all)) / ms_total, 4),
                         "final_gather_ms": round(ms_gather, 3), "gather": "one all_gather_into_tensor, fixed capacity "
                         f"{gather_cap} rows per rank, merged by offset on the device, no host sync",
                         "launcher": "moyolo_b200.sharding.run_sharded (LPT assignment, lock-step groups)",
                         "parallelism": f"sequence-sharded x{world}", "cuda_graphs_precaptured": n_graphs,
                         "host_pipeline": f"host enqueues up to {args.host_lag} frames ahead of the newest result it has read "
                                          "(speculative padded size, device-side abort + re-run)",
                         "speculation_aborts": int(aborts_value), "e2e_host_checksum": round(host_checksum, 3),
                         "e2e_repeats": reps, "e2e_ms_per_repeat": [round(x, 3) for x in e2e_ms],
                         "l2": f"inputs larger than L2: {K} distinct frame buffers of {feat_bytes / 1e6:.1f} MB cycled "
                               f"({K * in_bytes / 1e9:.2f} GB per rank)"},
            "e2e": {"value": round(frames_total / (ms_e2e * 1e-3), 2), "unit": UNIT,
                    "h2d_bytes_per_step": int(in_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                    "ms_per_step": round(ms_e2e / K, 4), "repeats": reps,
                    "spread": round((e2e_sorted[-1] - e2e_sorted[0]) / ms_e2e, 4)},
            "gpu_launches": int(launches),
            "launches_per_frame": round(launches / max(K, 1), 1),
            "roofline": roof,
            "roofline_batched": roof_batched,
            "clocks": clocks,
        }
        if table_check is not None:
            line["table_check"] = table_check
        if S == 1:
            line["frame_floor"] = frame_floor(spec, eng.Lv, rows_mean, args.n_detect, peaks, ms_total / K)
        cpu_src = [tuple(x[0].float().cpu() for x in b) for b in dev_batches[:args.cpu_frames]]
        if world == 1 and not args.no_selection:
            del dev_batches, host_batches
            torch.cuda.empty_cache()
            line["with_query_selection"] = selection_leg(args, spec, syn, shapes, device, S, min(K, 200), Wm)
        if not args.no_cpu_baseline and world == 1:
            base, recs = cpu_baseline(args, sd, spec, shapes, cpu_src)
            line["cpu_baseline"] = base
            line["parity_check"] = parity_check(gpu_rows, recs, 2e-2 if args.precision == "bf16" else 2e-5)
            line["gpu_eager_baseline"] = gpu_eager_baseline(args, sd, spec, shapes, cpu_src, device)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        _emit(line)


def parity_check(gpu_rows, recs, margin):
    """GPU result rows (e2e leg, sequence 0, first frames) against the oracle port's rows on the same inputs:
    track ids must be identical on every frame whose oracle scores keep `margin` to the 0.4 / 0.5 thresholds."""
    import numpy as np
    n = min(len(gpu_rows), len(recs))
    equal = excluded = 0
    max_box = max_score = 0.0
    alive = True
    for t in range(n):
        r, g = recs[t], gpu_rows[t]
        s = r["scores"]
        if not alive or float(np.minimum(np.abs(s - 0.4), np.abs(s - 0.5)).min()) < margin:
            alive = False       # after a marginal score the two trajectories may legitimately diverge
            excluded += 1
            continue
        ids = g["ids"].numpy()
        if ids.shape == r["ids"].shape and np.array_equal(ids, r["ids"]):
            equal += 1
            max_box = max(max_box, float(np.abs(g["boxes"].numpy() - r["boxes"]).max()))
            max_score = max(max_score, float(np.abs(g["scores"].numpy() - s).max()))
    return {"frames": n, "ids_equal_frames": equal, "excluded_by_margin": excluded, "max_box_err": round(max_box, 6),
            "max_score_err": round(max_score, 6), "tracks_in_last_frame": int(recs[n - 1]["n_tracks_in"]) if n else 0,
            "against": "oracle port (oracle/tracker_port.py) on the same first frames of sequence 0"}


def gpu_eager_baseline(args, sd, spec, shapes, frames_cpu, device):
    """The reference's modules as they would run after `.cuda()`: the oracle port's PyTorch ops executed eagerly on
    the same B200 (fp32, and under torch.autocast(bfloat16)), ID assignment on the host as in the reference."""
    from oracle.tracker_port import track_sequence_port
    out = {"unit": UNIT, "kind": "port on cuda (eager PyTorch ops, F.grid_sample core)", "frames": len(frames_cpu)}
    fr = [tuple(x.to(device) for x in f) for f in frames_cpu]
    for name, ac in (("fp32", None), ("autocast_bf16", torch.bfloat16)):
        run = lambda f: track_sequence_port(sd, f, shapes, spec.n_heads, spec.n_levels, spec.n_points, spec.n_layers,  # noqa: E731
                                            spec.nc, device=device, autocast_dtype=ac)
        try:
            run(fr[:2])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run(fr)
            torch.cuda.synchronize()
            out[name] = round(len(fr) / (time.perf_counter() - t0), 2)
        except Exception as e:  # a baseline must never take the bench line down
            out[name] = f"failed: {type(e).__name__}: {e}"[:160]
    return out


# ------------------------------------------------------------------------------------------ f1 leg
def selection_leg(args, spec, syn, shapes, device, S, K, Wm):
    """The same frame with the encoder-side query selection inside the frame graph: the step
    starts from the neck's channels-last maps (1x1 conv + BN, enc_output + scores over all Lv positions, top-k,
    box head + anchors) instead of from ready-made feats / detect queries. Rank-local, no gather."""
    from moyolo_b200.selector import QuerySelector
    from moyolo_b200.tracker import TrackEngine
    ch = (256, 512, 512)
    lp = torch.bfloat16 if args.precision == "bf16" else torch.float32
    sd2 = dict(syn.make_decoder_state(spec, 0))   # detect queries come from the selection here: no planted channels
    sd2.update(syn.make_selector_state(spec, ch, 0))
    g = torch.Generator(device=device).manual_seed(4242)
    cur = [torch.randn(S, h, w, c, generator=g, device=device) for (h, w), c in zip(shapes, ch)]
    frames = []
    for _ in range(K):   # temporally coherent maps so that tracks persist, appear and die
        cur = [m + 0.05 * torch.randn(m.shape, generator=g, device=device) for m in cur]
        frames.append(tuple(m.to(lp).contiguous() for m in cur))

    def engine(state):
        sel = QuerySelector(state, spec, shapes, ch, device, args.precision, args.n_detect, S)
        return TrackEngine(state, spec, shapes, device, args.precision, args.n_detect, S, selector=sel)

    eng = engine(sd2)
    out = eng.step(*frames[0])[0]
    sd2 = syn.calibrate_score_bias(sd2, out["logits"], spec, 0.035)
    del eng
    eng = engine(sd2)
    eng.prepare(args.max_tracks)
    res = {}
    host = [tuple(x.cpu().pin_memory() for x in f) for f in frames]
    for name, src, rows in (("value", frames, False), ("e2e", host, True)):
        eng.reset()
        for t in range(Wm):
            eng.submit(*src[t % len(src)], want_rows=rows)
            if rows and t > 0:
                eng.collect(t - 1)
        eng.reset()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        chk = 0.0
        e0.record()
        for t in range(K):
            eng.submit(*src[t], want_rows=rows)
            if rows and t > 0:
                chk += float(eng.collect(t - 1)[0]["scores"].sum())
        eng.drain()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        res[name] = {"value": round(K * S / (ms * 1e-3), 2), "ms_per_step": round(ms / K, 4)}
    res["unit"] = UNIT
    res["h2d_bytes_per_step"] = int(sum(x.numel() * x.element_size() for x in frames[0]))
    res["tracks_carried_end"] = eng.n_tracks_host()
    res["note"] = ("frame = input_proj of the neck maps + query selection (top-k of all Lv positions) + the decoder frame "
                   "(the selection of frame t+1 runs as its own graph on a side stream while frame t decodes); "
                   "value: maps resident in HBM, e2e: pinned host maps, H2D inside the timed region")
    return res


# ------------------------------------------------------------------------------------------ CPU arms
def cpu_baseline(args, sd, spec, shapes, frames_dev, n_frames=None):
    """The oracle port (reference algorithm restated in PyTorch CPU ops, oracle/torch_port.py +
    oracle/tracker_port.py) timed on this box's host cores over the first frames of the same sequence. Returns
    (baseline dict, the port's per-frame records for the parity check)."""
    from oracle.tracker_port import track_sequence_port
    n = min(n_frames or args.cpu_frames, len(frames_dev))
    torch.set_num_threads(os.cpu_count() or 1)
    frames = [tuple(x.float().cpu() for x in frames_dev[t]) for t in range(n)]
    sd_cpu = {k: v.float().cpu() for k, v in sd.items()}
    track_sequence_port(sd_cpu, frames[:1], shapes, spec.n_heads, spec.n_levels, spec.n_points, spec.n_layers, spec.nc)
    t0 = time.perf_counter()
    recs = track_sequence_port(sd_cpu, frames, shapes, spec.n_heads, spec.n_levels, spec.n_points, spec.n_layers, spec.nc)
    dt = time.perf_counter() - t0
    return {"value": round(n / dt, 3), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"first {n} frames of sequence 0 (same weights/inputs, fp32, {torch.get_num_threads()} threads)"}, recs


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path. The reference is Python and does
    not exist on the GPU box, so this times the oracle port (pinned to the reference by tests/golden)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from moyolo_b200 import synthetic as syn
    from oracle.tracker_port import track_sequence_port
    torch.set_num_threads(os.cpu_count() or 1)
    spec, shapes, sd, plant = syn.tracking_workload(args.workload, 0)
    K, Wm = args.steps, args.warmup
    g = syn.PlantedSequenceGenerator(syn.SequenceSpec(args.workload, K, args.n_detect, 1, shapes=shapes), spec, plant, "cpu")
    first = tuple(t.clone() for t in g.next_frame())
    warm = [first] * max(Wm, 1)
    track_sequence_port(sd, warm[:Wm], shapes, spec.n_heads, spec.n_levels, spec.n_points, spec.n_layers, spec.nc)
    # bounded sample: at most ~3 minutes of CPU work; a step is one frame of the same workload
    budget_s, done, t_total = 170.0, 0, 0.0
    frames = [first]
    chunk = 10
    while done < K and t_total < budget_s:
        n = min(chunk, K - done)
        while len(frames) < done + n:
            frames.append(tuple(t.clone() for t in g.next_frame()))
        # note: each chunk restarts from an empty track set (bounded sample), tracks ramp up inside it
        t0 = time.perf_counter()
        track_sequence_port(sd, frames[done:done + n], shapes, spec.n_heads, spec.n_levels, spec.n_points,
                            spec.n_layers, spec.nc)
        t_total += time.perf_counter() - t0
        done += n
    fps = done / t_total
    cores = torch.get_num_threads()
    sample = f"{done} of {K} frames (chunks of {chunk}, fp32, {cores} threads)"
    line = {"impl": "reference", "metric": METRIC, "value": round(fps, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": K, "warmup": Wm, "ms_per_step": round(1e3 / fps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, shapes),
            "cpu_baseline": {"value": round(fps, 3), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(fps, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


if __name__ == "__main__":
    a = parse()
    _quiet_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_moyolo(a)
