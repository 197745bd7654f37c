"""CPU: the `bench.py --impl reference` arm (the reference's CPU path = the oracle port, timed on the host cores)
prints exactly ONE JSON line on stdout with the contract's keys; other ranks of a torchrun launch print nothing."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--cpu-frames", "1", "--workload", "tiny", "--n-detect", "32"],
                          capture_output=True, text=True, env=env, cwd=str(ROOT), timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decoder_frames_per_sec" and d["unit"] == "frames/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["value"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""


def test_both_arms_describe_the_same_config():
    """The driver compares the `config` objects of the two arms for equality: both come from bench.workload_config and
    depend only on the workload flags (run-specific details live in `run_info`)."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from moyolo_b200 import synthetic as syn
    for wl in ("MOT17", "DanceTrack", "KITTI"):
        shapes = [list(s) for s in syn.PYRAMIDS[wl]]
        a = bench.workload_config(argparse.Namespace(workload=wl, n_detect=300, seqs_per_gpu=1, precision="bf16"), shapes)
        b = bench.workload_config(argparse.Namespace(workload=wl, n_detect=300, seqs_per_gpu=4, precision="fp32"), shapes)
        assert a == b and "workload" in a and "model" not in a
