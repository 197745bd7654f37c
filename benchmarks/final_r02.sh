#!/bin/bash
# Round-2 closing pass on one B200: full GPU suite, smoke, bench (both arms), fp32 line, launch list, gather capture, sweep.
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 300 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 300 python bench.py > $O/r02_bench.json 2> $O/r02_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 12 --warmup 3 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "reference rc=$?"
timeout 120 python bench.py --precision fp32 --steps 20 --warmup 5 --no-cpu-baseline --no-selection > $O/r02_bench_fp32.json 2> $O/r02_bench_fp32.err; echo "fp32 rc=$?"
BENCH="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-selection --max-tracks 32 --profiler-range"
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/r02_launches_raw.csv $BENCH > $O/launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:msda_gather -s 8 -c 3 -f \
    -o $O/r02_msda_gather $BENCH > $O/msda_gather.log 2>&1; echo "ncu gather rc=$?"
timeout 240 python benchmarks/msda_sweep.py > $O/msda_sweep.log 2>&1; echo "sweep rc=$?"; cp gpurun_out/msda_sweep.json $O/r02_msda_sweep.json 2>/dev/null
python - <<'PY'
import json
for f in ("r02_bench", "r02_bench_fp32", "r02_bench_reference"):
    try:
        d = json.load(open(f"gpurun_out/final/{f}.json"))
        print(f, d.get("value"), d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "launches/frame", d.get("launches_per_frame"),
              "roofline", (d.get("roofline") or {}).get("frac"), "parity", d.get("parity_check"))
    except Exception as e:
        print(f, "FAILED", e)
PY
