#!/bin/bash
# Profiling pass run under gpurun (one GPU). Outputs land in gpurun_out/; summaries are copied to profiles/ by hand.
#   $1 = tag (e.g. r01g)   $2.. = extra bench.py flags
TAG=${1:-r01}
shift
mkdir -p gpurun_out
BENCH="python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-selection --max-tracks 32 --profiler-range $*"
# (1) launch list of the timed `value` leg only (graph kernel nodes are profiled individually; cold-cache, serialised)
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches_bench.log 2>&1
# (2) full capture of the dominant kernels (a few launches each, from the middle of the timed leg)
for K in msda_gather self_attention gemm_stream gemm_tcgen05; do
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:${K} -s 8 -c 3 -f \
      -o gpurun_out/${TAG}_${K} $BENCH > gpurun_out/${TAG}_${K}.log 2>&1
done
# (3) the gather backward (not part of the frame): two launches at the MOT17 shapes
ncu --set full --clock-control none --import-source on -k regex:msda_backward -c 2 -f -o gpurun_out/${TAG}_msda_backward \
    python -m pytest tests/test_gpu_parity.py -q -k "backward_vs_c_oracle_full_size and MOT17" > gpurun_out/${TAG}_msda_backward.log 2>&1
