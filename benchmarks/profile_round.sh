#!/bin/bash
# Profiling pass run under gpurun (one GPU). Outputs land in gpurun_out/; summaries are copied to profiles/ by hand.
#   $1 = tag (e.g. r01b)
TAG=${1:-r01}
mkdir -p gpurun_out
# (1) launch list of the bench command (graph kernel nodes are profiled individually; cold-cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --max-tracks 32 > gpurun_out/${TAG}_launches_bench.log 2>&1
# (2) full capture of the dominant kernels (3 launches each)
for K in msda_gather self_attention gemm_tcgen05 add_layernorm; do
  ncu --set full --clock-control none --import-source on -k regex:${K} -s 40 -c 3 -f -o gpurun_out/${TAG}_${K} \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --max-tracks 32 > gpurun_out/${TAG}_${K}.log 2>&1
done
