#!/bin/bash
# Round-2 multi-GPU lines (run under `gpurun --gpus 8`): MOT17 1/2/4/8, DanceTrack 32 sequences over 8 GPUs
# (BASELINE.json configs[2]), KITTI 2/4/8 sweep (configs[3]); every line with --check-table.
mkdir -p gpurun_out/scale
run() {  # n, tag, extra args
  n=$1; tag=$2; shift 2
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-selection --no-cpu-baseline "$@" > gpurun_out/scale/${tag}_n$n.json 2> gpurun_out/scale/${tag}_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
      bench.py --gpus $n --steps 20 --warmup 5 --no-selection --no-cpu-baseline --check-table "$@" > gpurun_out/scale/${tag}_n$n.json 2> gpurun_out/scale/${tag}_n$n.err
  fi
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/scale/${tag}_n$n.json"))
    print("${tag} n=$n", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "gather_ms", d["run_info"]["final_gather_ms"], d.get("table_check"))
except Exception as e:
    print("${tag} n=$n FAILED", e)
PY
}
NS="${NS:-1 2 4 8}"
for n in $NS; do run $n mot17; done
for n in 1 8; do case " $NS " in *" $n "*) run $n dancetrack --workload DanceTrack --seqs-per-gpu 4;; esac; done
for n in $NS; do run $n kitti --workload KITTI --seqs-per-gpu 4; done
