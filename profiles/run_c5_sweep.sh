#!/bin/bash
# BASELINE config 5 (correspondence-search scaling sweep): the c2 map layout at other pose counts / beam counts.
# usage (under gpurun, from the repo root): bash profiles/run_c5_sweep.sh <tag> "<poses>x<beams> ..." [gpus]
TAG=${1:-r2}; POINTS=${2:-"1000x360 3000x720 10000x360 5000x2160"}; GPUS=${3:-1}
mkdir -p gpurun_out
OUT=gpurun_out/c5_sweep_${TAG}.jsonl
: > $OUT
for pb in $POINTS; do
  P=${pb%x*}; B=${pb#*x}
  if [ "$GPUS" = "1" ]; then
    python bench.py --workload c2 --poses $P --beams $B --steps 3 --warmup 4 --no-correction --cpu-seconds 6 >> $OUT 2>> gpurun_out/c5_sweep_${TAG}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $GPUS --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $GPUS --workload c2 --poses $P --beams $B \
      --steps 3 --warmup 4 --no-correction --no-largest-map >> $OUT 2>> gpurun_out/c5_sweep_${TAG}.err
  fi
done
python - <<PY
import json
for l in open("$OUT"):
    d = json.loads(l)
    c = d.get("cpu_baseline") or {}
    print(d["config"]["n_poses"], "poses,", d["config"]["n_points"], "points, N=%d:" % d["n_gpus"], round(d["ms_per_step"], 3), "ms/step,", round(d["value"]), "M evals/s, e2e", round(d["e2e"]["ms_per_step"], 2), "ms,",
          d["detail"]["queries_per_step"], "queries,", d["detail"]["jacobian_evals_per_step"], "matches,", d["detail"]["tree_walks_per_step"], "walks; cpu", c.get("kind"), round(c.get("value", 0), 1),
          "port", round((c.get("port") or {}).get("value", 0), 1), "parity", (d.get("parity_checked") or {}).get("all_ok"))
PY
