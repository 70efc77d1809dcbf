#!/bin/bash
# BASELINE config 5 (correspondence-search scaling sweep): the c2 map layout at other pose counts / beam counts, one B200.
# usage (under gpurun, from the repo root): bash profiles/run_c5_sweep.sh <tag> "<poses>x<beams> ..."
TAG=${1:-r1}; shift
POINTS=${1:-"1000x360 3000x720 10000x360 5000x2160"}
mkdir -p gpurun_out
: > gpurun_out/c5_sweep_${TAG}.jsonl
for pb in $POINTS; do
  P=${pb%x*}; B=${pb#*x}
  python bench.py --workload c2 --poses $P --beams $B --steps 3 --warmup 4 --no-e2e --no-correction --cpu-seconds 6 \
    >> gpurun_out/c5_sweep_${TAG}.jsonl 2>> gpurun_out/c5_sweep_${TAG}.err
done
python - <<PY
import json
for l in open("gpurun_out/c5_sweep_${TAG}.jsonl"):
    d = json.loads(l)
    c = d.get("cpu_baseline") or {}
    print(d["config"]["n_poses"], "poses,", d["config"]["n_points"], "points:", round(d["ms_per_step"], 3), "ms/step,", round(d["value"]), "M evals/s,",
          d["detail"]["queries_per_step"], "queries,", d["detail"]["jacobian_evals_per_step"], "matches; cpu", c.get("kind"), round(c.get("value", 0), 1),
          "port", round((c.get("port") or {}).get("value", 0), 1))
PY
