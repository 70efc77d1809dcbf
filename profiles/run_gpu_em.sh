#!/bin/bash
# One short GPU-box session for the correction path: host-mirror tests, the default bench line, the ncu launch list of the short
# bench command and ONE ncu --set full pass over the EM kernels of the last correction repetition + the E-step roofline launches
# (3 full passes with the chunk cull off, one pass that records the chunk boxes, one culled pass).
# usage (repo root, under gpurun): bash profiles/run_gpu_em.sh <tag>
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests/test_host_mirror_gpu.py -q --tb=short > gpurun_out/pytest_host_${TAG}.log 2>&1; tail -3 gpurun_out/pytest_host_${TAG}.log
python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}_n1.json"))
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"].get("steps_ms"), d["correction_latency"]["ms"], d["correction_latency"]["parts_ms"], d["parity_checked"]["all_ok"])
PY
B="python bench.py --steps 2 --warmup 4 --no-cpu --no-e2e --no-parity --replay 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/launch_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"em_inliers_kernel|em_fit_kernel" --launch-skip 40 -c 13 -f -o gpurun_out/prof_em_${TAG} $B > gpurun_out/prof_em_${TAG}.log 2>&1
ls -la gpurun_out/*_${TAG}*
