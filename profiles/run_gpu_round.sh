#!/bin/bash
# One GPU-box session (one B200): parity tests, ncu launch list, ncu --set full of the kernels the bench line's `roofline` cites.
# usage (from the repo root, under gpurun): bash profiles/run_gpu_round.sh <tag>
# Afterwards, here: python profiles/ncu_counters.py stf_search_kernel=gpurun_out/prof_search_<tag>.ncu-rep eval_stf_kernel=... em_inliers_kernel=... em_fit_kernel=...
TAG=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_${TAG}.log
B="python bench.py --steps 2 --warmup 4 --no-cpu --no-e2e --no-parity --replay 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/launch_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stf_search_kernel --launch-skip 5 -c 1 -f -o gpurun_out/prof_search_${TAG} $B > gpurun_out/prof_search_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_stf_kernel --launch-skip 5 -c 1 -f -o gpurun_out/prof_eval_${TAG} $B > gpurun_out/prof_eval_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:em_inliers_kernel --launch-skip 2 -c 1 -f -o gpurun_out/prof_em_${TAG} $B > gpurun_out/prof_em_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:em_fit_kernel --launch-skip 2 -c 1 -f -o gpurun_out/prof_emfit_${TAG} $B > gpurun_out/prof_emfit_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stf_order_kernel --launch-skip 10 -c 2 -f -o gpurun_out/prof_order_${TAG} $B > gpurun_out/prof_order_${TAG}.log 2>&1
ls -la gpurun_out/*_${TAG}*
