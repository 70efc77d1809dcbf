#!/bin/bash
# One GPU-box session: parity tests, bench line, ncu launch list, ncu --set full of the two top kernels.
# usage (from the repo root, under gpurun): bash profiles/run_gpu_round.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_${TAG}.log
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 3000 gpurun_out/bench_${TAG}.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/launch_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stf_search_kernel -c 1 -f -o gpurun_out/prof_search_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/prof_search_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_stf_kernel -c 1 -f -o gpurun_out/prof_eval_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/prof_eval_${TAG}.log 2>&1
ls -la gpurun_out
