#!/bin/bash
# First GPU call of the next round (one B200): what this round's GPU budget did not cover.
#   1. the drop-in comparison (reference JointOpt with the hot path bound to the C ABI) that is gated behind HITL_DROPIN_TEST
#   2. a full default bench line with the final code, and the e2e leg with the compact boundary formats (HITL_E2E_COMPACT=1:
#      u32 tree nodes, u16 point indices: 72 MB less H2D and 80 MB less D2H per step at c2)
#   3. the residency-variant view of a shard (is a shard's critical path shorter with fewer resident warps?)
#   4. BASELINE config 5 sweep points
# usage (from the repo root): gpurun --timeout 900 -- 'bash profiles/run_next_round_first.sh r2a'
# then, separately (8x the box time):  gpurun --gpus 8 --timeout 240 -- 'bash profiles/run_gpu_multi.sh r2a 8'
TAG=${1:-r2a}
mkdir -p gpurun_out
HITL_DROPIN_TEST=1 HITL_COMPACT_TEST=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_${TAG}.log
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 2500 gpurun_out/bench_${TAG}.json
HITL_E2E_COMPACT=1 python bench.py --no-cpu --no-correction > gpurun_out/bench_${TAG}_compact.json 2> gpurun_out/bench_${TAG}_compact.err; tail -c 1200 gpurun_out/bench_${TAG}_compact.json
python profiles/diag_variants.py c2 > gpurun_out/variants_${TAG}.txt 2>&1; tail -20 gpurun_out/variants_${TAG}.txt
bash profiles/run_c5_sweep.sh ${TAG} "1000x360 3000x720 10000x360 5000x2160"
