#!/bin/bash
# Kernel-iteration loop (one B200): parity of the search on the small maps, c1 and the c2 full-size chunks, then the c2 bench line
# (device-resident value, search-kernel time, counters, parity_checked).  usage: gpurun --timeout 900 -- 'bash profiles/kiter.sh <tag>'
TAG=${1:-k}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference.py -m gpu -q -x -k "find_stf or stf or search" --deselect "tests/test_gpu_parity.py::test_find_stf_full_size_matches_oracle_on_chunks[c3-5]" 2>&1 | tail -4 | tee gpurun_out/pytest_${TAG}.log
python bench.py --steps 10 --warmup 4 --no-cpu --no-e2e --no-correction > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
tail -c 400 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}.json"))
r = d["roofline"]
print("ms_per_step %.3f | search %.3f ms | find_stf %.3f | normal_eq %.3f | walks %d | tile_pairs %d | parity %s" % (
    d["ms_per_step"], r["ms_kernel"], d["detail"]["ms_find_stf"], d["detail"]["ms_normal_eq"], d["detail"]["tree_walks_per_step"], d["detail"]["tile_pairs_per_step"], d["parity_checked"]))
PY
