#!/bin/bash
# ncu --set full of one warmed stf_search_kernel launch on c2 (profiles/diag_cull.py runs the search 3x per setting).
TAG=${1:-x}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:stf_search_kernel --launch-skip 2 -c 1 -f -o gpurun_out/prof_search_${TAG} \
    python profiles/diag_cull.py > gpurun_out/prof_search_${TAG}.log 2>&1
tail -3 gpurun_out/prof_search_${TAG}.log
