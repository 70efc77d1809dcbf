#!/bin/bash
# N-GPU bench lines (launched as the driver does).  usage: bash profiles/run_gpu_multi.sh <tag> <N...>
TAG=${1:-r1}; shift
mkdir -p gpurun_out
for N in "$@"; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N \
      > gpurun_out/bench_${TAG}_n${N}.json 2> gpurun_out/bench_${TAG}_n${N}.err
  fi
  tail -c 2500 gpurun_out/bench_${TAG}_n${N}.json; tail -5 gpurun_out/bench_${TAG}_n${N}.err
done
