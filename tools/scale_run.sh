#!/bin/bash
# Multi-GPU evidence on ONE box with N GPUs visible (gpurun --gpus N): bare H2D ceiling and bench.py at N = 8 / 4 / 2 (whatever fits).
# usage: bash tools/scale_run.sh "8 4 2" [extra bench flags]
NS=${1:-"8 4 2"}; shift
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for N in $NS; do
  if [ "$N" -gt "$NG" ]; then continue; fi
  P=$((29500 + N))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P tools/measure_h2d.py > gpurun_out/r02_h2d_n$N.json 2> gpurun_out/r02_h2d_n$N.err
  tail -1 gpurun_out/r02_h2d_n$N.json
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P + 20)) bench.py --gpus $N "$@" > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
  python tools/show_bench.py gpurun_out/r02_bench_n$N.json 2>&1 | head -1
done
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1
