#!/bin/bash
# bench.py at N = 1, 2, 4, 8 the way the driver launches it (run under gpurun --gpus 8).  $1 = exchange mode override.
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
[ -n "$1" ] && export BENCH_EXCHANGE=$1
python bench.py --no-cpu-baseline --steps 30 > $OUT/scale_n1.json 2> $OUT/scale_n1.err; echo "n=1 rc=$?"
for n in 2 4 8; do
  timeout 600 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 30 --warmup 5 > $OUT/scale_n$n.json 2> $OUT/scale_n$n.err; echo "n=$n rc=$?"
done
