#!/bin/bash
# 8-GPU round (run under gpurun --gpus 8): bench.py at N = 2, 4, 8 the way the driver launches it (flag barrier), N = 8 with
# the NCCL barrier for comparison, and the C4 / C3-multibounce configs at 8 GPUs.
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 30 --warmup 5 > $OUT/scale_n$n.json 2> $OUT/scale_n$n.err; echo "n=$n rc=$?"
done
BENCH_BARRIER=nccl timeout 300 $TR --nproc-per-node 8 --master-port 29548 bench.py --gpus 8 --steps 30 --warmup 5 > $OUT/scale_n8_nccl.json 2> $OUT/scale_n8_nccl.err; echo "n=8 nccl rc=$?"
timeout 600 $TR --nproc-per-node 8 --master-port 29558 tools/bench_configs.py --configs c3d4,c4 --reps 3 > $OUT/configs_n8.jsonl 2> $OUT/configs_n8.err; echo "configs n=8 rc=$?"
for f in scale_n2 scale_n4 scale_n8 scale_n8_nccl; do python -c "import json; d=json.load(open('$OUT/$f.json')); print('$f', d['value'], d['ms_per_step'])"; done
