#!/bin/bash
# Multi-GPU round (run under gpurun --gpus 8): bench.py at N = 2, 4, 8 the way the driver launches it, the
# multi-process NCCL/IPC test, and the C4 / C5 / C3-multibounce configs at 8 GPUs.
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > $OUT/pytest_multigpu.log 2>&1; echo "pytest multigpu rc=$?"
for n in 2 4 8; do
  timeout 600 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --steps 30 --warmup 5 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err; echo "bench n=$n rc=$?"
done
timeout 900 $TR --nproc-per-node 8 --master-port 29538 tools/bench_configs.py --configs c3d4,c4,c5 > $OUT/configs_n8.jsonl 2> $OUT/configs_n8.err; echo "configs n=8 rc=$?"
