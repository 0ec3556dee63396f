#!/bin/bash
# 8-GPU round (run under gpurun --gpus 8): bench.py at N = 2, 4, 8 the way the driver launches it, the reference arm at N = 8
# (rank 0 alone works, all host cores), and the C4 / C3-multibounce configs at 8 GPUs.   usage: tools/gpu_scale8.sh [tag] [configs]
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 4 8; do
  timeout 400 $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n --steps 30 --warmup 5 > $OUT/${TAG}_scale_n$n.json 2> $OUT/${TAG}_scale_n$n.err; echo "n=$n rc=$?"
done
timeout 400 $TR --nproc-per-node 8 --master-port 29549 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $OUT/${TAG}_ref_n8.json 2> $OUT/${TAG}_ref_n8.err; echo "ref n=8 rc=$?"
if [ -n "$2" ]; then
  timeout 900 $TR --nproc-per-node 8 --master-port 29558 tools/bench_configs.py --configs $2 --reps 3 > $OUT/${TAG}_configs_n8.jsonl 2> $OUT/${TAG}_configs_n8.err; echo "configs n=8 rc=$?"
fi
for n in 2 4 8; do python -c "
import json; d=json.load(open('$OUT/${TAG}_scale_n$n.json')); print('n=$n', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['frame_matches_1gpu'], d['host_frame_matches_1gpu'], 'mb', round(d['multibounce']['ms_per_frame'],3))"; done
python -c "
import json; d=json.load(open('$OUT/${TAG}_ref_n8.json')); print('ref n=8', d['value'], d['cpu_baseline']['cores'], d['config']==json.load(open('$OUT/${TAG}_scale_n8.json'))['config'])"
