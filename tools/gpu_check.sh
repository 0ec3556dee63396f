#!/bin/bash
# Validation round under gpurun (one GPU): parity tests, smoke, bench (both arms).   usage: tools/gpu_check.sh [tag] [pytest -k expr]
TAG=${1:-chk}
OUT=gpurun_out
mkdir -p $OUT
nproc > $OUT/nproc.txt
if [ -n "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -k "$2" > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
else
  timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
fi
tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?"
tail -c 1500 $OUT/${TAG}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref rc=$?"
