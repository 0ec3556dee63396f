#!/bin/bash
# One GPU validation round (run under gpurun): parity tests, smoke, bench (both arms), then -- only after
# the plain runs exited 0 -- the ncu launch list of the same bench command and one --set full capture of
# the dominant kernel.  Everything lands in gpurun_out/.   usage: tools/gpu_round.sh [tag] [kernel-regex]
TAG=${1:-r01}
KREGEX=${2:-k_path}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; rc=$?; echo "bench rc=$rc"
tail -c 600 $OUT/bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
if [ $rc -eq 0 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 4 -c 1 -f -o $OUT/${TAG}_${KREGEX} \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
