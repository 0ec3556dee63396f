#!/bin/bash
# ncu captures of the kernels that have no committed evidence yet (run under gpurun, one GPU): C2 (tiny scene),
# C4 (10M triangles: the one working set beyond L2) and the C3 incoherent-bounce / camera-ray kernels, each with the
# L1 / L2 byte and throughput counters next to --set full.   usage: tools/gpu_prof.sh <tag> [c2] [c4] [c3d4] [c3]
TAG=${1:-r02a}; shift
WHAT=${@:-c2 c4 c3d4 c3}
OUT=gpurun_out
mkdir -p $OUT
M="l1tex__t_bytes.sum,lts__t_bytes.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sectors.sum,lts__t_sectors.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum"
NCU="ncu --set full --metrics $M --clock-control none --import-source on -f"
for w in $WHAT; do
  case $w in
    c2)   timeout 300 python tools/tune.py c2 > $OUT/${TAG}_tune_c2.log 2>&1; echo "tune c2 rc=$?"
          timeout 900 $NCU -k regex:'k_wf_trace|k_render|k_tiny' -c 8 -o $OUT/${TAG}_c2 python tools/tune.py c2 > $OUT/${TAG}_ncu_c2.log 2>&1; echo "ncu c2 rc=$?";;
    c4)   timeout 600 python tools/tune.py c4 > $OUT/${TAG}_tune_c4.log 2>&1; echo "tune c4 rc=$?"
          timeout 1200 $NCU -k regex:'k_wf_trace|k_wf_packet0' -s 4 -c 4 -o $OUT/${TAG}_c4 python tools/tune.py c4 > $OUT/${TAG}_ncu_c4.log 2>&1; echo "ncu c4 rc=$?";;
    c3d4) SPP=4 DEPTH=4 timeout 300 python tools/exp_bounce.py > $OUT/${TAG}_bounce_c3d4.log 2>&1; echo "bounce rc=$?"
          SPP=4 DEPTH=4 timeout 900 $NCU -k regex:'k_wf_trace' -s 3 -c 3 -o $OUT/${TAG}_c3d4 python tools/exp_bounce.py > $OUT/${TAG}_ncu_c3d4.log 2>&1; echo "ncu c3d4 rc=$?";;
    c3)   timeout 300 python tools/tune.py c3 > $OUT/${TAG}_tune_c3.log 2>&1; echo "tune c3 rc=$?"
          timeout 900 $NCU -k regex:'k_packet' -s 4 -c 1 -o $OUT/${TAG}_c3 python tools/tune.py c3 > $OUT/${TAG}_ncu_c3.log 2>&1; echo "ncu c3 rc=$?";;
  esac
done
ls -la $OUT | grep $TAG
