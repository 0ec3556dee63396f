M="l1tex__t_bytes.sum,lts__t_bytes.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"
for L in 0 6; do
SPP=4 DEPTH=4 SWEEP="treelet=$L" timeout 600 ncu --set full --metrics $M --clock-control none -f -k regex:'k_wf_trace|k_wf_packet0' -s 4 -c 2 -o gpurun_out/r02h_treelet$L python tools/exp_bounce.py > gpurun_out/r02h_ncu$L.log 2>&1; echo rc=$?
done
