timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edits_and_edge_cases.py -x -q -k "packet or auto or shared_host or two_streams" 2>&1 | tail -3
timeout 300 python tools/tune.py c3s8 fold=1,0,1 2>&1 | grep -v scene
timeout 300 python tools/tune.py c3s2 fold=1,0,1 2>&1 | grep -v scene
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02i_bench_n2.json 2> gpurun_out/r02i_bench_n2.err; echo "bench n2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02i_bench_n2.json'))
for k in ['value','ms_per_step','ms_per_step_stats','frame_matches_1gpu','host_frame_matches_1gpu','e2e','gpu_launches']: print(k, d[k])
"
