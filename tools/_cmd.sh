timeout 600 python -m pytest tests/test_gpu_tiny.py -x -q 2>&1 | tail -3
for lib in libb200rt.so libb200rt_mb5.so; do
 echo "== $lib"
 B200RT_LIB=$PWD/pgr_raytracing_project_b200/$lib timeout 300 python tools/tune.py c2 kernel=5 tiny_threads=256,128 2>&1 | grep -v scene
 B200RT_LIB=$PWD/pgr_raytracing_project_b200/$lib timeout 300 python tools/tune.py c1 kernel=5 tiny_threads=256,128 2>&1 | grep -v scene
done
timeout 300 python tools/tune.py c2 kernel=-1,-1,-1  2>&1 | grep -v scene
timeout 300 python tools/tune.py c1 kernel=-1,-1,-1  2>&1 | grep -v scene
