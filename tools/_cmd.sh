for i in 1 2 3; do timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -2; done
