timeout 600 python -m pytest tests/test_gpu_reference_host.py tests/test_gpu_edits_and_edge_cases.py -x -q -k "reference_host or banded" 2>&1 | tail -25
