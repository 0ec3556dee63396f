timeout 300 python tools/tune.py c2 kernel=5 tiny_mode=0,1 2>&1 | grep -v scene
timeout 300 python tools/tune.py c1 kernel=5 tiny_mode=0,1 2>&1 | grep -v scene
