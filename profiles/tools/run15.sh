timeout -s KILL 120 python profiles/tools/routed_bisect.py 2>&1 | tail -4
timeout -s KILL 400 python -m pytest tests/test_fused_gpu.py -q -x 2>&1 | tail -4
timeout -s KILL 200 python profiles/tools/sa_b3_ab.py 2>&1 | grep -v "^Trace" | grep "preload=1\|diff" | tail -20
