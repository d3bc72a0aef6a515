timeout -s KILL 300 python -m pytest tests/test_fused_gpu.py -q -x 2>&1 | tail -3
timeout -s KILL 200 python profiles/tools/sa_b3_ab.py "" 16384 2>&1 | grep -v "^Trace" | grep "preload=1\|diff\|cycles" | tail -20
