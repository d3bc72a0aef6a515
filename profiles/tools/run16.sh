timeout -s KILL 400 python -m pytest tests/test_fused_gpu.py -q -x 2>&1 | tail -2
timeout -s KILL 200 python profiles/tools/sa_b3_ab.py 1,2,5 0 all 2>&1 | grep -v "^Trace" | grep "all kernels" | tail -20
