timeout -s KILL 300 python -m pytest tests/test_fused_gpu.py -q -x 2>&1 | tail -3
timeout -s KILL 200 python profiles/tools/sa_b3_ab.py 2 0 all 2>&1 | grep "all kernels" | tail -1
