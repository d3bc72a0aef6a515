timeout -s KILL 300 python -m pytest tests/test_fused_gpu.py -q -x 2>&1 | tail -2
timeout -s KILL 150 python profiles/tools/sa_b3_ab.py "" 0 all 2>&1 | grep "all kernels" | cut -c1-230
