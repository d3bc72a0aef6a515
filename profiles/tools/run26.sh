timeout -s KILL 200 python -m pytest tests/test_fused_gpu.py -q -x -k "left_operand or matches_reference" 2>&1 | tail -2
timeout -s KILL 150 python profiles/tools/sa_b3_ab.py 1,2 16384 all 2>&1 | grep "all kernels\|cycles {copy" | cut -c1-200
