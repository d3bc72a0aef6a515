timeout 900 python -m pytest tests/test_fused_gpu.py -x -q 2>&1 | tail -4
timeout 300 python profiles/tools/sa_b3_ab.py "" 16384 2>&1 | grep -v "^Trace" | tail -32
