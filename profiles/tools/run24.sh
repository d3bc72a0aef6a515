timeout -s KILL 300 python -m pytest tests/test_fused_gpu.py -q -x 2>&1 | tail -4
echo "--- tl kernel"; timeout -s KILL 200 python profiles/tools/sa_b3_ab.py 1,2,5 0 all 2>&1 | grep "all kernels"
echo "--- knob 2048 (shared-memory kernel)"; timeout -s KILL 200 python profiles/tools/sa_b3_ab.py 1,2,5 2048 all 2>&1 | grep "all kernels"
