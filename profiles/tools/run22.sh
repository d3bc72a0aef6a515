timeout -s KILL 300 python -m pytest tests/test_fused_gpu.py -q -x 2>&1 | tail -2
echo "--- default"; timeout -s KILL 200 python profiles/tools/sa_b3_ab.py 1,2,5 0 all 2>&1 | grep "all kernels"
echo "--- knob 1024 (LAG=1)"; timeout -s KILL 200 python profiles/tools/sa_b3_ab.py 1,2,5 1024 all 2>&1 | grep "all kernels\|diff"
