timeout -s KILL 200 python -m pytest tests/test_fused_gpu.py -q -x -k "matches_reference_sequence and 300-25" 2>&1 | tail -4
