timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_ref_kernels_gpu.py -q -x -k "knn or nn or interp" 2>&1 | tail -4
timeout 200 python profiles/tools/knn_one.py
