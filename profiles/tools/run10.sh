timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "density" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_fullsize_gpu.py -q -x -k "pointconv or PointConv" 2>&1 | tail -3
timeout 400 python bench.py --workload pointconv --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/pc.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pointconv', d['ms_per_step'], d['value'], d['config']['cuda_graph'], d['config']['cuda_graph_error'])"; tail -2 gpurun_out/pc.err
