set -x
free -g | head -2; nproc
python -c "import torch;print(torch.cuda.get_device_name(0))"
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_fullsize_gpu.py 2>&1 | tail -40 > gpurun_out/t1.log
tail -5 gpurun_out/t1.log
timeout 900 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -s 2>&1 | tail -40 > gpurun_out/t1_full.log
tail -12 gpurun_out/t1_full.log
timeout 300 python profiles/tools/bq_sweep.py > gpurun_out/bq_sweep.txt 2>&1; tail -15 gpurun_out/bq_sweep.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; tail -c 600 gpurun_out/bench_r02_a.err
for w in dgcnn partseg pointconv; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_$w.json 2> gpurun_out/bench_r02_$w.err; tail -c 300 gpurun_out/bench_r02_$w.err; done
python - <<'P'
import json
for n in ("a","dgcnn","partseg","pointconv"):
    try:
        d=json.loads(open(f"gpurun_out/bench_r02_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["cuda_graph"], d["config"]["cuda_graph_error"])
    except Exception as e: print(n, "ERR", e)
P
