timeout 900 python -m pytest tests/test_fused_gpu.py tests/test_models_gpu.py tests/test_fullsize_gpu.py -q -x 2>&1 | tail -4
timeout 300 python profiles/tools/sa3_one.py 2>&1 | grep "dense=" | cut -c1-600 | head -1
for w in pointnet2_msg dgcnn partseg pointconv; do timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02m_$w.json 2> gpurun_out/bench_r02m_$w.err; python - <<P
import json
d=json.loads(open("gpurun_out/bench_r02m_$w.json").read().strip().splitlines()[-1])
print("$w", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["own_kernels_share_of_step"])
P
done
