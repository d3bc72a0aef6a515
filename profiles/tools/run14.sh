timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python profiles/tools/sa_b3_ab.py 2>&1 | grep -v "^Trace" | tail -20
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pointnet2_msg.json 2>gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_pointnet2_msg.json').read().strip().splitlines()[-1])
r=d['roofline']
print(d['ms_per_step'], d['value'], d['e2e']['value'], r['kernel'], r['frac'], r['own_kernels_share_of_step'], r['instrumented_step_us'])
for k in r['kernels'][:24]: print(k['call'], k['key'], round(k['mean_us']), round(k.get('hbm_frac',0),3), round(k['share_of_step'],3))
PY
