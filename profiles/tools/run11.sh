timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in pointnet2_msg dgcnn partseg pointconv; do
  timeout 500 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_$w.json 2>gpurun_out/bench_$w.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1])
print('$w', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline'].get('kernel'), d['roofline'].get('frac'), d['config'].get('cuda_graph'))
"
done
