set -x
timeout -s KILL 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout -s KILL 700 ncu --set full --clock-control none --import-source on -k regex:"rowgemm_ws2_kernel|rowgemm_ws_kernel|wgrad_ws_kernel|sel_outer|gather_bn_backward|gather_stats|routed_sort" -c 14 -o gpurun_out/ws_r02c python profiles/tools/sa_branch.py 3 1 > gpurun_out/ncu_ws_c.log 2>&1; tail -2 gpurun_out/ncu_ws_c.log
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r02c.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu_launches_c.log 2>&1; tail -2 gpurun_out/ncu_launches_c.log
for w in pointnet2_msg dgcnn partseg pointconv; do
  timeout -s KILL 500 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_r02c_$w.json 2>gpurun_out/bench_r02c_$w.err
  python - <<P
import json
d=json.loads(open("gpurun_out/bench_r02c_$w.json").read().strip().splitlines()[-1])
print("$w", d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["config"]["cuda_graph"], d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
P
done
