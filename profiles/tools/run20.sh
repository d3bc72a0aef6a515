run() { w=$1; n=$2
  timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --workload $w --steps 10 --warmup 3 > gpurun_out/bench_r02c_${n}gpu_$w.json 2>gpurun_out/bench_r02c_${n}gpu_$w.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_r02c_${n}gpu_$w.json").read().strip().splitlines()[-1])
    print("$w", d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"])
except Exception as e:
    print("$w $n failed", e)
P
}
run dgcnn 8; run pointconv 8; run pointconv 4; run dgcnn 4
