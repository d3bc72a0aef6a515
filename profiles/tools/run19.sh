for w in pointnet2_msg dgcnn pointconv; do
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload $w --steps 10 --warmup 3 > gpurun_out/bench_r02c_2gpu_$w.json 2>gpurun_out/bench_r02c_2gpu_$w.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_r02c_2gpu_$w.json").read().strip().splitlines()[-1])
    print("$w", d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["allreduce"], d["config"]["cuda_graph"])
except Exception as e:
    print("$w failed", e)
P
done
