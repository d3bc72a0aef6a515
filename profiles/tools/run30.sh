timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r02e_2gpu_pointnet2_msg.json 2>gpurun_out/bench_r02e_2gpu.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_r02e_2gpu_pointnet2_msg.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["allreduce"])
P
