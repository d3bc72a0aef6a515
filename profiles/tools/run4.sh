set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 profiles/tools/ddp_check.py > gpurun_out/ddp_check.txt 2>&1; tail -12 gpurun_out/ddp_check.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err; tail -c 300 gpurun_out/bench_r02_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 20 --warmup 5 --no-overlap > gpurun_out/bench_r02_2gpu_blocking.json 2> gpurun_out/bench_r02_2gpu_blocking.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r02_1gpu_ref.json 2>/dev/null
python - <<'P'
import json
for n in ("2gpu","2gpu_blocking","1gpu_ref"):
    try:
        d=json.loads(open(f"gpurun_out/bench_r02_{n}.json").read().strip().splitlines()[-1])
        print(n, d["n_gpus"], round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["config"]["allreduce"][:60])
    except Exception as e: print(n,"ERR",e)
P
