set -x
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; free -g | head -2; nproc
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/t2.log; tail -4 gpurun_out/t2.log
timeout 200 python profiles/tools/bq_sweep.py > gpurun_out/bq_sweep2.txt 2>&1; tail -14 gpurun_out/bq_sweep2.txt
python - <<'P'
import torch, time
# pure write / copy ceilings on this box (the denominators the ball-query+group writer is quoted against)
for n_mb in (700,):
    x = torch.empty(n_mb * 250000, device="cuda"); y = torch.empty_like(x)
    def t(fn, reps=20):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3
    print("fill_ GB/s", x.numel() * 4 / t(lambda: x.fill_(1.0)) / 1e9, " copy_ GB/s (r+w)", 2 * x.numel() * 4 / t(lambda: y.copy_(x)) / 1e9)
P
timeout 300 python profiles/tools/sa_branch_ab.py > gpurun_out/sa_branch_ab.txt 2>&1; cat gpurun_out/sa_branch_ab.txt | tail -20
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err; tail -c 400 gpurun_out/bench_r02_b.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_r02_b.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k in d["roofline"]["kernels"][:14]: print(k["call"], k["key"], round(k["mean_us"],1), round(k.get("hbm_frac",0),2))
P
# ncu: the two ball-query+group writers at the SA2 ns=128 shape, and the new last-layer backward kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ball_query" -c 6 -o gpurun_out/bq_r02 python profiles/tools/bq_one.py > gpurun_out/ncu_bq.log 2>&1; tail -3 gpurun_out/ncu_bq.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rowgemm_ws_kernel|wgrad_ws_kernel|sel_outer" -c 12 -o gpurun_out/ws_r02 python profiles/tools/sa_branch.py 3 1 > gpurun_out/ncu_ws.log 2>&1; tail -3 gpurun_out/ncu_ws.log
