#!/bin/bash
# usage: scratch/run_bench.sh [extra bench args]; prints a digest of the JSON line
python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/b.json 2> gpurun_out/b.err; tail -c 600 gpurun_out/b.err
python - <<PY
import json
d=json.load(open("gpurun_out/b.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"], "loss", d["config"]["final_loss"], d["clocks"])
r=d["roofline"]; print(r["kernel"], r["achieved"], r["frac"], "own share", r["own_kernels_share_of_step"])
for k in r["kernels"]: print(k["call"], k["key"], round(k["launches_per_step"],1), round(k["mean_us"],1), round(k["share_of_step"],3), round(k.get("GBps",0)), round(k.get("TFLOPs",0),1))
print(d.get("cpu_baseline"))
PY
