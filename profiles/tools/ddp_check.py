"""torchrun --nproc-per-node N profiles/tools/ddp_check.py — the bucketed, event-gated all-reduce (Trainer
overlap_allreduce=True, CUDA graph + external event-record node) against the plain blocking all-reduce:
same seeds, same per-rank batches, K steps each; the parameter buckets must agree to atomics noise and every rank
must hold identical parameters.  Prints per-step times of both."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, '.')
rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dev = torch.device("cuda", lr_)
dist.init_process_group("nccl", device_id=dev)
from pointcloudlib_b200.networks.cls.pointnet2 import PointNetMSG
from pointcloudlib_b200.synthetic import modelnet_batch
from pointcloudlib_b200.train import Trainer
B, N, K = 16, 2048, 8
batches = [tuple(t.to(dev) for t in modelnet_batch(B, N, seed=100 * rank + s)) for s in range(4)]
res = {}
for overlap in (False, True):
    torch.manual_seed(0)
    model = PointNetMSG(n_classes=40).to(dev).train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    tr = Trainer(model, lr=0.01, graph=True, overlap_allreduce=overlap)
    for i in range(6):
        x, n, l = batches[i % 4]
        tr.step(x, n, labels=l)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        x, n, l = batches[i % 4]
        loss = tr.step(x, n, labels=l)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    p = tr.opt.params.clone()
    # identical replicas?
    ref = p.clone(); dist.broadcast(ref, 0)
    same = float((p - ref).abs().max())
    res[overlap] = (p, ms, float(loss), same, tr.allreduce_mode, tr._graph is not None, tr.graph_error)
    if rank == 0:
        print(f"overlap={overlap}: {ms:.3f} ms/step, loss {float(loss):.5f}, max |param - rank0 param| {same:.2e}, graph={tr._graph is not None} err={tr.graph_error}\n   mode: {tr.allreduce_mode}", flush=True)
d = ((res[True][0] - res[False][0]).norm() / res[False][0].norm()).item()
if rank == 0:
    print(f"relative parameter difference overlapped vs blocking after {6 + K} steps: {d:.3e}")
    assert d < 5e-3, d
    assert res[True][3] == 0.0 and res[False][3] == 0.0
    print("ddp_check ok")
dist.destroy_process_group()
