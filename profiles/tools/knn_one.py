"""kNN at the DGCNN shapes (B=32, N=1024, k=20, C = 3 / 64 / 128): CUDA-event time per call; PCL_KNN_LEGACY=1
selects the round-1 staging.  Run under ncu for the stall breakdown."""
import sys, torch
sys.path.insert(0, '.')
from pointcloudlib_b200 import functional as F
torch.manual_seed(0)
for C in (3, 64, 128):
    x = torch.randn(32, C, 1024, device='cuda')
    F.knn(x, x, 20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): F.knn(x, x, 20)
    e1.record(); torch.cuda.synchronize()
    print(f"C={C}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us", flush=True)
