"""ball-query+group A/B on config 2's six shapes: round-1 warp-per-centroid writer (PCL_BQ_LEGACY=1) vs the
one-scan kernel with the CTA-cooperative float4 writer, centroids-per-CTA and unroll sweeps, and the three
radii of a level in ONE launch.  Prints GB/s of ALGORITHMIC bytes and the fraction of the HBM peak."""
import json, os, sys
import torch
sys.path.insert(0, '.')
from pointcloudlib_b200 import functional as PF
from pointcloudlib_b200.synthetic import modelnet_batch

peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
dev = "cuda"
xyz0, nrm0, _ = (t.to(dev) for t in modelnet_batch(32, 4096, seed=1000))
cen1 = PF.gather_xyz(xyz0, PF.furthest_point_sample(xyz0, 512))
cen2 = PF.gather_xyz(cen1, PF.furthest_point_sample(cen1, 128))
feat2 = torch.randn(32, 512, 320, device=dev)


def nbytes(B, N, S, ns, C):
    return 4 * (B * N * (3 + C) + 3 * B * S + B * S * ns + B * S * ns * (3 + C))


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps


levels = [("SA1", cen1, xyz0, nrm0, (0.1, 0.2, 0.4), (16, 32, 128)), ("SA2", cen2, cen1, feat2, (0.2, 0.4, 0.8), (32, 64, 128))]
configs = [("legacy", dict(PCL_BQ_LEGACY="1")), ("flat writer, 256 thr", dict(PCL_BQ_WIDE="0", PCL_BQ_THREADS="256", PCL_BQ_CPB="8"))]
for cpb in (2, 4, 8, 16):
    configs.append((f"new cpb={cpb} (wide only)", dict(PCL_BQ_CPB=str(cpb))))
configs.append(("new default", dict()))
for name, env in configs:
    for k in ("PCL_BQ_LEGACY", "PCL_BQ_CPB", "PCL_BQ_UNROLL", "PCL_BQ_WIDE", "PCL_BQ_THREADS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    row = []
    for tag, cen, pts, feat, radii, nss in levels:
        tot = 0
        for r, ns in zip(radii, nss):
            us = timeit(lambda: PF.ball_query_group(cen, pts, feat, r, ns))
            by = nbytes(32, pts.shape[1], cen.shape[1], ns, feat.shape[2])
            tot += by
            row.append(f"{tag} ns={ns}: {us:6.1f}us {by / us / 1e3 / peak:.2f}")
        if name != "legacy":
            us = timeit(lambda: PF.ball_query_group_msg(cen, pts, feat, radii, nss))
            row.append(f"{tag} x3: {us:6.1f}us {tot / us / 1e3 / peak:.2f}")
            usq = timeit(lambda: PF.ball_query_msg(cen, pts, radii, nss))
            row.append(f"{tag} query x3: {usq:6.1f}us")
        else:
            usq = sum(timeit(lambda: PF.ball_query(cen, pts, r, ns)) for r, ns in zip(radii, nss))
            row.append(f"{tag} query 3 launches: {usq:6.1f}us")
    print(f"{name:22s} | " + " | ".join(row), flush=True)
