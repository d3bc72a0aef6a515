import copy, sys, torch
sys.path.insert(0, '.')
from oracle import model_oracle
from pointcloudlib_b200.synthetic import modelnet_batch
from pointcloudlib_b200.train import soft_cross_entropy_loss
from pointcloudlib_b200.networks.cls import pointnet2
for B in (4, 16):
    torch.manual_seed(0)
    model = pointnet2.PointNet2_cls(n_classes=40); model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout): m.p = 0.0
    ref = copy.deepcopy(model); ref64 = copy.deepcopy(model).double()
    model = model.cuda()
    xyz, nrm, lab = modelnet_batch(B, 1024, seed=3)
    l = soft_cross_entropy_loss(model(xyz.cuda(), nrm.cuda()), lab.cuda()); l.backward()
    lr = model_oracle.soft_cross_entropy_loss(model_oracle.pointnet2_cls(ref, xyz, nrm), lab); lr.backward()
    l64 = model_oracle.soft_cross_entropy_loss(model_oracle.pointnet2_cls(ref64, xyz.double(), nrm.double()), lab); l64.backward()
    print("B", B, float(l), float(lr), float(l64))
    for (n, p), (_, q), (_, r) in zip(model.named_parameters(), ref.named_parameters(), ref64.named_parameters()):
        s = r.grad.abs().max().item()
        print(f"{n:45s} gpu-vs-f64 {(p.grad.cpu().double()-r.grad).abs().max().item()/s:.2e}  cpu32-vs-f64 {(q.grad.double()-r.grad).abs().max().item()/s:.2e}")
