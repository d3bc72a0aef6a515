"""Literal torch-CPU restatement of the reference's module graphs (TEST INFRASTRUCTURE ONLY).

Each function follows the reference's op sequence line by line — materialised gathers, the two
transposes, 1x1 Conv -> BatchNorm(train) -> ReLU on (B,C,S,ns), max over dim 2 — with the index
ops taken from the C oracle (oracle/pcl_oracle.c).  Parameters are read from a CPU copy of the
product model (same state-dict structure as the reference's), so the product's fused / re-ordered
arithmetic is checked against the reference's own order of operations.

PARITY UNPINNED for the dense layers: Conv/BatchNorm numerics live in Jittor (un-vendored, no
version pinned); BatchNorm(train) is taken as biased batch statistics, eps 1e-5.
"""
from __future__ import annotations

import numpy as np
import torch

import oracle


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def index_points_t(points: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """misc/ops.py:12-27 with torch fancy indexing (differentiable)."""
    B = points.shape[0]
    view_shape = [B] + [1] * (idx.dim() - 1)
    batch_indices = torch.arange(B).view(view_shape).expand_as(idx)
    return points[batch_indices, idx.long(), :]


def furthest_point_sampler(x: torch.Tensor, n_samples: int) -> torch.Tensor:
    """misc/ops.py:257-286."""
    idx = _t(oracle.fps(x.detach().numpy(), n_samples))
    return index_points_t(x, idx)


def ball_query_grouper(new_xyz, pointset, feature, radius, n_samples, use_xyz):
    """misc/ops.py:345-407."""
    idx, _ = oracle.ball_query(new_xyz.detach().numpy(), pointset.detach().numpy(),
                               float(str(radius)), n_samples)
    idx = _t(idx)
    new_pointset = index_points_t(pointset, idx)                      # (B,S,ns,3)
    new_feature = index_points_t(feature, idx) if feature is not None else None
    if use_xyz:
        local_xyz = new_pointset - new_xyz.unsqueeze(dim=2)
        new_feature = torch.cat([local_xyz, new_feature], dim=-1) if new_feature is not None else local_xyz
    return new_feature


def group_all(new_xyz, pointset, feature):
    """misc/ops.py:415-419."""
    return torch.cat([pointset, feature], dim=-1).unsqueeze(dim=1)


def pointnet_sa(module, xyz, feature, seg: bool = False):
    """networks/cls/pointnet2.py:33-62 (seg=True: networks/seg/pointnet2_partseg.py:42-72)."""
    if module.n_points is not None:
        new_xyz = furthest_point_sampler(xyz, module.n_points)
    else:
        new_xyz = torch.zeros((xyz.shape[0], 1, 3)) if seg else None
    outs = []
    for i, grouper in enumerate(module.groupers):
        if module.n_points is not None:
            nf = ball_query_grouper(new_xyz, xyz, feature, grouper.radius, grouper.n_samples,
                                    grouper.use_xyz)
        else:
            nf = group_all(new_xyz, xyz, feature)
        nf = nf.permute(0, 3, 1, 2)            # [B, C, n_points, n_samples]
        nf = module.mlps[i](nf)                # Conv2d 1x1 -> BatchNorm2d -> ReLU (torch CPU)
        nf = nf.permute(0, 2, 3, 1)            # [B, n_points, n_samples, C]
        nf = nf.max(dim=2).values              # jittor argmax(dim=2)[1] == max values
        outs.append(nf)
    return new_xyz, torch.cat(outs, dim=-1)


def pointnet2_cls(model, xyz, feature):
    """networks/cls/pointnet2.py:149-158."""
    for module in model.pointnet_modules:
        xyz, feature = pointnet_sa(module, xyz, feature)
    return model.fc_layer(feature.squeeze(dim=1))


def feature_propagation(fp, xyz1, xyz2, points1, points2):
    """misc/ops.py:66-107."""
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    if S == 1:
        interpolated = points2.repeat(1, N, 1)
    else:
        idx, dist, _w = oracle.three_nn(xyz1.detach().numpy(), xyz2.detach().numpy())
        idx, dists = _t(idx), _t(dist)
        dist_recip = 1.0 / (dists + 1e-8)
        norm = torch.sum(dist_recip, dim=2, keepdim=True)
        weight = dist_recip / norm
        interpolated = torch.sum(index_points_t(points2, idx) * weight.view(B, N, 3, 1), dim=2)
    new_points = torch.cat([points1, interpolated], dim=-1) if points1 is not None else interpolated
    new_points = new_points.permute(0, 2, 1)
    for i, conv in enumerate(fp.mlp_convs):
        new_points = fp.relu(fp.mlp_bns[i](conv(new_points)))
    return new_points.permute(0, 2, 1)


def pointnet2_partseg(model, xyz, feature, cls_label):
    """networks/seg/pointnet2_partseg.py:158-176."""
    B, N, _ = xyz.shape
    l1_xyz, l1_f = pointnet_sa(model.pointnet_modules[0], xyz, feature, seg=True)
    l2_xyz, l2_f = pointnet_sa(model.pointnet_modules[1], l1_xyz, l1_f, seg=True)
    l3_xyz, l3_f = pointnet_sa(model.pointnet_modules[2], l2_xyz, l2_f, seg=True)
    l2_f = feature_propagation(model.fp3, l2_xyz, l3_xyz, l2_f, l3_f)
    l1_f = feature_propagation(model.fp2, l1_xyz, l2_xyz, l1_f, l2_f)
    onehot = cls_label.view(B, 16, 1).repeat(1, 1, N).permute(0, 2, 1)
    f = feature_propagation(model.fp1, xyz, l1_xyz, torch.cat([onehot, xyz, feature], 2), l1_f)
    return model.fc_layer(f.permute(0, 2, 1))


def get_graph_feature(x, k):
    """networks/cls/dgcnn.py:29-50 with idx from the KNN oracle."""
    B, C, N = x.shape
    idx = _t(oracle.knn(x.detach().numpy(), x.detach().numpy(), k)).permute(0, 2, 1).long()
    idx = (idx + torch.arange(B).view(-1, 1, 1) * N).reshape(-1)
    xt = x.transpose(2, 1)
    feature = xt.reshape(B * N, -1)[idx, :].reshape(B, N, k, C)
    xr = xt.reshape(B, N, 1, C).repeat(1, 1, k, 1)
    return torch.cat((feature - xr, xr), dim=3).permute(0, 3, 1, 2)


def dgcnn(model, x):
    """networks/cls/dgcnn.py:95-122."""
    import torch.nn.functional as TF
    B = x.shape[0]
    feats = []
    h = x
    for conv in (model.conv1, model.conv2, model.conv3, model.conv4):
        h = conv(get_graph_feature(h, model.k)).max(dim=-1).values
        feats.append(h)
    h = model.conv5(torch.cat(feats, dim=1))
    x1 = h.max(dim=2).values.reshape(B, -1)
    x2 = h.mean(dim=2).reshape(B, -1)
    h = torch.cat((x1, x2), 1)
    h = model.dp1(TF.leaky_relu(model.bn6(model.linear1(h)), 0.2))
    h = model.dp2(TF.leaky_relu(model.bn7(model.linear2(h)), 0.2))
    return model.linear3(h)


def soft_cross_entropy_loss(output, target, smoothing=True):
    """train_cls.py:31-51 (label smoothing eps = 0.2)."""
    target = target.view(-1)
    if smoothing:
        eps = 0.2
        n_class = output.shape[1]
        one_hot = torch.zeros_like(output)
        for i in range(output.shape[0]):
            one_hot[i, int(target[i])] = 1
        one_hot = one_hot * (1 - eps) + (1 - one_hot) * eps / (n_class - 1)
        log_prb = torch.log(torch.softmax(output, dim=1))
        return -(one_hot * log_prb).sum(dim=1).mean()
    return torch.nn.functional.cross_entropy(output, target)


# ---------------------------------------------------------------------------------------------
# PointConv (misc/pointconv_utils.py, networks/cls/pointconv.py)
# ---------------------------------------------------------------------------------------------
def square_distance_t(src, dst):
    """misc/pointconv_utils.py:34-53 (differentiable torch form)."""
    B, N, _ = src.shape
    M = dst.shape[1]
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist = dist + torch.sum(src ** 2, -1).view(B, N, 1)
    dist = dist + torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist


def pointconv_sample_and_group(npoint, nsample, xyz, points, density_scale):
    """misc/pointconv_utils.py:133-170; the FPS start index comes from numpy's global RNG (:88)."""
    B, N, C = xyz.shape
    start = np.random.randint(0, N, B, dtype="l").astype(np.int32)
    fps_idx = _t(oracle.fps_pointconv(xyz.detach().numpy(), npoint, start))
    new_xyz = index_points_t(xyz, fps_idx)
    idx = _t(oracle.knn_point(nsample, xyz.detach().numpy(), new_xyz.detach().numpy()))
    grouped_xyz = index_points_t(xyz, idx)
    grouped_xyz_norm = grouped_xyz - new_xyz.view(B, npoint, 1, C)
    new_points = torch.cat([grouped_xyz_norm, index_points_t(points, idx)], dim=-1) \
        if points is not None else grouped_xyz_norm
    grouped_density = index_points_t(density_scale, idx)
    return new_xyz, new_points, grouped_xyz_norm, idx, grouped_density


def pointconv_sa(mod, xyz, points):
    """misc/pointconv_utils.py:361-400 (PointConvDensitySetAbstraction.execute)."""
    B, _, N = xyz.shape
    xyz = xyz.permute(0, 2, 1)
    if points is not None:
        points = points.permute(0, 2, 1)
    sq = square_distance_t(xyz, xyz)                                        # :179
    xyz_density = (torch.exp(-sq / (2.0 * mod.bandwidth * mod.bandwidth)) / (2.5 * mod.bandwidth)).mean(dim=-1)
    density_scale = mod.densitynet(xyz_density)
    if mod.group_all:
        new_xyz = torch.zeros((B, 1, 3), dtype=xyz.dtype)
        grouped_xyz_norm = xyz.view(B, 1, N, 3)
        new_points = torch.cat([grouped_xyz_norm, points.view(B, 1, N, -1)], dim=-1) \
            if points is not None else grouped_xyz_norm
        grouped_density = density_scale.reshape(B, N, 1).view(B, 1, N, 1)
    else:
        new_xyz, new_points, grouped_xyz_norm, _, grouped_density = pointconv_sample_and_group(
            mod.npoint, mod.nsample, xyz, points, density_scale.reshape(B, N, 1))
    new_points = new_points.permute(0, 3, 2, 1)
    for i in range(len(mod.mlp_convs)):
        new_points = mod.relu(mod.mlp_bns[i](mod.mlp_convs[i](new_points)))
    weights = mod.weightnet(grouped_xyz_norm.permute(0, 3, 2, 1))
    new_points = new_points * grouped_density.permute(0, 3, 2, 1)
    new_points = torch.matmul(new_points.permute(0, 3, 1, 2),
                              weights.permute(0, 3, 2, 1)).reshape(B, mod.npoint, -1)
    new_points = mod.linear(new_points)
    new_points = mod.relu(mod.bn_linear(new_points.permute(0, 2, 1)))
    return new_xyz.permute(0, 2, 1), new_points


def _density_conv_tail(mod, B, S, new_points, grouped_xyz_norm, grouped_density):
    """misc/pointconv_utils.py:384-397 == :307-321 (the two modules share this tail)."""
    new_points = new_points.permute(0, 3, 2, 1)
    for i in range(len(mod.mlp_convs)):
        new_points = mod.relu(mod.mlp_bns[i](mod.mlp_convs[i](new_points)))
    weights = mod.weightnet(grouped_xyz_norm.permute(0, 3, 2, 1))
    new_points = new_points * grouped_density.permute(0, 3, 2, 1)
    new_points = torch.matmul(new_points.permute(0, 3, 1, 2),
                              weights.permute(0, 3, 2, 1)).reshape(B, S, -1)
    new_points = mod.linear(new_points)
    return mod.relu(mod.bn_linear(new_points.permute(0, 2, 1)))


def pointconv_interp(mod, xyz1, xyz2, points1, points2):
    """misc/pointconv_utils.py:274-321 (PointConvDensitySetInterpolation.execute): 3-NN inverse-distance
    interpolation of points2 onto xyz1 (points1 is permuted at :289 and never used), KDE density ->
    DensityNet, sample_and_group with npoint = N (FPS over ALL points: a permutation from a random
    start), shared MLP, weight net, density-weighted (C x ns).(ns x 16) product, Linear + BN + ReLU."""
    xyz1 = xyz1.permute(0, 2, 1)
    xyz2 = xyz2.permute(0, 2, 1)
    points2 = points2.permute(0, 2, 1)
    B, N, _ = xyz1.shape
    idx, dist, _w = oracle.three_nn(xyz1.detach().numpy(), xyz2.detach().numpy())
    idx, dists = _t(idx), _t(dist).to(xyz1.dtype)
    dist_recip = 1.0 / (dists + 1e-8)
    weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
    interpolated = torch.sum(index_points_t(points2, idx) * weight.view(B, N, 3, 1), dim=2)
    sq = square_distance_t(xyz1, xyz1)
    xyz_density = (torch.exp(-sq / (2.0 * mod.bandwidth * mod.bandwidth)) / (2.5 * mod.bandwidth)).mean(dim=-1)
    density_scale = mod.densitynet(xyz_density)
    _new_xyz, new_points, grouped_xyz_norm, _, grouped_density = pointconv_sample_and_group(
        N, mod.nsample, xyz1, interpolated, density_scale.reshape(B, N, 1))
    return _density_conv_tail(mod, B, N, new_points, grouped_xyz_norm, grouped_density)


def pointconv_partseg(model, xyz, cls_label=None):
    """networks/seg/pointconv_partseg.py:42-63 (cls_label is accepted and ignored there)."""
    xyz = xyz.permute(0, 2, 1)
    l1_xyz, l1_points = pointconv_sa(model.sa0, xyz, None)
    l2_xyz, l2_points = pointconv_sa(model.sa1, l1_xyz, l1_points)
    l3_xyz, l3_points = pointconv_sa(model.sa2, l2_xyz, l2_points)
    l4_xyz, l4_points = pointconv_sa(model.sa3, l3_xyz, l3_points)
    l3_points = pointconv_interp(model.in0, l3_xyz, l4_xyz, l3_points, l4_points)
    l2_points = pointconv_interp(model.in1, l2_xyz, l3_xyz, l2_points, l3_points)
    l1_points = pointconv_interp(model.in2, l1_xyz, l2_xyz, l1_points, l2_points)
    l0_points = pointconv_interp(model.in3, xyz, l1_xyz, xyz, l1_points)
    x = model.drop1(model.relu(model.bn1(model.fc1(l0_points))))
    return model.fc3(x).permute(0, 2, 1)


def pointconv_cls(model, xyz):
    """networks/cls/pointconv.py:23-34."""
    xyz = xyz.permute(0, 2, 1)
    B = xyz.shape[0]
    l1_xyz, l1_points = pointconv_sa(model.sa1, xyz, None)
    l2_xyz, l2_points = pointconv_sa(model.sa2, l1_xyz, l1_points)
    l3_xyz, l3_points = pointconv_sa(model.sa3, l2_xyz, l2_points)
    x = l3_points.reshape(B, 1024)
    x = model.drop1(model.relu(model.bn1(model.fc1(x))))
    x = model.drop2(model.relu(model.bn2(model.fc2(x))))
    return model.fc3(x)


def dgcnn_partseg(model, x, l):
    """networks/seg/dgcnn_partseg.py:84-128."""
    B, _, N = x.shape
    h = get_graph_feature(x, model.k)
    h = model.conv2(model.conv1(h))
    x1 = h.max(dim=-1).values
    h = get_graph_feature(x1, model.k)
    h = model.conv4(model.conv3(h))
    x2 = h.max(dim=-1).values
    h = get_graph_feature(x2, model.k)
    h = model.conv5(h)
    x3 = h.max(dim=-1).values
    h = model.conv6(torch.cat((x1, x2, x3), dim=1)).max(dim=-1, keepdim=True).values
    ll = model.conv7(l.view(B, -1, 1))
    h = torch.cat((h, ll), dim=1).repeat(1, 1, N)
    h = torch.cat((h, x1, x2, x3), dim=1)
    h = model.dp1(model.conv8(h))
    h = model.dp2(model.conv9(h))
    return model.conv11(model.conv10(h))


def pointcnn_cls(model, xyz, normal=None):
    """networks/cls/pointcnn.py:34-47 + misc/layers.py:306-407 evaluated on the CPU: the dense
    layers are the model's own modules (torch CPU, any dtype); sampling and neighbourhoods come from
    the C oracle (oracle.fps / oracle.knn incl. the dilation slice), the regional gather is the
    reference's per-sample fancy index (layers.py:381-388)."""
    def rand_pointcnn(m, pts, fts):
        if 0 < m.P < pts.shape[1]:
            rep = furthest_point_sampler(pts, m.P)
        else:
            rep = pts
        pc = m.pointcnn
        f = pc.dense(fts) if fts is not None else fts
        q = rep.permute(0, 2, 1).contiguous().detach().float().numpy()
        r = pts.permute(0, 2, 1).contiguous().detach().float().numpy()
        idx = _t(oracle.knn(q, r, pc.K * pc.D))[:, 0::pc.D, :].permute(0, 2, 1)   # (N, P, K)
        region = lambda t: torch.stack([t[n][i.long(), :] for n, i in enumerate(torch.unbind(idx, dim=0))], dim=0)
        return rep, pc.x_conv((rep, region(pts), region(f) if f is not None else f))

    x = (xyz, xyz if normal is None else normal)
    x = rand_pointcnn(model.pcnn1, *x)
    for m in model.pcnn2:
        x = rand_pointcnn(m, *x)
    logits = model.fcn(x[1].permute(0, 2, 1))
    return torch.mean(logits, dim=2)
