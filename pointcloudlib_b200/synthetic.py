"""Synthetic ModelNet40 / ShapeNetPart shaped batches (no dataset is available offline).

Distribution: directions uniform on the unit sphere (CAD-surface stand-in), then the
reference loader's normalisation (data_utils/modelnet40_loader.py:121-125: subtract the centroid,
divide by the max norm) and, for training batches, its augmentation (:128-132: per-axis scale
U(2/3, 3/2), shift U(-0.2, 0.2)).  normals = the directions.  CPU generator, seeded.
"""
from __future__ import annotations

import torch


def modelnet_batch(B: int, N: int, seed: int = 0, augment: bool = True, n_classes: int = 40,
                   solid: bool = False):
    """-> xyz (B,N,3) f32, normals (B,N,3) f32, labels (B,) int64 — CPU tensors."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(B, N, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    xyz = d.clone()
    if solid:
        xyz = xyz * torch.rand(B, N, 1, generator=g).pow(1.0 / 3.0)
    xyz = xyz - xyz.mean(dim=1, keepdim=True)
    xyz = xyz / xyz.norm(dim=-1).max(dim=1).values.view(B, 1, 1)
    if augment:
        scale = torch.empty(B, 1, 3).uniform_(2.0 / 3.0, 3.0 / 2.0, generator=g)
        shift = torch.empty(B, 1, 3).uniform_(-0.2, 0.2, generator=g)
        xyz = xyz * scale + shift
    labels = torch.randint(0, n_classes, (B,), generator=g)
    return xyz.contiguous().float(), d.contiguous().float(), labels


def adversarial_cloud(B: int, N: int, seed: int = 0):
    """Parity stress input: exact duplicates, collinear runs and near-origin points (|p|^2<=1e-3)."""
    g = torch.Generator().manual_seed(seed)
    xyz = torch.rand(B, N, 3, generator=g) - 0.5
    q = max(N // 8, 1)
    xyz[:, q:2 * q] = xyz[:, :q]                                   # duplicates
    t = torch.linspace(-0.5, 0.5, q).view(1, q, 1)
    xyz[:, 2 * q:3 * q] = t * torch.tensor([1.0, 0.5, 0.25])       # collinear, equally spaced
    xyz[:, 3 * q:3 * q + max(q // 4, 1)] *= 0.02                   # inside the FPS skip radius
    xyz = (xyz * 64).round() / 64                                  # lattice -> many exact ties
    return xyz.contiguous().float()
