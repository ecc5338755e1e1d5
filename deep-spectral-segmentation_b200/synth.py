"""Seeded synthetic inputs for the deep-spectral hot path (images, features, VOC-shaped size tables).

There is no network and no dataset in the build/bench environment, so every measurement and parity test runs on
synthetic data of the shapes BASELINE.json names. The "blobs" image recipe gives a random-init
DINO ViT well separated Laplacian eigenvalues; i.i.d. noise images give tightly clustered ones.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# VOC-like (H, W) table with sampling weights (max side 500) used for BASELINE config 4
VOC_SHAPES = [((375, 500), 0.55), ((500, 375), 0.20), ((333, 500), 0.08), ((500, 333), 0.05),
              ((500, 500), 0.04), ((360, 480), 0.04), ((480, 360), 0.02), ((281, 500), 0.02)]


def blobs_image(H: int, W: int, seed: int) -> torch.Tensor:
    """uint8 RGB image (H, W, 3): smooth colour field + 2..5 uniform rectangles + N(0, 0.02) noise."""
    g = torch.Generator().manual_seed(int(seed))
    field = torch.rand(1, 3, 6, 6, generator=g)
    img = F.interpolate(field, size=(H, W), mode="bicubic", align_corners=False).clamp_(0, 1)[0]
    R = int(torch.randint(2, 6, (1,), generator=g))
    for _ in range(R):
        frac = 0.05 + 0.25 * float(torch.rand(1, generator=g))
        aspect = 0.5 + float(torch.rand(1, generator=g))
        h = int(min(H, max(1, round((frac * H * W * aspect) ** 0.5))))
        w = int(min(W, max(1, round(frac * H * W / h))))
        y0 = int(torch.randint(0, H - h + 1, (1,), generator=g))
        x0 = int(torch.randint(0, W - w + 1, (1,), generator=g))
        colour = torch.rand(3, 1, 1, generator=g)
        img[:, y0:y0 + h, x0:x0 + w] = colour
    img = img + 0.02 * torch.randn(3, H, W, generator=g)
    return (img.clamp_(0, 1) * 255.0).round_().to(torch.uint8).permute(1, 2, 0).contiguous()


def blobs_batch(n: int, H: int, W: int, seed0: int = 0) -> torch.Tensor:
    """(n, H, W, 3) uint8, image i uses seed seed0 + i."""
    return torch.stack([blobs_image(H, W, seed0 + i) for i in range(n)])


def structured_features(N: int, d: int, rank: int = 6, seed: int = 0) -> torch.Tensor:
    """(N, d) float32 features F = randn(N,d) + 3 randn(N,r) randn(r,d): eigen-gaps 0.008 .. 0.34."""
    g = torch.Generator().manual_seed(int(seed))
    return (torch.randn(N, d, generator=g) + 3.0 * torch.randn(N, rank, generator=g) @ torch.randn(rank, d, generator=g))


def clustered_features(N: int, d: int, n_clusters: int = 4, noise: float = 0.35, seed: int = 0) -> torch.Tensor:
    """(N, d) float32 features drawn around ``n_clusters`` random centres: a few well separated small eigenvalues."""
    g = torch.Generator().manual_seed(int(seed))
    centres = torch.randn(n_clusters, d, generator=g)
    assign = torch.randint(0, n_clusters, (N,), generator=g)
    return centres[assign] + noise * torch.randn(N, d, generator=g) * (d ** 0.5) / (d ** 0.5)


def voc_shapes(n: int, seed: int = 0):
    """n (H, W) pairs drawn from VOC_SHAPES with a seeded generator."""
    rng = np.random.default_rng(seed)
    p = np.array([w for _, w in VOC_SHAPES], np.float64)
    idx = rng.choice(len(VOC_SHAPES), size=n, p=p / p.sum())
    return [VOC_SHAPES[i][0] for i in idx]
