"""Host side of the segmentation kernels (csrc/segment.cu): the first consumers of the eigenvectors in the reference
(extract/extract.py:283-349 multi-region K-means, :364-390 single-region threshold), run on the device tensors the
eigensolver produced -- no .pth round trip in between when called from the fused pipeline."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


@torch.no_grad()
def threshold_masks(evecs: torch.Tensor, threshold: float = 0.0, which: int = 1) -> torch.Tensor:
    """evecs [B, K, N] fp32 CUDA -> uint8 [B, N]: 255 where eigenvector ``which`` > threshold, else 0 -- the pixels of
    ``Image.fromarray(eigenvector > threshold).convert('L')`` (extract.py:383-387)."""
    _lib.require_cuda(evecs, "evecs")
    ev = evecs.to(torch.float32).contiguous()
    B, K, N = ev.shape
    with torch.cuda.device(ev.device):
        mask = torch.empty(B, N, dtype=torch.uint8, device=ev.device)
        _lib.check(_lib.load().dss_segment_threshold(ev.data_ptr(), B, K, N, int(which), float(threshold), mask.data_ptr(),
                                                     _lib.stream_ptr(ev.device)), "dss_segment_threshold")
    return mask


def adaptive_num_clusters(eigenvalues) -> int:
    """extract.py:312-316: position of the largest eigengap, the gap after the constant vector excluded."""
    ev = np.asarray(eigenvalues, dtype=np.float64)
    by_gap = np.argsort(np.diff(ev))[::-1]
    return int(by_gap[by_gap != 0][0] + 1)


@torch.no_grad()
def kmeans_labels(points: torch.Tensor, n_clusters, grid: Optional[Tuple[int, int]] = None, infer_bg_index: bool = True,
                  layout: str = "eigenvectors", seed: int = 0, max_iter: int = 300, tol: float = 1e-4,
                  image_keys: Optional[Sequence[int]] = None):
    """Batched K-means on the device.

    layout='eigenvectors': points [B, dims, N] (rows of an eigenvector stack, e.g. evecs[:, 1:1+m]) -- point n has the
    coordinates points[b, :, n]; layout='features': points [B, N, dims] (kmeans_baseline, extract.py:323-325).
    n_clusters: int or a length-B sequence (the reference's adaptive mode). grid=(H, W) with H*W == N enables the
    background rule (the label with the largest border share becomes 0, extract.py:337-345).
    image_keys (length B): keys of the device random generator, e.g. the data-set index of each image, so that an
    image is clustered identically whatever batch it arrives in (default: the position in the batch).
    Returns (labels uint8 [B, N], info int32 [B, 2] = {iterations, converged}, inertia fp32 [B])."""
    _lib.require_cuda(points, "points")
    pts = points.to(torch.float32)
    if layout == "eigenvectors":
        B, dims, N = pts.shape
        if pts.stride(2) != 1:
            pts = pts.contiguous()
        strides = (pts.stride(0), 1, pts.stride(1))
    elif layout == "features":
        pts = pts.contiguous()
        B, N, dims = pts.shape
        strides = (N * dims, dims, 1)
    else:
        raise ValueError(layout)
    ks = [int(n_clusters)] * B if np.isscalar(n_clusters) else [int(k) for k in n_clusters]
    assert len(ks) == B and min(ks) >= 1
    dev = pts.device
    with torch.cuda.device(dev):
        kdev = torch.tensor(ks, dtype=torch.int32, device=dev)
        keys = None
        if image_keys is not None:
            assert len(image_keys) == B
            keys = torch.tensor([int(k) & 0x7FFFFFFF for k in image_keys], dtype=torch.int32, device=dev)
        labels = torch.empty(B, N, dtype=torch.uint8, device=dev)
        info = torch.empty(B, 2, dtype=torch.int32, device=dev)
        inertia = torch.empty(B, dtype=torch.float32, device=dev)
        gh, gw = grid if grid is not None else (1, N)
        use_bg = bool(infer_bg_index and grid is not None)
        _lib.check(_lib.load().dss_segment_kmeans(pts.data_ptr(), strides[0], strides[1], strides[2], B, N, dims,
                                                  kdev.data_ptr(), _lib.ptr(keys), max(ks), int(gh), int(gw),
                                                  1 if use_bg else 0,
                                                  int(seed) & 0xFFFFFFFF, int(max_iter), float(tol), labels.data_ptr(),
                                                  info.data_ptr(), inertia.data_ptr(), _lib.stream_ptr(dev)),
                   "dss_segment_kmeans")
    return labels, info, inertia
