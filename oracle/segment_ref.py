"""ORACLE (test infrastructure, not product): CPU restatement of the reference's two segmentation workers,
/root/reference/extract/extract.py:283-349 (``_extract_multi_region_segmentations``) and :364-390
(``_extract_single_region_segmentations``), plus extract_utils.py:124-135 (``get_border_fraction``), from the loaded
dicts to the array the reference hands to ``Image.fromarray(...).convert('L').save``.

The multi-region worker clusters with scikit-learn's KMeans, unseeded in the reference (k-means++ draws from numpy's
global RNG): label numbers are not reproducible there, only the partition on well-separated data. Pinning: the
fixtures tests/golden/seg_*.npz are produced by the reference's OWN functions (oracle/make_golden.py, numpy RNG seeded
before each call); tests/test_cpu_oracle.py checks this restatement against them and against the live reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module."""
from __future__ import annotations

import numpy as np


def get_border_fraction(segmap: np.ndarray):
    """extract_utils.py:124-135."""
    num_border_pixels = 2 * (segmap.shape[0] + segmap.shape[1])
    counts_map = {idx: 0 for idx in np.unique(segmap)}
    for border in [segmap[:, 0], segmap[:, -1], segmap[0, :], segmap[-1, :]]:
        unique, counts = np.unique(border, return_counts=True)
        for idx, count in zip(unique.tolist(), counts.tolist()):
            counts_map[idx] += count
    indices = np.array(list(counts_map.keys()))
    normlized_counts = np.array(list(counts_map.values())) / num_border_pixels
    return indices, normlized_counts


def single_region(eigenvectors: np.ndarray, H_patch: int, W_patch: int, threshold: float = 0.0) -> np.ndarray:
    """extract.py:383-387 -> the uint8 'L' image (0 / 255) PIL makes of the boolean map."""
    segmap = (np.asarray(eigenvectors)[1] > threshold).reshape(H_patch, W_patch)
    return (segmap * 255).astype(np.uint8)


def multi_region(eigenvalues: np.ndarray, eigenvectors: np.ndarray, feats_k, H_patch: int, W_patch: int,
                 adaptive: bool = False, non_adaptive_num_segments: int = 4, infer_bg_index: bool = True,
                 kmeans_baseline: bool = False, num_eigenvectors: int = 1_000_000, random_state=None) -> np.ndarray:
    """extract.py:306-348 -> label map (uint8 after convert('L'))."""
    from sklearn.cluster import KMeans
    if adaptive:
        indices_by_gap = np.argsort(np.diff(np.asarray(eigenvalues)))[::-1]
        index_largest_gap = indices_by_gap[indices_by_gap != 0][0]
        n_clusters = index_largest_gap + 1
    else:
        n_clusters = non_adaptive_num_segments
    kmeans = KMeans(n_clusters=n_clusters) if random_state is None else KMeans(n_clusters=n_clusters,
                                                                                random_state=random_state)
    if kmeans_baseline:
        clusters = kmeans.fit_predict(np.asarray(feats_k).squeeze())
    else:
        clusters = kmeans.fit_predict(np.asarray(eigenvectors)[1:1 + num_eigenvectors].T)
    if clusters.size == H_patch * W_patch:
        segmap = clusters.reshape(H_patch, W_patch)
    elif clusters.size == H_patch * W_patch * 4:
        segmap = clusters.reshape(H_patch * 2, W_patch * 2)
    else:
        raise ValueError()
    if infer_bg_index:
        indices, normlized_counts = get_border_fraction(segmap)
        bg_index = indices[np.argmax(normlized_counts)].item()
        bg_region = (segmap == bg_index)
        zero_region = (segmap == 0)
        segmap[bg_region] = 0
        segmap[zero_region] = bg_index
    return np.clip(segmap, 0, 255).astype(np.uint8)


def same_partition(a: np.ndarray, b: np.ndarray) -> bool:
    """True iff the two label maps induce the same partition (labels may be permuted)."""
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    if a.shape != b.shape:
        return False
    pairs = set(zip(a.tolist(), b.tolist()))
    return len(pairs) == len(set(a.tolist())) == len(set(b.tolist()))
