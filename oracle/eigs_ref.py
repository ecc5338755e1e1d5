"""ORACLE (test infrastructure, not product): CPU restatement of the reference's ``_extract_eig`` arithmetic.

Follows /root/reference/extract/extract.py:119-244 line by line (numpy / scipy / torch-CPU):

  :146-148  feats = data[which_features].squeeze(); F.normalize(p=2, dim=-1)
  :160-163  which_matrix == 'affinity_svd'  -> torch.linalg.svd
  :166-172  which_matrix == 'affinity'      -> eigsh(W, which='LM', k=K)
  :175-235  which_matrix == 'laplacian'     -> W_feat, optional colour affinity, degree, eigsh(D-W, sigma=0, M=D)
  :237-240  sign rule
and the helpers extract/extract_utils.py:151-188 (knn_affinity) and :207-220 (get_diagonal).

The eigen-arithmetic itself lives in scipy (``scipy.sparse.linalg.eigsh`` -> ARPACK ssaupd/sseupd + LAPACK
sgetrf), which the reference leaves unpinned (requirements.txt:6); this container has scipy 1.18.1.
pymatting (unpinned, not installed) supplies ``knn`` and ``row_sum``; their published behaviour is restated in
``knn_exact`` / ``row_sum`` below (exact k-nearest neighbours including the query itself; A.dot(ones)).

Pinning: the reference ships no tests or golden vectors, so this restatement is pinned against the reference's
own function run in this container (oracle/ref_shim.py imports the real ``_extract_eig``); the outputs are
committed under tests/golden/ by oracle/make_golden.py and checked by tests/test_oracle_golden.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse
import torch
import torch.nn.functional as F
from scipy.sparse.linalg import eigsh


def row_sum(A):
    """pymatting.util.util.row_sum restated: A.dot(ones(A.shape[1], A.dtype))."""
    return A.dot(np.ones(A.shape[1], A.dtype))


def knn_exact(data: np.ndarray, query: np.ndarray, k: int):
    """pymatting.util.kdtree.knn restated: exact k nearest neighbours (euclidean, float32 points), self included.

    Ties are broken by the lower point index (pymatting's KD-tree order is implementation defined; the tests use
    inputs without distance ties at the k-th neighbour).
    Returns (distances float32 (n,k), indices int (n,k)) sorted by (distance, index).
    """
    data = np.asarray(data, np.float32)
    query = np.asarray(query, np.float32)
    n = query.shape[0]
    dist = np.empty((n, k), np.float32)
    idx = np.empty((n, k), np.int64)
    step = max(1, (1 << 24) // max(1, data.shape[0]))
    for s in range(0, n, step):
        q = query[s:s + step]
        # squared distance accumulated in float32, dimension by dimension (matches the CUDA kernel's order)
        d2 = np.zeros((q.shape[0], data.shape[0]), np.float32)
        for c in range(data.shape[1]):
            diff = q[:, c:c + 1] - data[None, :, c]
            d2 += diff * diff
        order = np.lexsort((np.broadcast_to(np.arange(data.shape[0]), d2.shape), d2), axis=1)[:, :k]
        idx[s:s + step] = order
        dist[s:s + step] = np.sqrt(np.take_along_axis(d2, order, axis=1))
    return dist, idx


def knn_affinity(image, n_neighbors=(20, 10), distance_weights=(2.0, 0.1), knn=knn_exact):
    """extract_utils.py:151-188 restated. image (h, w, 3) float in [0,1] -> csr (n, n) float64."""
    h, w = image.shape[:2]
    r, g, b = image.reshape(-1, 3).T
    n = w * h
    x = np.tile(np.linspace(0, 1, w), h)
    y = np.repeat(np.linspace(0, 1, h), w)
    i, j = [], []
    for k, dw in zip(n_neighbors, distance_weights):
        f = np.stack([r, g, b, dw * x, dw * y], axis=1, out=np.zeros((n, 5), dtype=np.float32))
        _, neighbors = knn(f, f, k=k)
        i.append(np.repeat(np.arange(n), k))
        j.append(neighbors.flatten())
    ij = np.concatenate(i + j)
    ji = np.concatenate(j + i)
    coo_data = np.ones(2 * sum(n_neighbors) * n)
    return scipy.sparse.csr_matrix((coo_data, (ij, ji)), (n, n))


def rw_laplacian_values(image, sigma, r):
    """pymatting.laplacian.rw_laplacian._rw_laplacian restated (pymatting is unpinned in requirements.txt:10 and not
    installed here; restated from its published source, releases 1.0-1.1): for every pixel and every offset in
    [-r, r]^2 the CLAMPED neighbour gets exp(-900 * ||z_i - z_j||^2). The published code hard-codes 900 and never
    uses ``sigma`` (1 / 0.033^2 = 918); PARITY UNPINNED for this constant -- the CUDA entry point takes it as an
    argument. Returns (values f8[m], i_inds i4[m], j_inds i4[m]), m = n (2r+1)^2, in pymatting's order."""
    image = np.asarray(image, np.float64)
    h, w = image.shape[:2]
    n = h * w
    ys, xs = np.mgrid[0:h, 0:w]
    vals, ii, jj = [], [], []
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            x2 = np.clip(xs + dx, 0, w - 1)
            y2 = np.clip(ys + dy, 0, h - 1)
            diff = image[ys, xs] - image[y2, x2]
            nrm = np.sqrt((diff * diff).sum(-1))
            vals.append(np.exp(-900 * nrm ** 2))
            ii.append(xs + ys * w)
            jj.append(x2 + y2 * w)
    # pymatting's loop order is (y, x, dy, dx): offsets vary fastest
    order = lambda a: np.stack(a, axis=-1).reshape(n * (2 * r + 1) ** 2)
    return order(vals), order(ii).astype(np.int32), order(jj).astype(np.int32)


def rw_affinity(image, sigma=0.033, radius=1, rw=rw_laplacian_values):
    """extract_utils.py:191-204 restated. image (h, w, 3) float64 in [0,1] -> csr (n, n) float64 (duplicates summed)."""
    h, w = image.shape[:2]
    n = h * w
    values, i_inds, j_inds = rw(image, sigma, radius)
    return scipy.sparse.csr_matrix((values, (i_inds, j_inds)), shape=(n, n))


def get_diagonal(W, threshold: float = 1e-12):
    """extract_utils.py:207-220 restated."""
    D = row_sum(W)
    D[D < threshold] = 1.0
    return scipy.sparse.diags(D)


def image_sizes(shape, patch_size):
    """extract_utils.py:73-79 restated."""
    B, C, H, W = shape
    assert B == 1
    Hp, Wp = H // patch_size, W // patch_size
    return B, C, H, W, patch_size, Hp, Wp, Hp * patch_size, Wp * patch_size


def sign_rule_(eigenvectors: torch.Tensor) -> torch.Tensor:
    """extract.py:237-240: flip v_k iff 0.5 < mean(v_k > 0) < 1.0 (in place)."""
    for k in range(eigenvectors.shape[0]):
        if 0.5 < torch.mean((eigenvectors[k] > 0).float()).item() < 1.0:
            eigenvectors[k] = 0 - eigenvectors[k]
    return eigenvectors


def affinity_matrices(feats: torch.Tensor, normalize=True, threshold_at_zero=True, image_lr=None,
                      image_color_lambda=0.0, knn=knn_exact, which_color_matrix="knn"):
    """extract.py:148,191-222 restated -> (W_comb float32 (n,n) ndarray, D_comb float32 dense diag (n,n))."""
    feats = feats.squeeze()
    if normalize:
        feats = F.normalize(feats, p=2, dim=-1)
    W_feat = feats @ feats.T
    if threshold_at_zero:
        W_feat = W_feat * (W_feat > 0)
    W_feat = W_feat / W_feat.max()
    W_feat = W_feat.cpu().numpy()
    if image_color_lambda > 0:
        if which_color_matrix == "knn":
            W_lr = knn_affinity(image_lr, knn=knn)
        elif which_color_matrix == "rw":
            W_lr = rw_affinity(image_lr)
        else:
            raise ValueError(which_color_matrix)   # the reference leaves W_lr undefined here (UnboundLocalError)
        W_color = np.array(W_lr.todense().astype(np.float32))
    else:
        W_color = 0
    W_comb = W_feat + W_color * image_color_lambda
    D_comb = np.array(get_diagonal(W_comb).todense())
    return W_comb, D_comb


def upsample_features(feats: torch.Tensor, grid, lr_size) -> torch.Tensor:
    """extract.py:184-188 restated: (N, d) features on a (H_patch, W_patch) grid -> (H_lr*W_lr, d), bilinear."""
    (Hp, Wp), (Hl, Wl) = grid, lr_size
    return F.interpolate(feats.T.reshape(1, -1, Hp, Wp), size=(Hl, Wl), mode="bilinear",
                         align_corners=False).reshape(-1, Hl * Wl).T


def extract_eig(feats: torch.Tensor, K: int, which_matrix="laplacian", normalize=True, lapnorm=True,
                threshold_at_zero=True, image_lr=None, image_color_lambda=0.0, knn=knn_exact, rng_seed=None,
                grid=None, lr_size=None, which_color_matrix="knn", stats=None):
    """The arithmetic of reference ``_extract_eig`` from a feature tensor to (eigenvalues, eigenvectors).

    ``feats`` is the (1, N, d) / (N, d) float32 ``data_dict['k']``; ``image_lr`` the (H_lr, W_lr, 3) float64 /255
    low-resolution image the reference builds at extract.py:199-204 (only used if image_color_lambda > 0).
    ``rng_seed``: scipy>=1.15 draws ARPACK's start vector from ``rng``; the reference leaves it unseeded.
    Returns (eigenvalues, eigenvectors (K, N) float32 torch) exactly as saved at extract.py:242-244.
    """
    kw = {} if rng_seed is None else {"rng": np.random.default_rng(rng_seed)}
    feats = feats.squeeze()
    if which_matrix == "affinity_svd":
        f = F.normalize(feats, p=2, dim=-1) if normalize else feats
        USV = torch.linalg.svd(f, full_matrices=False)
        eigenvectors = USV[0][:, :K].T
        eigenvalues = USV[1][:K]
    elif which_matrix == "affinity":
        f = F.normalize(feats, p=2, dim=-1) if normalize else feats
        W = f @ f.T
        if threshold_at_zero:
            W = W * (W > 0)
        W = W.cpu().numpy()
        eigenvalues, eigenvectors = eigsh(W, which="LM", k=K, **kw)
        eigenvectors = torch.flip(torch.from_numpy(eigenvectors), dims=(-1,)).T
    elif which_matrix in ("matting_laplacian", "laplacian"):
        if grid is not None and lr_size is not None and tuple(grid) != tuple(lr_size):
            # extract.py:148 normalises BEFORE the up-sampling at :184-188 and never re-normalises
            if normalize:
                feats = F.normalize(feats, p=2, dim=-1)
                normalize = False
            feats = upsample_features(feats, grid, lr_size)
        W_comb, D_comb = affinity_matrices(feats, normalize, threshold_at_zero, image_lr, image_color_lambda, knn,
                                           which_color_matrix)
        # ``stats`` (optional dict) records which route the reference's try/except took: D - W is singular, so its
        # float32 LU can hit an exactly-zero pivot; the shift-invert solve then yields NaN/inf, ARPACK raises and the
        # reference falls back to which='SM' (extract.py:226-229)
        if lapnorm:
            try:
                eigenvalues, eigenvectors = eigsh(D_comb - W_comb, k=K, sigma=0, which="LM", M=D_comb, **kw)
                if stats is not None:
                    stats["route"] = "shift-invert"
            except Exception:
                eigenvalues, eigenvectors = eigsh(D_comb - W_comb, k=K, which="SM", M=D_comb, **kw)
                if stats is not None:
                    stats["route"] = "SM-fallback"
        else:
            try:
                eigenvalues, eigenvectors = eigsh(D_comb - W_comb, k=K, sigma=0, which="LM", **kw)
                if stats is not None:
                    stats["route"] = "shift-invert"
            except Exception:
                eigenvalues, eigenvectors = eigsh(D_comb - W_comb, k=K, which="SM", **kw)
                if stats is not None:
                    stats["route"] = "SM-fallback"
        eigenvalues, eigenvectors = torch.from_numpy(eigenvalues), torch.from_numpy(eigenvectors.T).float()
    else:
        raise ValueError(which_matrix)
    eigenvectors = sign_rule_(eigenvectors.clone() if torch.is_tensor(eigenvectors) else eigenvectors)
    return eigenvalues, eigenvectors


def eigh_f64(feats: torch.Tensor, K: int, normalize=True, lapnorm=True, threshold_at_zero=True, W_color=None,
             image_color_lambda=0.0):
    """float64 dense ground truth: scipy.linalg.eigh(D-W, D) for the K smallest pairs (D-orthonormal vectors).

    The affinity itself is formed in float32 exactly like the reference (so the *matrix* is the same), only the
    eigensolve is done in float64; used to judge both the oracle's and the CUDA solver's accuracy.
    """
    import scipy.linalg
    W, D = affinity_matrices(feats, normalize, threshold_at_zero, None, 0.0)
    if W_color is not None and image_color_lambda > 0:
        W = W + W_color.astype(np.float32) * np.float32(image_color_lambda)
        D = np.array(get_diagonal(W).todense())
    W64 = W.astype(np.float64)
    d64 = np.diag(D).astype(np.float64)
    L = np.diag(d64) - W64
    if lapnorm:
        vals, vecs = scipy.linalg.eigh(L, np.diag(d64), subset_by_index=[0, K - 1])
    else:
        vals, vecs = scipy.linalg.eigh(L, subset_by_index=[0, K - 1])
    return vals, vecs.T


def eigs_from_affinity(W_feat: np.ndarray, K: int, lapnorm: bool = True, rng_seed=None, stats=None):
    """extract.py:216-240 from the point where the feature affinity has come back from the GPU (``W_feat.cpu().numpy()``,
    :195): combine (no colour term), degree, dense diag, eigsh with the SM fallback, sign rule. This is the CPU half
    of the reference AS SHIPPED (GPU matmul + CPU eigsh); bench.py times it for the `as_shipped` baseline."""
    import time
    kw = {} if rng_seed is None else {"rng": np.random.default_rng(rng_seed)}
    t0 = time.perf_counter()
    W_comb = W_feat + 0 * 0.0
    D_comb = np.array(get_diagonal(W_comb).todense())
    t1 = time.perf_counter()
    try:
        eigenvalues, eigenvectors = eigsh(D_comb - W_comb, k=K, sigma=0, which="LM", M=D_comb if lapnorm else None, **kw)
        route = "shift-invert"
    except Exception:
        eigenvalues, eigenvectors = eigsh(D_comb - W_comb, k=K, which="SM", M=D_comb if lapnorm else None, **kw)
        route = "SM-fallback"
    t2 = time.perf_counter()
    if stats is not None:
        stats.update(route=route, degree_s=t1 - t0, eigsh_s=t2 - t1)
    eigenvectors = sign_rule_(torch.from_numpy(eigenvectors.T).float())
    return torch.from_numpy(eigenvalues), eigenvectors
