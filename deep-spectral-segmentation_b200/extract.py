"""Drop-in replacements for the reference's hot-path callables (extract/extract.py):

    extract_features  (:21-116)   images -> features/{id}.pth      {'k','indices','file','id','model_name','patch_size','shape'}
    _extract_eig      (:119-244)  one features file -> eigs/{image_id}.pth  {'eigenvalues','eigenvectors'}
    extract_eigs      (:247-280)  directory of features files -> eigs files (batched on the GPU)
    extract_all                   fused: images -> both file layouts without the disk round-trip (new)

Same argument names, defaults, file layouts and skip-if-exists behaviour; all arithmetic runs in libdss_b200."""
from __future__ import annotations

from collections import defaultdict
from pathlib import Path
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib, spectral
from . import extract_utils as utils

torch.set_grad_enabled(False)  # extract.py:838


def _device():
    if not torch.cuda.is_available():
        raise _lib.DssError("a CUDA device is required: the hot path has no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device())


def _feature_dict(k: torch.Tensor, index: int, file: str, model_name: str, patch_size: int, H: int, W: int) -> dict:
    """The dict the reference saves at extract.py:98-110 (k is (1, N, d) fp32 on the CPU; shape is un-cropped)."""
    return {"k": k, "indices": torch.tensor(index), "file": file, "id": Path(file).stem, "model_name": model_name,
            "patch_size": patch_size, "shape": (1, 3, H, W)}


def extract_features(images_list: str, images_root: Optional[str], model_name: str, batch_size: int, output_dir: str,
                     which_block: int = -1, checkpoint: Optional[str] = None, seed: Optional[int] = None,
                     random_init: bool = False, yes: Optional[bool] = None):
    """
    Extract features from a list of images.

    Example:
        python extract.py extract_features \
            --images_list "./data/VOC2012/lists/images.txt" \
            --images_root "./data/VOC2012/images" \
            --output_dir "./data/VOC2012/features/dino_vits16" \
            --model_name dino_vits16 \
            --batch_size 1
    """
    utils.make_output_dir(output_dir, assume_yes=yes)
    model_name = model_name.lower()
    if not ("dino" in model_name or "mocov3" in model_name):
        raise ValueError(model_name)
    dev = _device()
    model, _, patch_size, _ = utils.get_model(model_name, checkpoint=checkpoint, seed=seed, device=dev,
                                              random_init=random_init)
    filenames = Path(images_list).read_text().splitlines()
    dataset = utils.ImagesDataset(filenames=filenames, images_root=images_root)
    print(f"Dataset size: {len(dataset)=}")

    pending: dict = defaultdict(list)  # (H, W) -> [(image, file, index)]

    def flush(shape_key):
        items = pending.pop(shape_key, [])
        if not items:
            return
        H, W = shape_key
        batch = torch.stack([it[0] for it in items]).pin_memory().to(dev, non_blocking=True)
        k = model.forward_k(batch, which_block=which_block).cpu()
        for j, (_, file, index) in enumerate(items):
            out = _feature_dict(k[j:j + 1].clone(), index, file, model_name, patch_size, H, W)
            torch.save(out, str(Path(output_dir) / f"{out['id']}.pth"))

    for i in range(len(dataset)):
        file = dataset.filenames[i]
        output_file = Path(output_dir) / f"{Path(file).stem}.pth"
        if output_file.is_file():
            print(f"Skipping existing file {str(output_file)}")
            continue
        image, file, index = dataset[i]
        key = (int(image.shape[0]), int(image.shape[1]))
        pending[key].append((image, file, index))
        if len(pending[key]) >= max(1, int(batch_size)):
            flush(key)
    for key in list(pending.keys()):
        flush(key)
    print(f"Saved features to {output_dir}")


def _load_image_lr(images_root: str, image_id: str, W_lr: int, H_lr: int) -> np.ndarray:
    """extract.py:199-204: PIL open, BILINEAR resize of the whole image to (W_lr, H_lr), /255 -> (H_lr, W_lr, 3)."""
    from PIL import Image
    image_file = str(Path(images_root) / f"{image_id}.jpg")
    image_lr = Image.open(image_file).resize((W_lr, H_lr), Image.BILINEAR)
    return np.array(image_lr) / 255.0


def _check_supported(which_matrix, which_color_matrix, image_color_lambda):
    if which_matrix == "affinity_torch":
        raise RuntimeError("which_matrix='affinity_torch' calls torch.eig, which PyTorch removed (dead in the reference)")
    if which_matrix not in ("laplacian", "matting_laplacian", "affinity", "affinity_svd"):
        raise ValueError(f"unknown which_matrix={which_matrix!r}")
    if which_matrix in ("laplacian", "matting_laplacian") and image_color_lambda > 0 and which_color_matrix != "knn":
        # 'rw' needs pymatting's _rw_laplacian, a third-party routine that is neither installed nor restatable offline
        raise NotImplementedError(f"which_color_matrix={which_color_matrix!r}: only 'knn' is built")


def _eigs_for_group(data_dicts: List[dict], K: int, images_root, which_features, normalize, lapnorm, threshold_at_zero,
                    image_downsample_factor, image_color_lambda, dev, which_matrix="laplacian"
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """One GPU batch: feature dicts whose patch grids have the same size -> (eigenvalues [B,K], eigenvectors [B,K,N]) CPU."""
    feats = torch.stack([d[which_features].squeeze() for d in data_dicts]).to(torch.float32)
    feats = feats.pin_memory().to(dev, non_blocking=True)
    if which_matrix in ("affinity", "affinity_svd"):
        evals, evecs, info = spectral.affinity_eigs(feats, K, which_matrix, normalize, threshold_at_zero)
    else:
        rgb_lr, lr_size = None, None
        sizes = utils.get_image_sizes(data_dicts[0])
        P, H_patch, W_patch, H_pad, W_pad = sizes[4:]
        factor = P if image_downsample_factor is None else image_downsample_factor
        H_lr, W_lr = H_pad // factor, W_pad // factor
        if (H_patch, W_patch) != (H_lr, W_lr):
            # extract.py:148,179-188: normalise first, then bilinear up-sampling (no re-normalisation afterwards)
            if normalize:
                feats = spectral.normalize_rows(feats)
                normalize = False
            feats = spectral.upsample_bilinear(feats, H_patch, W_patch, H_lr, W_lr)
        if image_color_lambda > 0:
            lr = [_load_image_lr(images_root, d["file"][:-4], W_lr, H_lr) for d in data_dicts]
            rgb_lr = torch.from_numpy(np.stack(lr).reshape(len(lr), H_lr * W_lr, 3).astype(np.float32)).to(dev)
            lr_size = (H_lr, W_lr)
        evals, evecs, info, _ = spectral.laplacian_eigs(feats, K, normalize, threshold_at_zero, lapnorm, rgb_lr,
                                                        lr_size, image_color_lambda)
    evals, evecs, info = evals.cpu(), evecs.cpu(), info.cpu()
    bad = (info[:, 1] == 0).nonzero().flatten().tolist()
    if bad:
        print(f"Warning: eigensolver did not reach its tolerance for {[data_dicts[i]['id'] for i in bad]}")
    if which_matrix == "affinity" and bool((info[:, 2] != 0).any()):
        print("Warning: a negative eigenvalue exceeds the K-th largest in magnitude; eigsh(which='LM') would pick it")
    return evals, evecs


def _save_eigs(output_file, which_matrix, evals_k: torch.Tensor, evecs_k: torch.Tensor):
    """extract.py:242-244. The 'affinity' branch of the reference keeps `eigenvalues` as the ascending numpy array
    eigsh returned while the eigenvectors are flipped to descending order (extract.py:171-172); mirrored here."""
    if which_matrix == "affinity":
        eigenvalues = evals_k.flip(0).numpy().copy()
    else:
        eigenvalues = evals_k.clone()
    Path(output_file).parent.mkdir(parents=True, exist_ok=True)
    torch.save({"eigenvalues": eigenvalues, "eigenvectors": evecs_k.clone()}, str(output_file))


def _extract_eig(inp: Tuple[int, str], K: int, images_root: str, output_dir: str, which_matrix: str = "laplacian",
                 which_features: str = "k", normalize: bool = True, lapnorm: bool = True,
                 which_color_matrix: str = "knn", threshold_at_zero: bool = True,
                 image_downsample_factor: Optional[int] = None, image_color_lambda: float = 10):
    """Single-image worker with the reference's signature (extract.py:119-132); writes {output_dir}/{image_id}.pth."""
    index, features_file = inp
    data_dict = torch.load(features_file, map_location="cpu")
    image_id = data_dict["file"][:-4]
    output_file = str(Path(output_dir) / f"{image_id}.pth")
    if Path(output_file).is_file():
        print(f"Skipping existing file {str(output_file)}")
        return
    _check_supported(which_matrix, which_color_matrix, image_color_lambda)
    evals, evecs = _eigs_for_group([data_dict], K, images_root, which_features, normalize, lapnorm, threshold_at_zero,
                                   image_downsample_factor, image_color_lambda, _device(), which_matrix)
    _save_eigs(output_file, which_matrix, evals[0], evecs[0])


def extract_eigs(images_root: str, features_dir: str, output_dir: str, which_matrix: str = "laplacian",
                 which_color_matrix: str = "knn", which_features: str = "k", normalize: bool = True,
                 threshold_at_zero: bool = True, lapnorm: bool = True, K: int = 20,
                 image_downsample_factor: Optional[int] = None, image_color_lambda: float = 0.0,
                 multiprocessing: int = 0, batch_size: int = 128, yes: Optional[bool] = None):
    """
    Extracts eigenvalues from features.

    Example:
        python extract.py extract_eigs \
            --images_root "./data/VOC2012/images" \
            --features_dir "./data/VOC2012/features/dino_vits16" \
            --which_matrix "laplacian" \
            --output_dir "./data/VOC2012/eigs/laplacian" \
            --K 5
    """
    utils.make_output_dir(output_dir, assume_yes=yes)
    kwargs = dict(K=K, which_matrix=which_matrix, which_features=which_features, which_color_matrix=which_color_matrix,
                  normalize=normalize, threshold_at_zero=threshold_at_zero, images_root=images_root,
                  output_dir=output_dir, image_downsample_factor=image_downsample_factor,
                  image_color_lambda=image_color_lambda, lapnorm=lapnorm)
    print(kwargs)
    _check_supported(which_matrix, which_color_matrix, image_color_lambda)
    dev = _device()
    inputs = list(enumerate(sorted(Path(features_dir).iterdir())))
    import time
    start = time.time()
    groups: dict = defaultdict(list)  # N -> [data_dict]

    def flush(key):
        dds = groups.pop(key, [])
        if not dds:
            return
        evals, evecs = _eigs_for_group(dds, K, images_root, which_features, normalize, lapnorm, threshold_at_zero,
                                       image_downsample_factor, image_color_lambda, dev, which_matrix)
        for j, d in enumerate(dds):
            _save_eigs(Path(output_dir) / f"{d['file'][:-4]}.pth", which_matrix, evals[j], evecs[j])

    for index, features_file in inputs:
        data_dict = torch.load(str(features_file), map_location="cpu")
        image_id = data_dict["file"][:-4]
        if (Path(output_dir) / f"{image_id}.pth").is_file():
            print(f"Skipping existing file {str(Path(output_dir) / (image_id + '.pth'))}")
            continue
        key = (tuple(data_dict[which_features].shape[-2:]), tuple(data_dict["shape"]))
        groups[key].append(data_dict)
        if len(groups[key]) >= batch_size:
            flush(key)
    for key in list(groups.keys()):
        flush(key)
    print(f"Finished in {time.time() - start:.1f}s")


def _extract_single_region_segmentations(inp, threshold: float, output_dir: str):
    """Worker with the reference's signature (extract.py:364-390): eigenvector 1 > threshold on the patch grid -> PNG."""
    from PIL import Image
    index, (feature_path, eigs_path) = inp
    data_dict = torch.load(feature_path, map_location="cpu")
    data_dict.update(torch.load(eigs_path, map_location="cpu", weights_only=False))
    id = Path(data_dict["id"])
    output_file = str(Path(output_dir) / f"{id}.png")
    if Path(output_file).is_file():
        print(f"Skipping existing file {str(output_file)}")
        return
    B, C, H, W, P, H_patch, W_patch, H_pad, W_pad = utils.get_image_sizes(data_dict)
    eigenvector = data_dict["eigenvectors"][1].numpy()  # smallest non-zero eigenvector
    segmap = (eigenvector > threshold).reshape(H_patch, W_patch)
    Image.fromarray(segmap).convert("L").save(output_file)


def extract_single_region_segmentations(features_dir: str, eigs_dir: str, output_dir: str, threshold: float = 0.0,
                                        multiprocessing: int = 0, yes: Optional[bool] = None):
    """
    First consumer of the eigs files, same command / file contract as the reference
    (extract/extract.py:393-411): thresholds the Fiedler-like eigenvector of every image into a patch-grid mask.

    Example:
    python extract.py extract_single_region_segmentations \
        --features_dir "./data/VOC2012/features/dino_vits16" \
        --eigs_dir "./data/VOC2012/eigs/laplacian" \
        --output_dir "./data/VOC2012/single_region_segmentation/patches" \
    """
    utils.make_output_dir(output_dir, assume_yes=yes)
    inputs = utils.get_paired_input_files(features_dir, eigs_dir)
    utils.parallel_process(inputs, lambda inp: _extract_single_region_segmentations(inp, threshold, output_dir),
                           multiprocessing)


def _extract_multi_region_segmentations(inp, adaptive: bool, non_adaptive_num_segments: int, infer_bg_index: bool,
                                        kmeans_baseline: bool, output_dir: str, num_eigenvectors: int,
                                        random_state: Optional[int] = None):
    """Worker with the reference's signature (extract.py:283-349): K-means on the non-constant eigenvectors (or on the
    raw features, ``kmeans_baseline``) of one image -> label map on the patch grid, background label swapped to 0.

    The clustering is scikit-learn's KMeans exactly as in the reference (unseeded there; ``random_state`` is an
    extension, None = reference behaviour). It runs on the host: N <= a few thousand points in <= K dimensions.
    """
    from PIL import Image
    from sklearn.cluster import KMeans
    index, (feature_path, eigs_path) = inp
    data_dict = torch.load(feature_path, map_location="cpu")
    data_dict.update(torch.load(eigs_path, map_location="cpu", weights_only=False))
    id = Path(data_dict["id"])
    output_file = str(Path(output_dir) / f"{id}.png")
    if Path(output_file).is_file():
        print(f"Skipping existing file {str(output_file)}")
        return
    B, C, H, W, P, H_patch, W_patch, H_pad, W_pad = utils.get_image_sizes(data_dict)
    if adaptive:   # number of segments = position of the largest eigengap (the gap after the constant vector excluded)
        by_gap = np.argsort(np.diff(data_dict["eigenvalues"].numpy()))[::-1]
        n_clusters = by_gap[by_gap != 0][0] + 1
    else:
        n_clusters = non_adaptive_num_segments
    kmeans = KMeans(n_clusters=n_clusters) if random_state is None else KMeans(n_clusters=n_clusters,
                                                                                random_state=random_state)
    if kmeans_baseline:
        clusters = kmeans.fit_predict(data_dict["k"].squeeze().numpy())
    else:
        clusters = kmeans.fit_predict(data_dict["eigenvectors"][1:1 + num_eigenvectors].numpy().T)
    if clusters.size == H_patch * W_patch:
        segmap = clusters.reshape(H_patch, W_patch)
    elif clusters.size == H_patch * W_patch * 4:     # eigenvectors of the image_downsample_factor = P/2 mode
        segmap = clusters.reshape(H_patch * 2, W_patch * 2)
    else:
        raise ValueError(f"{clusters.size} labels do not fit a {H_patch} x {W_patch} patch grid")
    if infer_bg_index:   # the label owning most of the border becomes 0 (labels 0 and bg are swapped)
        labels, share = utils.get_border_fraction(segmap)
        bg_index = labels[np.argmax(share)].item()
        bg_region, zero_region = segmap == bg_index, segmap == 0
        segmap[bg_region] = 0
        segmap[zero_region] = bg_index
    Image.fromarray(segmap).convert("L").save(output_file)


def extract_multi_region_segmentations(features_dir: str, eigs_dir: str, output_dir: str, adaptive: bool = False,
                                       non_adaptive_num_segments: int = 4, infer_bg_index: bool = True,
                                       kmeans_baseline: bool = False, num_eigenvectors: int = 1_000_000,
                                       multiprocessing: int = 0, random_state: Optional[int] = None,
                                       yes: Optional[bool] = None):
    """
    Second consumer of the eigs files, same command / file contract as the reference
    (extract/extract.py:352-376).

    Example:
    python extract.py extract_multi_region_segmentations \
        --features_dir "./data/VOC2012/features/dino_vits16" \
        --eigs_dir "./data/VOC2012/eigs/laplacian" \
        --output_dir "./data/VOC2012/multi_region_segmentation/fixed" \
    """
    utils.make_output_dir(output_dir, assume_yes=yes)
    inputs = utils.get_paired_input_files(features_dir, eigs_dir)
    utils.parallel_process(
        inputs, lambda inp: _extract_multi_region_segmentations(inp, adaptive, non_adaptive_num_segments, infer_bg_index,
                                                                kmeans_baseline, output_dir, num_eigenvectors,
                                                                random_state), multiprocessing)


def extract_all(images_list: str, images_root: Optional[str], model_name: str, features_dir: Optional[str],
                eigs_dir: str, K: int = 20, batch_size: int = 16, which_block: int = -1, normalize: bool = True,
                threshold_at_zero: bool = True, lapnorm: bool = True, image_color_lambda: float = 0.0,
                checkpoint: Optional[str] = None, seed: int = 0, yes: Optional[bool] = None):
    """Fused extract_features + extract_eigs: features never leave the GPU between the two stages. Writes the eigs
    files (and, if features_dir is given, the features files) in the reference's layouts."""
    if features_dir:
        utils.make_output_dir(features_dir, assume_yes=yes)
    utils.make_output_dir(eigs_dir, assume_yes=yes)
    model_name = model_name.lower()
    dev = _device()
    model, _, patch_size, _ = utils.get_model(model_name, checkpoint=checkpoint, seed=seed, device=dev)
    filenames = Path(images_list).read_text().splitlines()
    dataset = utils.ImagesDataset(filenames=filenames, images_root=images_root)
    pending: dict = defaultdict(list)

    def flush(key):
        items = pending.pop(key, [])
        if not items:
            return
        H, W = key
        batch = torch.stack([it[0] for it in items]).pin_memory().to(dev, non_blocking=True)
        k = model.forward_k(batch, which_block=which_block)
        rgb_lr, lr_size = None, None
        Hp, Wp = H // patch_size, W // patch_size
        if image_color_lambda > 0:
            lr = [_load_image_lr(images_root, it[1][:-4], Wp, Hp) for it in items]
            rgb_lr = torch.from_numpy(np.stack(lr).reshape(len(lr), Hp * Wp, 3).astype(np.float32)).to(dev)
            lr_size = (Hp, Wp)
        evals, evecs, info, _ = spectral.laplacian_eigs(k, K, normalize, threshold_at_zero, lapnorm, rgb_lr, lr_size,
                                                        image_color_lambda)
        evals, evecs = evals.cpu(), evecs.cpu()
        k_cpu = k.cpu() if features_dir else None
        for j, (_, file, index) in enumerate(items):
            if features_dir:
                fd = _feature_dict(k_cpu[j:j + 1].clone(), index, file, model_name, patch_size, H, W)
                torch.save(fd, str(Path(features_dir) / f"{fd['id']}.pth"))
            out = Path(eigs_dir) / f"{file[:-4]}.pth"
            out.parent.mkdir(parents=True, exist_ok=True)
            torch.save({"eigenvalues": evals[j].clone(), "eigenvectors": evecs[j].clone()}, str(out))

    for i in range(len(dataset)):
        file = dataset.filenames[i]
        if (Path(eigs_dir) / f"{file[:-4]}.pth").is_file():
            print(f"Skipping existing file {str(Path(eigs_dir) / (file[:-4] + '.pth'))}")
            continue
        image, file, index = dataset[i]
        key = (int(image.shape[0]), int(image.shape[1]))
        pending[key].append((image, file, index))
        if len(pending[key]) >= max(1, int(batch_size)):
            flush(key)
    for key in list(pending.keys()):
        flush(key)
    print(f"Saved eigs to {eigs_dir}")
