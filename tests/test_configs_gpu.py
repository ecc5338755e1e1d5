"""-m gpu: the larger BASELINE.json configurations at full size, checked through size-independent properties and, on
a subset, against the oracle: C3 (dino_vitb8, 480x480 -> 3600 patches, colour-KNN affinity, K=15) and C2's exact shape."""
import numpy as np
import pytest
import torch

from conftest import load_pkg
from test_cpu_oracle import _aligned_err

pytestmark = pytest.mark.gpu


def _props(evals, evecs, Wm, N, lapnorm=True):
    """lambda ascending, |lambda_0| small, V^T D V = I, v0 constant, residual ||(D-W)v - lambda D v|| small."""
    W = Wm[:, :N].double()
    d = W.sum(1)
    V = evecs.double()
    G = (V * d[None]) @ V.T
    assert (G - torch.eye(V.shape[0], dtype=torch.float64, device=V.device)).abs().max().item() <= 2e-5
    assert (evals[1:] - evals[:-1]).min().item() >= -1e-6 and abs(evals[0].item()) <= 1e-6
    R = (d[None] * V - V @ W) - evals.double()[:, None] * (d[None] * V)      # rows: (D - W) v - lambda D v
    rel = (R.norm(dim=1) / (d[None] * V).norm(dim=1)).max().item()
    assert rel <= 2e-5, rel
    assert (V[0].max() - V[0].min()).item() <= 1e-7 * V[0].abs().max().item() + 1e-12
    return rel


def test_config3_vitb8_480_colour_knn_K15(cuda):
    from oracle import eigs_ref
    vit = load_pkg("vit")
    spectral = load_pkg("spectral")
    synth = load_pkg("synth")
    model, _, P, _ = vit.get_model("dino_vitb8", seed=0, device=cuda)
    imgs = synth.blobs_batch(2, 480, 480, seed0=7)
    k = model.forward_k(imgs.to(cuda))
    assert tuple(k.shape) == (2, 3600, 768) and torch.isfinite(k).all()
    # low-resolution colour image exactly as extract.py:199-204 derives it (PIL bilinear of the whole image)
    from PIL import Image
    lr = np.stack([np.array(Image.fromarray(im.numpy()).resize((60, 60), Image.BILINEAR)) / 255.0 for im in imgs])
    rgb = torch.from_numpy(lr.reshape(2, 3600, 3).astype(np.float32)).to(cuda)
    cc = spectral.knn_color_counts(rgb, 60, 60)
    Wm = spectral.affinity(k, True, True, cc, 10.0)
    ev, vec, info, resid = spectral.eigsh_laplacian(Wm, 3600, 15)
    torch.cuda.synchronize()
    assert int(info[:, 1].min()) == 1, info
    for b in range(2):
        rel = _props(ev[b], vec[b], Wm[b], 3600)
        print(f"C3 image {b}: lanczos steps {int(info[b, 0])}, max relative residual {rel:.2e}")
    # one image against the oracle (the reference's scipy route: ~10 s on the CPU at N = 3600)
    ev_o, vec_o = eigs_ref.extract_eig(k[0].cpu(), 15, image_lr=lr[0], image_color_lambda=10.0, rng_seed=0)
    assert np.abs(ev[0].cpu().numpy() - ev_o.numpy()).max() <= 2e-5
    err = _aligned_err(vec[0].cpu().numpy(), vec_o.numpy())
    ev64 = ev[0].double().cpu().numpy()
    gaps = np.minimum(np.diff(ev64, prepend=-1.0), np.diff(ev64, append=2.0))
    print("C3 rel-L2 vs oracle:", err, "gaps", gaps)
    assert np.all(err <= np.maximum(1e-4, 2e-6 / np.maximum(gaps, 1e-9)))
    assert np.array_equal(cc[0].cpu().numpy().astype(np.float64), eigs_ref.knn_affinity(lr[0]).toarray())


def test_config2_shape_batch_properties(cuda):
    pipeline = load_pkg("pipeline")
    synth = load_pkg("synth")
    pipe = pipeline.SpectralPipeline("dino_vits16", K=5, device=cuda, vit_batch=16)
    imgs = synth.blobs_batch(24, 480, 480, seed0=100)
    ev, vec, info = (t.clone() for t in pipe.run_host(imgs.pin_memory()))   # run_host returns reusable pinned buffers
    assert int(info[:, 1].min()) == 1 and tuple(vec.shape) == (24, 5, 900)
    Wm = pipe._bufs["W"]
    for b in (0, 7, 23):
        _props(ev[b].to(cuda), vec[b].to(cuda), Wm[b], 900)
    # linearity / idempotence style checks: permuting the batch permutes the outputs bitwise
    perm = torch.randperm(24, generator=torch.Generator().manual_seed(0))
    ev2, vec2, _ = pipe.run_host(imgs[perm].contiguous().pin_memory())
    assert torch.equal(vec2, vec[perm]) and torch.equal(ev2, ev[perm])


def test_config5_vitb8_640_K32_properties(cuda):
    """BASELINE config 5 shape: dino_vitb8 at 640x640 (6400 patches, T = 6401), K = 32."""
    vit = load_pkg("vit")
    spectral = load_pkg("spectral")
    synth = load_pkg("synth")
    model, _, P, _ = vit.get_model("dino_vitb8", seed=0, device=cuda)
    imgs = synth.blobs_batch(1, 640, 640, seed0=3)
    k = model.forward_k(imgs.to(cuda))
    assert tuple(k.shape) == (1, 6400, 768) and torch.isfinite(k).all()
    Wm = spectral.affinity(k)
    ev, vec, info, resid = spectral.eigsh_laplacian(Wm, 6400, 32)
    torch.cuda.synchronize()
    print(f"C5: lanczos steps {int(info[0, 0])}, converged {int(info[0, 1])}, eigenvalues {ev[0, :6].tolist()}")
    assert int(info[0, 1]) == 1
    _props(ev[0], vec[0], Wm[0], 6400)
