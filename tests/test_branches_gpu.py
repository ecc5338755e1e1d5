"""-m gpu: the non-default branches of _extract_eig (which_matrix='affinity' / 'affinity_svd', extract.py:160-172, and
image_downsample_factor != patch size, extract.py:179-188) against the oracle, plus the file-level drop-in path."""
import numpy as np
import pytest
import torch

from conftest import load_pkg
from test_cpu_oracle import _aligned_err

pytestmark = pytest.mark.gpu


def test_which_matrix_affinity(cuda):
    from oracle import eigs_ref
    spectral = load_pkg("spectral")
    synth = load_pkg("synth")
    for N, K, seed in [(196, 5, 0), (400, 8, 1)]:
        feats = synth.structured_features(N, 384, 6, seed)
        ev_o, vec_o = eigs_ref.extract_eig(feats, K, which_matrix="affinity", rng_seed=0)   # ev ascending (numpy), vec descending
        ev, vec, info = spectral.affinity_eigs(feats[None].to(cuda), K, "affinity")
        torch.cuda.synchronize()
        assert int(info[0, 1]) == 1 and int(info[0, 2]) == 0
        ev, vec = ev[0].cpu().numpy(), vec[0].cpu().numpy()
        assert np.abs(ev[::-1] - np.asarray(ev_o)).max() <= 1e-4 * np.abs(ev_o).max()
        err = _aligned_err(vec, vec_o.numpy())
        print(f"affinity N={N} K={K}: rel-L2 {err}")
        assert err.max() <= 1e-4
        assert np.abs(np.linalg.norm(vec, axis=1) - 1).max() <= 1e-5        # unit 2-norm, not D-normalised
        for k in range(K):
            assert not (0.5 < float((vec[k] > 0).mean()) < 1.0)


def test_which_matrix_affinity_svd(cuda):
    from oracle import eigs_ref
    spectral = load_pkg("spectral")
    synth = load_pkg("synth")
    feats = synth.structured_features(300, 384, 6, 2)
    s_o, u_o = eigs_ref.extract_eig(feats, 6, which_matrix="affinity_svd")
    s, u, info = spectral.affinity_eigs(feats[None].to(cuda), 6, "affinity_svd")
    torch.cuda.synchronize()
    assert int(info[0, 1]) == 1
    assert np.abs(s[0].cpu().numpy() - s_o.numpy()).max() <= 1e-4 * float(s_o.max())
    err = _aligned_err(u[0].cpu().numpy(), u_o.numpy())
    print("affinity_svd rel-L2", err)
    assert err.max() <= 1e-4


@pytest.mark.parametrize("Hp,Wp,Hl,Wl", [(14, 14, 28, 28), (10, 15, 20, 30), (7, 9, 21, 18)])
def test_bilinear_upsample_matches_torch(cuda, Hp, Wp, Hl, Wl):
    from oracle import eigs_ref
    spectral = load_pkg("spectral")
    g = torch.Generator().manual_seed(Hp)
    f = torch.randn(2, Hp * Wp, 64, generator=g)
    got = spectral.upsample_bilinear(f.to(cuda), Hp, Wp, Hl, Wl).cpu()
    for b in range(2):
        want = eigs_ref.upsample_features(f[b], (Hp, Wp), (Hl, Wl))
        assert (got[b] - want).abs().max().item() <= 1e-5
    n = spectral.normalize_rows(f.to(cuda)).cpu()
    assert (n - torch.nn.functional.normalize(f, p=2, dim=-1)).abs().max().item() <= 1e-6


def _write_dataset(tmp_path, n, H, W, fmt="png"):
    from PIL import Image
    synth = load_pkg("synth")
    root = tmp_path / "images"
    root.mkdir()
    names = []
    for i in range(n):
        name = f"img_{i:03d}.{fmt}"
        Image.fromarray(synth.blobs_image(H, W, 50 + i).numpy()).save(root / name, quality=95)
        names.append(name)
    (tmp_path / "list.txt").write_text("\n".join(names + names[:1]) + "\n")  # with a duplicate line
    return root, tmp_path / "list.txt", names


def test_drop_in_files_features_then_eigs(cuda, tmp_path):
    """extract_features -> extract_eigs through files, checked against the oracle end to end (same weights)."""
    from oracle import dino_vit, eigs_ref
    ex = load_pkg("extract")
    utils = load_pkg("extract_utils")
    root, lst, names = _write_dataset(tmp_path, 3, 100, 132)
    fdir, edir = tmp_path / "features", tmp_path / "eigs"
    ex.extract_features(str(lst), str(root), "dino_vits16", 2, str(fdir), seed=0)
    files = sorted(fdir.iterdir())
    assert [f.name for f in files] == [n[:-4] + ".pth" for n in names]
    vit = load_pkg("vit")
    ref = dino_vit.DinoViT(dino_vit.cfg_for("dino_vits16")).eval()
    ref.load_state_dict(vit.random_state_dict("dino_vits16", 0))
    for i, f in enumerate(files):
        d = torch.load(f, map_location="cpu")
        assert d["file"] == names[i] and d["id"] == names[i][:-4] and d["shape"] == (1, 3, 100, 132)
        assert d["patch_size"] == 16 and d["model_name"] == "dino_vits16" and int(d["indices"]) == i
        assert d["k"].shape == (1, 6 * 8, 384) and d["k"].dtype == torch.float32
        img = torch.from_numpy(utils.read_image_rgb(root / names[i]))
        k_ref = ref.forward_k(dino_vit.preprocess_u8(img, 16))
        assert ((d["k"] - k_ref).norm() / k_ref.norm()).item() <= 3e-3
    ex.extract_eigs(str(root), str(fdir), str(edir), K=4)
    # skip-if-exists: a second run must not touch the files
    before = {p.name: p.stat().st_mtime_ns for p in edir.iterdir()}
    ex.extract_eigs(str(root), str(fdir), str(edir), K=4)
    assert before == {p.name: p.stat().st_mtime_ns for p in edir.iterdir()}
    for i, f in enumerate(files):
        d = torch.load(f, map_location="cpu")
        out = torch.load(edir / f.name, map_location="cpu")
        assert out["eigenvectors"].shape == (4, 48) and out["eigenvalues"].shape == (4,)
        ev_o, vec_o = eigs_ref.extract_eig(d["k"], 4, rng_seed=0)
        assert np.abs(out["eigenvalues"].numpy() - ev_o.numpy()).max() <= 2e-5
        assert _aligned_err(out["eigenvectors"].numpy(), vec_o.numpy()).max() <= 1e-4


def test_drop_in_single_worker_colour_and_downsample(cuda, tmp_path):
    """_extract_eig with the reference's own signature: colour-KNN affinity read from a JPEG + up-sampled features."""
    from oracle import eigs_ref
    from PIL import Image
    ex = load_pkg("extract")
    synth = load_pkg("synth")
    root, lst, names = _write_dataset(tmp_path, 1, 112, 96, fmt="jpg")
    feats = synth.structured_features(7 * 6, 64, 4, 3)
    fd = ex._feature_dict(feats[None].clone(), 0, names[0], "dino_vits16", 16, 112, 96)
    (tmp_path / "features").mkdir()
    ffile = tmp_path / "features" / "img_000.pth"
    torch.save(fd, ffile)
    # (a) colour affinity at the patch resolution
    ex._extract_eig((0, str(ffile)), K=4, images_root=str(root), output_dir=str(tmp_path / "e1"), image_color_lambda=10)
    lr = np.array(Image.open(root / names[0]).resize((6, 7), Image.BILINEAR)) / 255.0
    ev_o, vec_o = eigs_ref.extract_eig(feats, 4, image_lr=lr, image_color_lambda=10.0, rng_seed=0)
    out = torch.load(tmp_path / "e1" / "img_000.pth", map_location="cpu")
    assert np.abs(out["eigenvalues"].numpy() - ev_o.numpy()).max() <= 2e-5
    assert _aligned_err(out["eigenvectors"].numpy(), vec_o.numpy()).max() <= 1e-4
    # (b) image_downsample_factor=8: features bilinearly up-sampled 2x, colour affinity at that resolution
    ex._extract_eig((0, str(ffile)), K=4, images_root=str(root), output_dir=str(tmp_path / "e2"),
                    image_downsample_factor=8, image_color_lambda=1.0)
    lr = np.array(Image.open(root / names[0]).resize((12, 14), Image.BILINEAR)) / 255.0
    ev_o, vec_o = eigs_ref.extract_eig(feats, 4, image_lr=lr, image_color_lambda=1.0, rng_seed=0, grid=(7, 6), lr_size=(14, 12))
    out = torch.load(tmp_path / "e2" / "img_000.pth", map_location="cpu")
    assert out["eigenvectors"].shape == (4, 14 * 12)
    assert np.abs(out["eigenvalues"].numpy() - ev_o.numpy()).max() <= 2e-5
    assert _aligned_err(out["eigenvectors"].numpy(), vec_o.numpy()).max() <= 2e-4
    # (c) which_matrix='affinity' keeps the reference's quirk: ascending numpy eigenvalues, descending vectors
    ex._extract_eig((0, str(ffile)), K=3, images_root=str(root), output_dir=str(tmp_path / "e3"), which_matrix="affinity",
                    image_color_lambda=0)
    out = torch.load(tmp_path / "e3" / "img_000.pth", map_location="cpu", weights_only=False)
    assert isinstance(out["eigenvalues"], np.ndarray) and np.all(np.diff(out["eigenvalues"]) >= 0)
    ev_o, vec_o = eigs_ref.extract_eig(feats, 3, which_matrix="affinity", rng_seed=0)
    assert _aligned_err(out["eigenvectors"].numpy(), vec_o.numpy()).max() <= 1e-4


def test_extract_all_fused_matches_two_stage(cuda, tmp_path):
    ex = load_pkg("extract")
    root, lst, names = _write_dataset(tmp_path, 3, 96, 96)
    ex.extract_features(str(lst), str(root), "dino_vits16", 4, str(tmp_path / "f"), seed=0)
    ex.extract_eigs(str(root), str(tmp_path / "f"), str(tmp_path / "e"), K=3)
    ex.extract_all(str(lst), str(root), "dino_vits16", str(tmp_path / "f2"), str(tmp_path / "e2"), K=3, batch_size=4, seed=0)
    for n in names:
        a = torch.load(tmp_path / "e" / (n[:-4] + ".pth"))
        b = torch.load(tmp_path / "e2" / (n[:-4] + ".pth"))
        assert torch.equal(a["eigenvectors"], b["eigenvectors"]) and torch.equal(a["eigenvalues"], b["eigenvalues"])
        fa = torch.load(tmp_path / "f" / (n[:-4] + ".pth"))
        fb = torch.load(tmp_path / "f2" / (n[:-4] + ".pth"))
        assert torch.equal(fa["k"], fb["k"]) and fa["shape"] == fb["shape"]
