"""Generates tests/golden/*.npz by running the reference's OWN ``_extract_eig`` (imported from /root/reference through
oracle/ref_shim.py) on small seeded inputs. Run in the dev container only:

    python -m oracle.make_golden

Each fixture stores the inputs (features, optional JPEG bytes + the low-res image the reference derived from it)
and the tensors the reference saved. tests/test_cpu_oracle.py checks the oracle restatement against them;
tests/test_golden_gpu.py checks the CUDA path against them."""
from __future__ import annotations

import importlib
import io
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import ref_shim  # noqa: E402

synth = importlib.import_module("deep-spectral-segmentation_b200.synth")
GOLDEN = ROOT / "tests" / "golden"

CASES = [
    # name, N-grid (Hp, Wp), d, K, rank, seed, kwargs
    ("lap_196_k5", (14, 14), 64, 5, 6, 0, {}),
    ("lap_143_k5_odd", (11, 13), 64, 5, 4, 1, {}),
    ("lap_196_k8", (14, 14), 96, 8, 6, 2, {}),
    ("lap_100_nolapnorm", (10, 10), 64, 4, 4, 3, {"lapnorm": False}),
    ("lap_120_nonorm_nothr", (10, 12), 64, 5, 6, 4, {"normalize": False, "threshold_at_zero": False}),
    ("lap_196_color10", (14, 14), 64, 5, 6, 5, {"image_color_lambda": 10.0}),
    ("lap_150_color1", (10, 15), 64, 6, 6, 6, {"image_color_lambda": 1.0}),
]


def main():
    assert ref_shim.available(), "reference sources not present"
    GOLDEN.mkdir(parents=True, exist_ok=True)
    from PIL import Image
    for name, (Hp, Wp), d, K, rank, seed, kw in CASES:
        N = Hp * Wp
        P = 16
        H, W = Hp * P + 3, Wp * P + 5  # un-cropped size: exercises the crop arithmetic of get_image_sizes
        feats = synth.structured_features(N, d, rank, seed)
        if kw.get("normalize", True) is False:
            feats = feats - feats.min() + 0.1
        fd = {"k": feats[None].clone(), "indices": torch.tensor(0), "file": f"{name}.jpg", "id": name,
              "model_name": "dino_vits16", "patch_size": P, "shape": (1, 3, H, W)}
        jpeg = np.zeros(0, np.uint8)
        image_lr = np.zeros(0)
        with tempfile.TemporaryDirectory() as td:
            if kw.get("image_color_lambda", 0) > 0:
                img = synth.blobs_image(H, W, 100 + seed).numpy()
                buf = io.BytesIO()
                Image.fromarray(img).save(buf, format="JPEG", quality=95)
                jpeg = np.frombuffer(buf.getvalue(), np.uint8)
                (Path(td) / f"{name}.jpg").write_bytes(buf.getvalue())
                image_lr = np.array(Image.open(str(Path(td) / f"{name}.jpg")).resize((Wp, Hp), Image.BILINEAR)) / 255.0
            out = ref_shim.run_reference_extract_eig(fd, td, K=K, images_root=td, **kw)
        np.savez_compressed(GOLDEN / f"{name}.npz", feats=feats.numpy(), K=K, patch=P, shape=np.array([1, 3, H, W]),
                            kwargs=np.array(repr(kw)), jpeg=jpeg, image_lr=image_lr,
                            eigenvalues=np.asarray(out["eigenvalues"], dtype=np.float32),
                            eigenvectors=out["eigenvectors"].numpy())
        print(name, "lambda:", np.asarray(out["eigenvalues"]))


if __name__ == "__main__":
    main()
