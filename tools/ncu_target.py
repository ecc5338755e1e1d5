"""Short single-GPU run of the pipeline for ncu captures: 2 warm-up passes + 1 measured pass of B images."""
import importlib
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
PKG = "deep-spectral-segmentation_b200"
pipeline = importlib.import_module(PKG + ".pipeline")
synth = importlib.import_module(PKG + ".synth")
torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
vb = int(sys.argv[3]) if len(sys.argv) > 3 else 32
pipe = pipeline.SpectralPipeline("dino_vits16", K=5, device=dev, vit_batch=vb)
imgs = synth.blobs_batch(8, 480, 480, 0).repeat((B + 7) // 8, 1, 1, 1)[:B].contiguous().to(dev)
for _ in range(passes):
    pipe.run_device(imgs)
torch.cuda.synchronize()
print("done")
