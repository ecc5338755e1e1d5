#!/bin/bash
# the CLI under torchrun on 2 GPUs: every rank takes a strided share of the image list, 512 JPEGs -> 512 eigs files
mkdir -p gpurun_out
D=$(mktemp -d /dev/shm/dss_cli.XXXX)
python - "$D" <<'PY'
import sys, importlib
from pathlib import Path
sys.path.insert(0, ".")
import cv2
synth = importlib.import_module("deep-spectral-segmentation_b200.synth")
root = Path(sys.argv[1]) / "images"; root.mkdir()
imgs = synth.blobs_batch(64, 480, 480, seed0=0).numpy()
names = []
for i in range(512):
    name = f"im{i:05d}.jpg"
    cv2.imwrite(str(root / name), cv2.cvtColor(imgs[i % 64], cv2.COLOR_RGB2BGR), [cv2.IMWRITE_JPEG_QUALITY, 90])
    names.append(name)
(Path(sys.argv[1]) / "list.txt").write_text("\n".join(names) + "\n")
PY
CMD="extract/extract.py extract_all --images_list $D/list.txt --images_root $D/images --model_name dino_vits16 --K 5 --batch_size 128 --seed 0 --random_init True --yes True"
timeout 300 python $CMD --features_dir None --eigs_dir $D/eigs1 > gpurun_out/cli_1gpu.log 2>&1; echo "1 GPU rc $? files $(ls $D/eigs1 | wc -l)"; grep "Saved eigs" gpurun_out/cli_1gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 $CMD --features_dir None --eigs_dir $D/eigs2 > gpurun_out/cli_2gpu.log 2>&1; echo "2 GPU rc $? files $(ls $D/eigs2 | wc -l)"; grep "Saved eigs" gpurun_out/cli_2gpu.log
python - "$D" <<'PY'
import sys, torch
from pathlib import Path
d = Path(sys.argv[1]); bad = 0
for f in sorted((d / "eigs1").iterdir()):
    a, b = torch.load(f), torch.load(d / "eigs2" / f.name)
    bad += not (torch.equal(a["eigenvectors"], b["eigenvectors"]) and torch.equal(a["eigenvalues"], b["eigenvalues"]))
print("files compared", len(list((d / "eigs1").iterdir())), "different", bad)
PY
rm -rf "$D"
