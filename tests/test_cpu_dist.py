"""CPU suite: the N>1 host logic under world_size-2 gloo -- the single weight broadcast and the strided image shard."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, outdir):
    import importlib
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pl = importlib.import_module("deep-spectral-segmentation_b200.pipeline")
    vit = importlib.import_module("deep-spectral-segmentation_b200.vit")
    name = "dino_vits16"
    # rank 0 owns the real weights (seed 7); the others must end up with exactly those after ONE broadcast
    sd0 = vit.random_state_dict(name, 7) if rank == 0 else None
    sd = pl.broadcast_weights(name, seed=0, device="cpu", src=0, state_dict=sd0)
    flat = pl.flatten_state_dict(sd, name)
    shard = pl.shard_indices(11, rank, world)
    torch.save({"sum": flat.double().sum().item(), "abs": flat.abs().double().sum().item(), "n": flat.numel(),
                "shard": shard}, os.path.join(outdir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_weight_broadcast_and_sharding_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"r{i}.pt") for i in range(world)]
    assert r[0]["n"] == r[1]["n"] > 21_000_000                      # ViT-S/16: ~21.6 M parameters in one buffer
    assert r[0]["sum"] == r[1]["sum"] and r[0]["abs"] == r[1]["abs"]  # bitwise the same weights on both ranks
    import importlib
    vit = importlib.import_module("deep-spectral-segmentation_b200.vit")
    pl = importlib.import_module("deep-spectral-segmentation_b200.pipeline")
    want = pl.flatten_state_dict(vit.random_state_dict("dino_vits16", 7), "dino_vits16")
    assert r[1]["sum"] == want.double().sum().item()                # ... and they are rank 0's (seed 7), not seed 0
    assert sorted(r[0]["shard"] + r[1]["shard"]) == list(range(11)) and r[0]["shard"] == [0, 2, 4, 6, 8, 10]


def test_flatten_unflatten_roundtrip():
    import importlib
    vit = importlib.import_module("deep-spectral-segmentation_b200.vit")
    pl = importlib.import_module("deep-spectral-segmentation_b200.pipeline")
    sd = vit.random_state_dict("dino_vits16", 1)
    back = pl.unflatten_state_dict(pl.flatten_state_dict(sd, "dino_vits16"), sd, "dino_vits16")
    assert all(torch.equal(back[k], sd[k]) for k in back)
