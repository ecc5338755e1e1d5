"""-m gpu: the hand-written ViT forward against the fp32 PyTorch oracle (oracle/dino_vit.py) on identical weights
and images. Operands are fp16 (11-bit significand) with fp32 accumulation and an fp32 residual stream, so the
tolerance is the fp16 rounding level, stated per test."""
import pytest
import torch

from conftest import load_pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models(cuda):
    from oracle import dino_vit
    vit = load_pkg("vit")
    ref = dino_vit.build("dino_vits16", seed=0)
    mine = vit.DinoViT("dino_vits16", ref.state_dict(), device=cuda)
    return ref.to(cuda), mine


def _images(n, H, W, seed0=0):
    synth = load_pkg("synth")
    return synth.blobs_batch(n, H, W, seed0)


@pytest.mark.parametrize("Hp,Wp", [(14, 14), (30, 30), (23, 31), (31, 23), (5, 9)])
def test_pos_embed_matches_upstream_interpolation(cuda, models, Hp, Wp):
    ref, mine = models
    got = mine.pos_embed(Hp, Wp)
    want = ref.interpolate_pos_encoding(Hp * Wp, Hp * 16, Wp * 16)[0]
    err = (got - want).abs().max().item()
    assert err <= 2e-6, err


@pytest.mark.parametrize("n_blocks", [0, 1, 2, 6, 11])
def test_residual_stream_per_depth(cuda, models, n_blocks):
    from oracle import dino_vit
    ref, mine = models
    imgs = _images(2, 224, 224)
    x_ref = torch.cat([ref.forward_tokens(dino_vit.preprocess_u8(im, 16).to(cuda), n_blocks) for im in imgs])
    x = mine.forward_tokens(imgs.to(cuda), n_blocks)
    torch.cuda.synchronize()
    assert torch.isfinite(x).all()
    rel = ((x - x_ref).norm() / x_ref.norm()).item()
    mx = (x - x_ref).abs().max().item()
    print(f"tokens after {n_blocks} blocks: rel-L2 {rel:.3e} max-abs {mx:.3e} (|x|max {x_ref.abs().max().item():.3f})")
    assert rel <= 2e-3, (n_blocks, rel)   # fp16-operand rounding accumulated over <= 11 blocks


@pytest.mark.parametrize("H,W,B", [(224, 224, 1), (480, 480, 2), (375, 500, 2), (250, 333, 1)])
def test_forward_k_matches_oracle(cuda, models, H, W, B):
    from oracle import dino_vit
    ref, mine = models
    imgs = _images(B, H, W, seed0=11)
    k_ref = torch.cat([ref.forward_k(dino_vit.preprocess_u8(im, 16).to(cuda)) for im in imgs])
    k = mine.forward_k(imgs.to(cuda))
    torch.cuda.synchronize()
    assert k.shape == k_ref.shape and torch.isfinite(k).all()
    rel = ((k - k_ref).norm() / k_ref.norm()).item()
    cos = torch.nn.functional.cosine_similarity(k.flatten(0, 1), k_ref.flatten(0, 1), dim=-1).min().item()
    print(f"K features {H}x{W}: rel-L2 {rel:.3e}, min row cosine {cos:.6f}")
    assert rel <= 3e-3 and cos >= 0.9999


def test_forward_k_which_block_and_batch_invariance(cuda, models):
    from oracle import dino_vit
    ref, mine = models
    imgs = _images(3, 224, 224, seed0=3)
    k3 = mine.forward_k(imgs.to(cuda), which_block=3).clone()
    k3_ref = torch.cat([ref.forward_k(dino_vit.preprocess_u8(im, 16).to(cuda), which_block=3) for im in imgs])
    assert ((k3 - k3_ref).norm() / k3_ref.norm()).item() <= 2e-3
    # same image alone or inside a batch gives bitwise identical features (no cross-image coupling)
    k1 = mine.forward_k(imgs[1:2].to(cuda), which_block=3)
    assert torch.equal(k1[0], k3[1])


def test_vitb8_small_image(cuda):
    from oracle import dino_vit
    vit = load_pkg("vit")
    ref = dino_vit.build("dino_vitb8", seed=1)
    mine = vit.DinoViT("dino_vitb8", ref.state_dict(), device=cuda)
    ref = ref.to(cuda)
    imgs = _images(1, 96, 120, seed0=5)
    k_ref = ref.forward_k(dino_vit.preprocess_u8(imgs[0], 8).to(cuda))
    k = mine.forward_k(imgs.to(cuda))
    torch.cuda.synchronize()
    rel = ((k - k_ref).norm() / k_ref.norm()).item()
    print(f"vitb8 96x120: rel-L2 {rel:.3e}")
    assert rel <= 3e-3
