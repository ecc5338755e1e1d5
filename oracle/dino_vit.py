"""ORACLE (test infrastructure, not product): fp32 PyTorch restatement of the DINO VisionTransformer.

The network itself is third-party code that the reference pulls at run time with
``torch.hub.load('facebookresearch/dino:main', name)`` (reference extract/extract_utils.py:40-50); the
source is not vendored under /root/reference and cannot be fetched here (no network), so this file restates
the *published* architecture of ``facebookresearch/dino@main vision_transformer.py`` (unpinned upstream):

  * PatchEmbed = Conv2d(3, d, kernel=P, stride=P), flatten, prepend CLS, add positional embedding that is
    bicubically interpolated from the training grid with ``scale_factor=((h0+0.1)/sqrt(N0), (w0+0.1)/sqrt(N0))``
    (the ``+0.1`` quirk), CLS slot untouched.
  * depth x Block: x += Attn(LN1(x)); x += MLP(LN2(x)); LayerNorm eps = 1e-6, qkv_bias=True,
    softmax(q k^T * dh^-0.5) v, proj; MLP = Linear(d,4d) -> GELU(erf) -> Linear(4d,d).
  * the reference hooks ``blocks[which_block].attn.qkv`` (extract/extract.py:49-53) and keeps the K third of
    its output for the patch tokens (extract/extract.py:96-98) == qkv_out[:, 1:, d:2d].

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Parity status: *unpinned* by the reference (it ships no tests or golden vectors); cross-checked offline against
``transformers.ViTModel`` block math in tests/test_oracle_vit.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

# name -> (patch, dim, depth, heads)   (upstream hubconf.py: dino_vits16 / vits8 / vitb16 / vitb8)
ARCHS = {
    "dino_vits16": (16, 384, 12, 6),
    "dino_vits8": (8, 384, 12, 6),
    "dino_vitb16": (16, 768, 12, 12),
    "dino_vitb8": (8, 768, 12, 12),
}
TRAIN_IMG = 224  # upstream img_size=[224] => pos-embed grid 14x14 (P16) or 28x28 (P8)


@dataclass
class VitCfg:
    patch: int
    dim: int
    depth: int
    heads: int
    mlp_ratio: int = 4
    eps: float = 1e-6

    @property
    def grid0(self) -> int:
        return TRAIN_IMG // self.patch


def cfg_for(name: str) -> VitCfg:
    name = name.lower()
    if name not in ARCHS:
        raise ValueError(f"Cannot get model: {name}")
    p, d, L, h = ARCHS[name]
    return VitCfg(p, d, L, h)


class Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.scale = (dim // heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)) * self.scale
        attn = attn.softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, heads, mlp_ratio, eps):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = Mlp(dim, dim * mlp_ratio)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        return x


class PatchEmbed(nn.Module):
    def __init__(self, patch, dim):
        super().__init__()
        self.patch_size = patch
        self.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


def _trunc_normal_(t, std=0.02):
    return nn.init.trunc_normal_(t, std=std)


class DinoViT(nn.Module):
    """State-dict compatible with upstream DINO checkpoints (same parameter names)."""

    def __init__(self, cfg: VitCfg):
        super().__init__()
        self.cfg = cfg
        d = cfg.dim
        self.patch_embed = PatchEmbed(cfg.patch, d)
        n0 = cfg.grid0 * cfg.grid0
        self.cls_token = nn.Parameter(torch.zeros(1, 1, d))
        self.pos_embed = nn.Parameter(torch.zeros(1, n0 + 1, d))
        self.blocks = nn.ModuleList([Block(d, cfg.heads, cfg.mlp_ratio, cfg.eps) for _ in range(cfg.depth)])
        self.norm = nn.LayerNorm(d, eps=cfg.eps)
        # upstream init: trunc_normal_(std=.02) for pos/cls/Linear weights, zero biases, LN (1, 0); conv default
        _trunc_normal_(self.pos_embed)
        _trunc_normal_(self.cls_token)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                _trunc_normal_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

    # upstream: interpolate_pos_encoding(x, w, h) with (w, h) = x.shape[2:] i.e. "w" is the image HEIGHT
    def interpolate_pos_encoding(self, npatch, Hc, Wc):
        N = self.pos_embed.shape[1] - 1
        if npatch == N and Hc == Wc:
            return self.pos_embed
        class_pos = self.pos_embed[:, 0]
        patch_pos = self.pos_embed[:, 1:]
        dim = patch_pos.shape[-1]
        P = self.cfg.patch
        w0, h0 = Hc // P + 0.1, Wc // P + 0.1
        s = int(math.sqrt(N))
        patch_pos = F.interpolate(
            patch_pos.reshape(1, s, s, dim).permute(0, 3, 1, 2),
            scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)),
            mode="bicubic",
        )
        assert int(w0) == patch_pos.shape[-2] and int(h0) == patch_pos.shape[-1]
        patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
        return torch.cat((class_pos.unsqueeze(0), patch_pos), dim=1)

    def prepare_tokens(self, x):
        B, _, Hc, Wc = x.shape
        x = self.patch_embed(x)
        x = torch.cat((self.cls_token.expand(B, -1, -1), x), dim=1)
        return x + self.interpolate_pos_encoding(x.shape[1] - 1, Hc, Wc)

    @torch.no_grad()
    def forward_k(self, images: torch.Tensor, which_block: int = -1) -> torch.Tensor:
        """images (B,3,Hc,Wc) normalised fp32, Hc/Wc multiples of P -> K features (B, N, d) fp32.

        Equals the reference's ``output_dict['k']`` (extract/extract.py:94-98). Blocks after ``which_block`` and
        the final LayerNorm do not influence the hooked tensor and are skipped.
        """
        d = self.cfg.dim
        x = self.prepare_tokens(images)
        blk_idx = which_block % len(self.blocks)
        for i, blk in enumerate(self.blocks):
            if i == blk_idx:
                qkv = blk.attn.qkv(blk.norm1(x))
                return qkv[:, 1:, d:2 * d].contiguous()
            x = blk(x)
        raise AssertionError

    @torch.no_grad()
    def forward(self, images: torch.Tensor) -> torch.Tensor:
        """Upstream VisionTransformer.forward: all blocks, final LayerNorm, CLS token (B, d) -- what the reference's
        extract_bbox_features calls on every crop (extract/extract.py:537-541)."""
        x = self.prepare_tokens(images)
        for blk in self.blocks:
            x = blk(x)
        return self.norm(x)[:, 0]

    @torch.no_grad()
    def forward_tokens(self, images: torch.Tensor, n_blocks: int) -> torch.Tensor:
        """Residual stream after ``n_blocks`` full blocks (debug / per-layer parity)."""
        x = self.prepare_tokens(images)
        for blk in self.blocks[:n_blocks]:
            x = blk(x)
        return x


def build(name: str, seed: int = 0) -> DinoViT:
    """Random-init model of the named architecture (upstream init recipe), deterministic in ``seed``."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        m = DinoViT(cfg_for(name)).eval()
    finally:
        torch.random.set_rng_state(g)
    for p in m.parameters():
        p.requires_grad_(False)
    return m


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def preprocess_u8(img_hwc_u8: torch.Tensor, patch: int) -> torch.Tensor:
    """uint8 RGB (H,W,3) -> (1,3,Hc,Wc) fp32: ToTensor + Normalize (extract_utils.py:53-59) then the top-left
    crop to patch multiples (extract/extract.py:82-88)."""
    H, W, _ = img_hwc_u8.shape
    Hc, Wc = (H // patch) * patch, (W // patch) * patch
    x = img_hwc_u8.permute(2, 0, 1).to(torch.float32).div(255.0)
    mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
    x = (x - mean) / std
    return x[None, :, :Hc, :Wc].contiguous()
