"""Bench / test HARNESS model: a DiT-XL/2-shaped transformer in plain PyTorch.

Not product code — the forward/backward of the reference's models stays in PyTorch and is out of the
hot path's scope.  The reference's DiT (DiT/models.py:145-266) needs `timm`, which is absent here, so
the end-to-end steps/s measurement uses this stand-in with the SAME parameter names, shapes and
count (675,129,632 with the frozen `pos_embed`; depth 28, width 1152, 16 heads, patch 2, 4 latent
channels, learn_sigma -> 8 output channels, 1000 classes + 1 null class), randomly initialised
(the reference's zero-initialised adaLN / final layers are re-randomised N(0, 0.02): with the
constructor's zeros 99.96 % of the gradients are exactly zero).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _modulate(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


class _PatchEmbed(nn.Module):
    def __init__(self, patch, in_ch, width):
        super().__init__()
        self.proj = nn.Conv2d(in_ch, width, kernel_size=patch, stride=patch)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class _Attention(nn.Module):
    def __init__(self, width, heads):
        super().__init__()
        self.heads = heads
        self.qkv = nn.Linear(width, 3 * width)
        self.proj = nn.Linear(width, width)

    def forward(self, x):
        b, t, w = x.shape
        q, k, v = self.qkv(x).view(b, t, 3, self.heads, w // self.heads).permute(2, 0, 3, 1, 4)
        return self.proj(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, t, w))


class _Mlp(nn.Module):
    def __init__(self, width, hidden):
        super().__init__()
        self.fc1 = nn.Linear(width, hidden)
        self.fc2 = nn.Linear(hidden, width)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x), approximate="tanh"))


class _TimestepEmbedder(nn.Module):
    def __init__(self, width, freq=256):
        super().__init__()
        self.freq = freq
        self.mlp = nn.Sequential(nn.Linear(freq, width), nn.SiLU(), nn.Linear(width, width))

    def forward(self, t):
        half = self.freq // 2
        freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
        args = t[:, None].float() * freqs[None]
        emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        return self.mlp(emb.to(self.mlp[0].weight.dtype))


class _LabelEmbedder(nn.Module):
    def __init__(self, classes, width, dropout):
        super().__init__()
        self.embedding_table = nn.Embedding(classes + (dropout > 0), width)
        self.classes, self.dropout = classes, dropout

    def forward(self, y, train):
        if train and self.dropout > 0:
            y = torch.where(torch.rand(y.shape[0], device=y.device) < self.dropout, self.classes, y)
        return self.embedding_table(y)


class _Block(nn.Module):
    def __init__(self, width, heads, ratio):
        super().__init__()
        self.norm1 = nn.LayerNorm(width, elementwise_affine=False, eps=1e-6)
        self.attn = _Attention(width, heads)
        self.norm2 = nn.LayerNorm(width, elementwise_affine=False, eps=1e-6)
        self.mlp = _Mlp(width, int(width * ratio))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(width, 6 * width))

    def forward(self, x, c):
        s1, c1, g1, s2, c2, g2 = self.adaLN_modulation(c).chunk(6, dim=1)
        x = x + g1.unsqueeze(1) * self.attn(_modulate(self.norm1(x), s1, c1))
        return x + g2.unsqueeze(1) * self.mlp(_modulate(self.norm2(x), s2, c2))


class _FinalLayer(nn.Module):
    def __init__(self, width, patch, out_ch):
        super().__init__()
        self.norm_final = nn.LayerNorm(width, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(width, patch * patch * out_ch)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(width, 2 * width))

    def forward(self, x, c):
        shift, scale = self.adaLN_modulation(c).chunk(2, dim=1)
        return self.linear(_modulate(self.norm_final(x), shift, scale))


class DiTXL2Harness(nn.Module):
    def __init__(self, input_size=32, patch=2, in_ch=4, width=1152, depth=28, heads=16, ratio=4.0,
                 classes=1000, class_dropout=0.1, learn_sigma=True):
        super().__init__()
        self.patch, self.in_ch = patch, in_ch
        self.out_ch = in_ch * 2 if learn_sigma else in_ch
        self.x_embedder = _PatchEmbed(patch, in_ch, width)
        self.t_embedder = _TimestepEmbedder(width)
        self.y_embedder = _LabelEmbedder(classes, width, class_dropout)
        tokens = (input_size // patch) ** 2
        self.pos_embed = nn.Parameter(torch.zeros(1, tokens, width), requires_grad=False)
        self.blocks = nn.ModuleList([_Block(width, heads, ratio) for _ in range(depth)])
        self.final_layer = _FinalLayer(width, patch, self.out_ch)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.normal_(p, std=0.02)
            else:
                nn.init.normal_(p, std=0.02)

    def forward(self, x, t, y):
        b, _, h, w = x.shape
        x = self.x_embedder(x) + self.pos_embed
        c = self.t_embedder(t) + self.y_embedder(y, self.training)
        for blk in self.blocks:
            x = blk(x, c)
        x = self.final_layer(x, c)
        gh = h // self.patch
        x = x.view(b, gh, gh, self.patch, self.patch, self.out_ch)
        return torch.einsum("nhwpqc->nchpwq", x).reshape(b, self.out_ch, h, w)


def synthetic_loss(model, latents, t, y, noise, alphas_cumprod):
    """Noise-prediction MSE on the epsilon channels — the same shape of computation as
    GaussianDiffusion.training_losses' MSE term (DiT/diffusion/gaussian_diffusion.py:715-787),
    without the learned-variance VB term (negligible cost)."""
    a = alphas_cumprod[t].view(-1, 1, 1, 1)
    x_t = a.sqrt() * latents + (1 - a).sqrt() * noise
    out = model(x_t, t, y)
    return F.mse_loss(out[:, :latents.shape[1]], noise)


if __name__ == "__main__":
    m = DiTXL2Harness()
    n = sum(p.numel() for p in m.parameters())
    nt = sum(p.numel() for p in m.parameters() if p.requires_grad)
    print(n, nt, len(list(m.parameters())))
