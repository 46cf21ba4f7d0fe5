"""Bench / test HARNESS model: the Stable Diffusion v1.x latent U-Net in plain PyTorch (BASELINE config 4).

Not product code — forward/backward stays in PyTorch and is outside the hot path.  The reference's network
(SD/ldm/modules/diffusionmodules/openaimodel.py:428-847 configured by SD/configs/stable-diffusion/
v1-inference.yaml:29-44) lives under /root/reference, which does not exist on the GPU box, and its package
imports omegaconf / pytorch_lightning, absent here.  This stand-in has the SAME state-dict names, shapes, order
and count — 859,520,964 parameters in 686 tensors: 320 base channels x (1,2,4,4), 2 res blocks per level,
8-head spatial transformers with 768-wide cross-attention at the three finest levels — so Fisher / mask files keyed
by U-Net-local names (SD/train-scripts/generate_fisher.py:73-79) fit either.  `tests/test_harness_models.py`
compares names / shapes / order and outputs with the reference module when /root/reference is present.
Gradient checkpointing (the reference config's use_checkpoint) is a memory knob, off here: 180 GB of HBM.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class _GN32(nn.GroupNorm):
    """GroupNorm evaluated in fp32 whatever the activation dtype."""

    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


class _Res(nn.Module):
    takes = "emb"

    def __init__(self, c_in: int, c_out: int, emb: int):
        super().__init__()
        self.in_layers = nn.Sequential(_GN32(32, c_in), nn.SiLU(), nn.Conv2d(c_in, c_out, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb, c_out))
        self.out_layers = nn.Sequential(_GN32(32, c_out), nn.SiLU(), nn.Dropout(0.0), nn.Conv2d(c_out, c_out, 3, padding=1))
        self.skip_connection = nn.Identity() if c_in == c_out else nn.Conv2d(c_in, c_out, 1)

    def forward(self, x, emb):
        h = self.in_layers(x) + self.emb_layers(emb).type(x.dtype)[:, :, None, None]
        return self.skip_connection(x) + self.out_layers(h)


class _Attention(nn.Module):
    """Multi-head attention of a token sequence over itself or over a context sequence."""

    def __init__(self, dim: int, ctx_dim: int, heads: int):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim, dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(dim, dim), nn.Dropout(0.0))

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        b, n, d = x.shape
        split = lambda t: t.view(b, t.shape[1], self.heads, d // self.heads).transpose(1, 2)
        o = F.scaled_dot_product_attention(split(self.to_q(x)), split(self.to_k(ctx)), split(self.to_v(ctx)))
        return self.to_out(o.transpose(1, 2).reshape(b, n, d))


class _GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, 2 * inner)

    def forward(self, x):
        a, gate = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(gate)


class _FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.Sequential(_GEGLU(dim, 4 * dim), nn.Dropout(0.0), nn.Linear(4 * dim, dim))

    def forward(self, x):
        return self.net(x)


class _TransformerBlock(nn.Module):
    def __init__(self, dim: int, ctx_dim: int, heads: int):
        super().__init__()
        self.attn1 = _Attention(dim, dim, heads)
        self.ff = _FeedForward(dim)
        self.attn2 = _Attention(dim, ctx_dim, heads)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        return x + self.ff(self.norm3(x))


class _SpatialTransformer(nn.Module):
    takes = "ctx"

    def __init__(self, ch: int, ctx_dim: int, heads: int, depth: int = 1):
        super().__init__()
        self.norm = nn.GroupNorm(32, ch, eps=1e-6)
        self.proj_in = nn.Conv2d(ch, ch, 1)
        self.transformer_blocks = nn.ModuleList(_TransformerBlock(ch, ctx_dim, heads) for _ in range(depth))
        self.proj_out = nn.Conv2d(ch, ch, 1)

    def forward(self, x, ctx):
        b, c, hh, ww = x.shape
        t = self.proj_in(self.norm(x)).flatten(2).transpose(1, 2)
        for blk in self.transformer_blocks:
            t = blk(t, ctx)
        return x + self.proj_out(t.transpose(1, 2).reshape(b, c, hh, ww))


class _Down(nn.Module):
    takes = None

    def __init__(self, ch: int):
        super().__init__()
        self.op = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.op(x)


class _Up(nn.Module):
    takes = None

    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2, mode="nearest"))


class _Stage(nn.Sequential):
    """A numbered chain of layers; each gets the extra input it declares (`takes`)."""

    def forward(self, h, emb, ctx):
        for layer in self:
            kind = getattr(layer, "takes", None)
            h = layer(h, emb) if kind == "emb" else layer(h, ctx) if kind == "ctx" else layer(h)
        return h


class SDUNet(nn.Module):
    def __init__(self, in_channels=4, out_channels=4, model_channels=320, channel_mult=(1, 2, 4, 4), num_res_blocks=2,
                 attention_levels=(0, 1, 2), num_heads=8, context_dim=768, transformer_depth=1):
        super().__init__()
        self.model_channels = model_channels
        emb = 4 * model_channels
        self.time_embed = nn.Sequential(nn.Linear(model_channels, emb), nn.SiLU(), nn.Linear(emb, emb))

        def attn(ch):
            return _SpatialTransformer(ch, context_dim, num_heads, transformer_depth)

        self.input_blocks = nn.ModuleList([_Stage(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        ch, skip_chs = model_channels, [model_channels]
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [_Res(ch, mult * model_channels, emb)]
                ch = mult * model_channels
                if level in attention_levels:
                    layers.append(attn(ch))
                self.input_blocks.append(_Stage(*layers))
                skip_chs.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(_Stage(_Down(ch)))
                skip_chs.append(ch)
        self.middle_block = _Stage(_Res(ch, ch, emb), attn(ch), _Res(ch, ch, emb))
        self.output_blocks = nn.ModuleList()
        for level in reversed(range(len(channel_mult))):
            for i in range(num_res_blocks + 1):
                layers = [_Res(ch + skip_chs.pop(), channel_mult[level] * model_channels, emb)]
                ch = channel_mult[level] * model_channels
                if level in attention_levels:
                    layers.append(attn(ch))
                if level and i == num_res_blocks:
                    layers.append(_Up(ch))
                self.output_blocks.append(_Stage(*layers))
        self.out = nn.Sequential(_GN32(32, ch), nn.SiLU(), nn.Conv2d(model_channels, out_channels, 3, padding=1))

    def forward(self, x, timesteps, context):
        half = self.model_channels // 2
        freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=x.device) / half)
        ang = timesteps[:, None].float() * freqs[None]
        emb = self.time_embed(torch.cat([ang.cos(), ang.sin()], dim=-1).to(self.time_embed[0].weight.dtype))
        hs, h = [], x
        for stage in self.input_blocks:
            h = stage(h, emb, context)
            hs.append(h)
        h = self.middle_block(h, emb, context)
        for stage in self.output_blocks:
            h = stage(torch.cat([h, hs.pop()], dim=1), emb, context)
        return self.out(h)

    def randomise_zero_layers(self, std: float = 0.02) -> None:
        """A constructor-initialised SD U-Net zeroes every block's last conv (`zero_module`); with those zeros
        most gradients vanish identically.  The synthetic-data runs re-randomise them, as the DiT harness does."""
        with torch.no_grad():
            for m in self.modules():
                if isinstance(m, _Res):
                    nn.init.normal_(m.out_layers[3].weight, std=std)
                elif isinstance(m, _SpatialTransformer):
                    nn.init.normal_(m.proj_out.weight, std=std)
            nn.init.normal_(self.out[2].weight, std=std)


def sd_alphas_cumprod(device, steps=1000, linear_start=0.00085, linear_end=0.012):
    """The "linear" schedule of latent diffusion: linear in sqrt(beta) (v1-inference.yaml:5-6)."""
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, steps, dtype=torch.float64) ** 2
    return torch.cumprod(1 - betas, 0).float().to(device)


def guided_eps_loss(model, z0, t, ctx, null_ctx, noise, ac, cond_scale=7.5):
    """Fisher loss of SD/train-scripts/generate_fisher.py:47-66: q_sample, conditional and unconditional
    passes, guided prediction (1+c)*cond - c*null, negative MSE against the noise."""
    a = ac.index_select(0, t).view(-1, 1, 1, 1).to(z0.dtype)
    zt = a.sqrt() * z0 + (1 - a).sqrt() * noise
    pred = (1 + cond_scale) * model(zt, t, ctx) - cond_scale * model(zt, t, null_ctx)
    return -F.mse_loss(noise, pred)


def eps_mse_loss(model, z0, t, ctx, noise, ac):
    """Plain noise-prediction MSE (LatentDiffusion.p_losses, eps parameterisation, l2): the remain loss of
    nsfw_removal.py:165-170."""
    a = ac.index_select(0, t).view(-1, 1, 1, 1).to(z0.dtype)
    return F.mse_loss(model(a.sqrt() * z0 + (1 - a).sqrt() * noise, t, ctx), noise)
