"""Bench / test HARNESS model: the class-conditional DDPM CIFAR-10 U-Net in plain PyTorch (BASELINE config 2).

Not product code — forward/backward stays in PyTorch and is outside the hot path.  The reference's network
(DDPM/models/diffusion.py:195-413) lives under /root/reference, which does not exist on the GPU box, so the
end-to-end measurement uses this stand-in with the SAME state-dict names, shapes, order and count
(38,632,323 parameters in 334 tensors at ch 128, ch_mult (1,2,2,2), 2 res blocks, attention at 16x16,
10 classes + a learned null class for classifier-free guidance).  `tests/test_harness_models.py` checks
names / shapes / order — and, with the reference's weights loaded, the outputs — against the reference
module whenever /root/reference is present.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def sinusoidal_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(torch.arange(half, dtype=torch.float32, device=t.device) * (-math.log(10000) / (half - 1)))
    ang = t.float()[:, None] * freqs[None]
    return torch.cat([ang.sin(), ang.cos()], dim=1)


def _gn(ch: int) -> nn.GroupNorm:
    return nn.GroupNorm(32, ch, eps=1e-6)


class _Dense2(nn.Module):
    """`.dense.0`, `.dense.1`: Linear -> swish -> Linear."""

    def __init__(self, d_in: int, d_out: int):
        super().__init__()
        self.dense = nn.ModuleList([nn.Linear(d_in, d_out), nn.Linear(d_out, d_out)])

    def forward(self, x):
        return self.dense[1](F.silu(self.dense[0](x)))


class _Res(nn.Module):
    def __init__(self, c_in: int, c_out: int, emb: int, dropout: float):
        super().__init__()
        self.norm1 = _gn(c_in)
        self.conv1 = nn.Conv2d(c_in, c_out, 3, padding=1)
        self.temb_cemb_proj = nn.Linear(2 * emb, c_out)       # one projection of [time ; class] embeddings
        self.norm2 = _gn(c_out)
        self.conv2 = nn.Conv2d(c_out, c_out, 3, padding=1)
        if c_in != c_out:
            self.nin_shortcut = nn.Conv2d(c_in, c_out, 1)
        self.p_drop = dropout

    def forward(self, x, emb_act):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.temb_cemb_proj(emb_act)[:, :, None, None]
        h = F.dropout(F.silu(self.norm2(h)), self.p_drop, self.training)
        h = self.conv2(h)
        return (self.nin_shortcut(x) if hasattr(self, "nin_shortcut") else x) + h


class _Attn(nn.Module):
    """Single-head self-attention over the h*w positions (1x1 conv projections)."""

    def __init__(self, ch: int):
        super().__init__()
        self.norm = _gn(ch)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(ch, ch, 1) for _ in range(4))

    def forward(self, x):
        b, c, hh, ww = x.shape
        y = self.norm(x)
        q, k, v = (f(y).flatten(2).transpose(1, 2) for f in (self.q, self.k, self.v))      # b, hw, c
        o = F.scaled_dot_product_attention(q, k, v)                                           # scale c**-0.5
        return x + self.proj_out(o.transpose(1, 2).reshape(b, c, hh, ww))


class _Resample(nn.Module):
    def __init__(self, ch: int, down: bool):
        super().__init__()
        self.down = down
        self.conv = nn.Conv2d(ch, ch, 3, stride=2 if down else 1, padding=0 if down else 1)

    def forward(self, x):
        if self.down:
            return self.conv(F.pad(x, (0, 1, 0, 1)))
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Level(nn.Module):
    def __init__(self):
        super().__init__()
        self.block = nn.ModuleList()
        self.attn = nn.ModuleList()


class _Mid(nn.Module):
    def __init__(self, ch, emb, dropout):
        super().__init__()
        self.block_1 = _Res(ch, ch, emb, dropout)
        self.attn_1 = _Attn(ch)
        self.block_2 = _Res(ch, ch, emb, dropout)


class DDPMCondUNet(nn.Module):
    def __init__(self, ch=128, ch_mult=(1, 2, 2, 2), num_res_blocks=2, attn_resolutions=(16,), resolution=32,
                 in_channels=3, out_ch=3, n_classes=10, dropout=0.1, cond_drop_prob=0.1):
        super().__init__()
        self.ch, self.resolution, self.cond_drop_prob = ch, resolution, cond_drop_prob
        self.num_res_blocks, self.levels = num_res_blocks, len(ch_mult)
        emb = 4 * ch
        self.temb = _Dense2(ch, emb)
        self.classes_emb = nn.Embedding(n_classes, ch)
        self.null_classes_emb = nn.Parameter(torch.randn(ch))
        self.cemb = _Dense2(ch, emb)
        self.conv_in = nn.Conv2d(in_channels, ch, 3, padding=1)

        widths = [ch * m for m in ch_mult]
        self.down = nn.ModuleList()
        res, c = resolution, ch
        skips = [ch]
        for lvl, w in enumerate(widths):
            level = _Level()
            for _ in range(num_res_blocks):
                level.block.append(_Res(c, w, emb, dropout))
                c = w
                if res in attn_resolutions:
                    level.attn.append(_Attn(c))
                skips.append(c)
            if lvl != len(widths) - 1:
                level.downsample = _Resample(c, down=True)
                res //= 2
                skips.append(c)
            self.down.append(level)
        self.mid = _Mid(c, emb, dropout)
        ups = []
        for lvl in reversed(range(len(widths))):
            level = _Level()
            for _ in range(num_res_blocks + 1):
                level.block.append(_Res(c + skips.pop(), widths[lvl], emb, dropout))
                c = widths[lvl]
                if res in attn_resolutions:
                    level.attn.append(_Attn(c))
            if lvl != 0:
                level.upsample = _Resample(c, down=False)
                res *= 2
            ups.append(level)
        self.up = nn.ModuleList(reversed(ups))            # up.0 is the full-resolution level
        self.norm_out = _gn(c)
        self.conv_out = nn.Conv2d(c, out_ch, 3, padding=1)

    def _net(self, x, t, c, cond_drop_prob):
        temb = self.temb(sinusoidal_embedding(t, self.ch))
        cls = self.classes_emb(c)
        if cond_drop_prob > 0:
            if cond_drop_prob >= 1:
                keep = torch.zeros(x.shape[0], dtype=torch.bool, device=x.device)
            else:
                keep = torch.rand(x.shape[0], device=x.device) < (1 - cond_drop_prob)
            cls = torch.where(keep[:, None], cls, self.null_classes_emb[None].expand_as(cls))
        emb_act = F.silu(torch.cat([temb, self.cemb(cls)], dim=-1))
        hs = [self.conv_in(x)]
        for lvl, level in enumerate(self.down):
            for i, blk in enumerate(level.block):
                h = blk(hs[-1], emb_act)
                if len(level.attn):
                    h = level.attn[i](h)
                hs.append(h)
            if lvl != self.levels - 1:
                hs.append(level.downsample(hs[-1]))
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(hs[-1], emb_act)), emb_act)
        for lvl in reversed(range(self.levels)):
            level = self.up[lvl]
            for i, blk in enumerate(level.block):
                h = blk(torch.cat([h, hs.pop()], dim=1), emb_act)
                if len(level.attn):
                    h = level.attn[i](h)
            if lvl != 0:
                h = level.upsample(h)
        return self.conv_out(F.silu(self.norm_out(h)))

    def forward(self, x, t, c, mode="train", cond_drop_prob=None, cond_scale=1.0):
        """mode "train": one pass with class dropout; "test": classifier-free-guided output
        (1+s)*cond - s*null, two passes (the form generate_fisher / generate_mask differentiate)."""
        if mode == "train":
            return self._net(x, t, c, self.cond_drop_prob if cond_drop_prob is None else cond_drop_prob)
        cond = self._net(x, t, c, 0.0)
        if cond_scale == 0:
            return cond
        return (1 + cond_scale) * cond - cond_scale * self._net(x, t, c, 1.0)


def ddpm_alphas_cumprod(device, steps=1000, beta_start=1e-4, beta_end=0.02):
    betas = torch.linspace(beta_start, beta_end, steps, dtype=torch.float64)
    return torch.cumprod(1 - betas, 0).float().to(device)


def eps_loss(model, x0, t, c, noise, ac, *, mode="train", cond_scale=2.0, keepdim=False):
    """Noise-prediction loss, summed over pixels, mean over the batch
    (DDPM/functions/losses.py:22-38; the "test"-mode form of runners/diffusion.py:1255-1265)."""
    a = ac.index_select(0, t).view(-1, 1, 1, 1)
    xt = x0 * a.sqrt() + noise * (1.0 - a).sqrt()
    out = model(xt, t.float(), c, mode=mode, cond_scale=cond_scale)
    per = (noise - out).square().sum(dim=(1, 2, 3))
    return per if keepdim else per.mean(dim=0)
