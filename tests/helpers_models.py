"""Tiny models / loaders shared by the tests; the same definitions tests/golden/make_golden.py used
when it recorded the reference runs (kept in sync by hand: 554-parameter TinyNet, TinyDiT)."""
import torch
import torch.nn as nn
from torch.utils.data import DataLoader, TensorDataset


class TinyNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(3, 8, 3, padding=1)
        self.fc = nn.Linear(32, 10)

    def forward(self, x):
        x = torch.relu(self.conv(x))
        x = torch.nn.functional.adaptive_avg_pool2d(x, 2).flatten(1)
        return self.fc(x)


class TinyDiT(nn.Module):
    def __init__(self):
        super().__init__()
        self.pos_embed = nn.Parameter(torch.randn(1, 6, 8), requires_grad=False)
        self.fc = nn.Linear(33, 10)
        self.out = nn.Linear(10, 4)


class TinyCond(nn.Module):
    """Parameter structure of the 411-parameter stand-in network make_golden.py's `ddpm_runner` part
    gave the reference's Diffusion runner (named_parameters order: null_classes_emb first)."""

    def __init__(self):
        super().__init__()
        self.conv_in = nn.Conv2d(3, 6, 3, padding=1)
        self.classes_emb = nn.Embedding(10, 6)
        self.null_classes_emb = nn.Parameter(torch.randn(6))
        self.temb = nn.Linear(1, 6)
        self.conv_out = nn.Conv2d(6, 3, 3, padding=1)


class TinyCondNet(TinyCond):
    """TinyCond with the call interface of the reference's Conditional_Model (`model(x, t, c, mode="train"|"test",
    cond_drop_prob=, cond_scale=)`, classifier-free guidance through `null_classes_emb`), so that the runner-level
    drop-in (sfron_b200.methods.ddpm.Diffusion) can be driven end to end on synthetic data."""

    def __init__(self, config=None):
        super().__init__()

    def _eps(self, x, t, c, drop):
        cemb = self.classes_emb(c)
        if drop > 0:
            keep = torch.rand(x.shape[0], device=x.device) < (1 - drop)
            cemb = torch.where(keep[:, None], cemb, self.null_classes_emb[None].expand_as(cemb))
        h = self.conv_in(x) + (cemb + self.temb(t[:, None] / 1000.0))[:, :, None, None]
        return self.conv_out(torch.tanh(h))

    def forward(self, x, t, c, mode="train", cond_drop_prob=None, cond_scale=None):
        if mode == "train":
            return self._eps(x, t, c, 0.1 if cond_drop_prob is None else cond_drop_prob)
        eps = self._eps(x, t, c, 0.0)
        return eps if not cond_scale else (1 + cond_scale) * eps - cond_scale * self._eps(x, t, c, 1.0)


class TinyLatentUNet(nn.Module):
    """Parameter structure of the 586-parameter stand-in U-Net make_golden.py's `sd_scripts` part put behind
    `model.model.diffusion_model` (keys are U-Net-local, as in the SD Fisher / mask files)."""

    def __init__(self):
        super().__init__()
        self.conv_in = nn.Conv2d(4, 6, 3, padding=1)
        self.temb = nn.Linear(1, 6)
        self.attn2 = nn.Module()
        self.attn2.to_q = nn.Linear(6, 6, bias=False)
        self.attn2.to_k = nn.Linear(8, 6, bias=False)
        self.attn2.to_v = nn.Linear(8, 6, bias=False)
        self.conv_out = nn.Conv2d(6, 4, 3, padding=1)


def loaders(seed):
    g = torch.Generator().manual_seed(seed)
    fx, fy = torch.randn(12, 3, 8, 8, generator=g), torch.randint(0, 10, (12,), generator=g)
    rx, ry = torch.randn(16, 3, 8, 8, generator=g), torch.randint(0, 10, (16,), generator=g)
    mk = lambda x, y: DataLoader(TensorDataset(x, y), batch_size=4, shuffle=False)
    return dict(forget_train=mk(fx, fy), retain_train=mk(rx, ry), forget_valid=None, retain_valid=None)


def inject(model, flat_grad):
    """A 'loss' whose gradient w.r.t. the trainable parameters is exactly `flat_grad`:
    sum_i <p_i, g_i>.  Lets a test drive a real backward pass with recorded gradients."""
    loss, off = 0.0, 0
    for p in model.parameters():
        if p.requires_grad:
            g = flat_grad[off:off + p.numel()].reshape(p.shape).to(p.device)
            loss = loss + (p * g).sum()
            off += p.numel()
    assert off == flat_grad.numel()
    return loss
