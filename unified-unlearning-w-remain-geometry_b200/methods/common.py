"""Shared machinery of the drop-in entry points: a model adopted into flat buffers + a HotPath.

The reference's loops are `for name, param in model.named_parameters()` Python loops around
`loss.backward()`.  `ModelHotPath` keeps the same call structure (the caller still computes the
loss and calls `.backward()` in PyTorch) and replaces everything between backward and the next
forward by the flat-vector kernels.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from .. import formats
from ..engine import HotPath, OptConfig
from ..flat import FlatParams


def cycle(dl):
    """sfron.py:14-17; DDPM/runners/diffusion.py `cycle`; DiT/forget.py `cycle`."""
    while True:
        for data in dl:
            yield data


def cosine_lr_scheduler(base_lr, current_epoch, T_max):
    """sfron.py:45-46 = DDPM/functions/losses.py:71-72 = DiT/forget.py:35-36."""
    return base_lr * (1 + math.cos(math.pi * current_epoch / T_max)) / 2


def linear_lr_scheduler(base_lr, current_epoch, T_max, base=1):
    """sfron.py:42-43."""
    return base_lr * (1 - current_epoch / T_max) ** base


def expdecay_lr_scheduler(base_lr, current_epoch, T_max, base=2):
    """sfron.py:39-40."""
    return base_lr * (1 - current_epoch / T_max) ** base


class ModelHotPath:
    """A torch model whose parameters / gradients live in flat device vectors, plus the kernels.

    key_prefix: prefix of the reference's file keys relative to `model.named_parameters()` names
    ("module." when the reference wrapped the model in DataParallel and we did not).
    """

    def __init__(self, model: torch.nn.Module, opt: OptConfig, *, ema_mode: str = "none", ema_a: float = 0.0,
                 key_prefix: str = "", device=None, grads_as_views: bool = True):
        self.model = model
        self.flat = FlatParams(model, device, grads_as_views=grads_as_views)
        self.layout = self.flat.layout
        self.key_prefix = key_prefix
        self.hp = HotPath(self.flat.n, self.flat.device, opt, ema_mode=ema_mode, ema_a=ema_a)
        self.frozen_slow: Optional[torch.Tensor] = None
        if ema_mode != "none":
            self.hp.init_slow(self.flat.p)
            if ema_mode == "dit" and self.flat.frozen.numel():
                self.frozen_slow = self.flat.frozen.clone()      # update_ema walks frozen params too

    # ---- gradients -------------------------------------------------------------------------------
    def zero_grad(self) -> None:
        """optimizer.zero_grad() of the reference loops (one memset; grads stay views)."""
        if self.flat.grads_as_views:
            self.flat.zero_grad()
        else:
            for p in self.flat._train_params:
                p.grad = None

    def grads(self) -> torch.Tensor:
        return self.flat.collect_grads()

    # ---- Fisher ----------------------------------------------------------------------------------
    def fisher_accumulate(self, which: str, divisor: float, clip_max_norm: Optional[float] = None) -> None:
        self.hp.fisher_accumulate(which, self.grads(), divisor, clip_max_norm=clip_max_norm)

    def save_fisher(self, which: str, path: str) -> None:
        buf = self.hp.forget_fisher if which == "forget" else self.hp.remain_fisher
        formats.save_fisher(path, self.layout, buf, all_names=self.flat.all_names, prefix=self.key_prefix)

    def load_fisher(self, which: str, path: str) -> None:
        flat = formats.load_fisher(path, self.layout, prefix=self.key_prefix, device=self.flat.device)
        self.hp.set_buffer("forget_fisher" if which == "forget" else "remain_fisher", flat)

    # ---- masks -----------------------------------------------------------------------------------
    def ratio_mask(self, threshold: float) -> Dict[str, object]:
        """Builds the ratio mask on the device; returns it in the reference's dict format and prints
        the reference's sparsity line."""
        mask = self.hp.ratio_mask(threshold)
        zeros = int(self.hp.zero_count[0])
        print(f"Total sparsity th:{threshold} weight:{formats.sparsity_percent(zeros, self.layout.numel)}")
        return formats.ratio_mask_to_dict(self.layout, mask, all_names=self.flat.all_names, prefix=self.key_prefix)

    def load_mask(self, path_or_dict) -> None:
        self.hp.set_buffer("mask", formats.load_mask(path_or_dict, self.layout, prefix=self.key_prefix,
                                                     device=self.flat.device))

    # ---- update ----------------------------------------------------------------------------------
    def forget_step(self, *, use_mask: bool = True, max_norm: Optional[float] = None, lr: Optional[float] = None,
                    mask_order: str = "mask_then_clip") -> None:
        self.hp.forget_step(self.flat.p, self.grads(), use_mask=use_mask, max_norm=max_norm, lr=lr,
                            mask_order=mask_order, zero_grad=self.flat.grads_as_views,
                            p_bf16=self.flat.p_work)

    def remain_step(self, *, max_norm: Optional[float] = None, lr: Optional[float] = None, ema: bool = True) -> None:
        self.hp.remain_step(self.flat.p, self.grads(), max_norm=max_norm, lr=lr, ema=ema,
                            zero_grad=self.flat.grads_as_views, p_bf16=self.flat.p_work)
        if ema and self.frozen_slow is not None:
            self.hp.ema_only(self.flat.frozen, self.frozen_slow)

    def joint_step(self, *, use_mask: bool = True, max_norm: Optional[float] = None, lr: Optional[float] = None,
                   mask_order: str = "clip_then_mask", ema: bool = True) -> None:
        self.hp.joint_step(self.flat.p, self.grads(), use_mask=use_mask, max_norm=max_norm, lr=lr,
                           mask_order=mask_order, ema=ema, zero_grad=self.flat.grads_as_views,
                           p_bf16=self.flat.p_work)
        if ema and self.frozen_slow is not None:
            self.hp.ema_only(self.flat.frozen, self.frozen_slow)

    # ---- state export ------------------------------------------------------------------------------
    def slow_state_dict(self) -> Dict[str, torch.Tensor]:
        """EMA shadow / slow weights per parameter name (EMAHelper.state_dict(), ema.state_dict())."""
        out = dict(self.layout.views(self.hp.slow))
        if self.frozen_slow is not None:
            out.update(self.flat.frozen_layout.views(self.frozen_slow))
        return {self.key_prefix + n: out[n].detach().clone() for n in self.flat.all_names if n in out}

    def optimizer_state_dict(self) -> dict:
        o = self.hp.opt
        trainable = set(self.layout.names)
        n_frozen_before, seen = [], 0
        for name in self.flat.all_names:
            if name in trainable:
                n_frozen_before.append(seen)
            else:
                seen += 1
        step = self.hp.step_count if self.hp.step_dev is None else int(self.hp.step_dev)
        return formats.adam_state_dict(self.layout, self.hp.m, self.hp.v, step, lr=o.lr,
                                       betas=(o.beta1, o.beta2), eps=o.eps, weight_decay=o.weight_decay,
                                       n_frozen_before=n_frozen_before, param_count=len(self.flat.all_names),
                                       decoupled=o.kind == "adamw")
