"""Drop-in for Classification/unlearn/{unlearn_method,sfron,salun}.py — same class names,
constructor, `prepare_unlearn / get_unlearned_model / get_params`, hyper-parameter attributes and
Fisher cache files, with the per-parameter Python loops replaced by the flat-vector kernels.

    method = create_unlearn_method("SFRon")(model, loss_function, save_path, args)   # main_random.py:106
    method.prepare_unlearn(unlearn_dataloaders)                                        # :107
    model = method.get_unlearned_model()                                               # :108

The model must already be on a CUDA device (the reference calls `.cuda()` itself, sfron.py:195).
Evaluation helpers of the reference (`validate`, MIA, CSV) are out of scope; `validate_fn` may be
set to get the periodic validation calls of sfron.py:243-247.
"""
from __future__ import annotations

import os
import time
from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import formats
from ..engine import OptConfig
from .common import (ModelHotPath, cosine_lr_scheduler, cycle, expdecay_lr_scheduler, linear_lr_scheduler)


class UnlearnMethod:
    """Classification/unlearn/unlearn_method.py:4-21."""

    def __init__(self, model, loss_function, save_path, args) -> None:
        self.unlearn_dataloaders = None
        self.model = model
        self.loss_function = loss_function
        self.save_path = save_path
        self.args = args
        self.params = {}

    def prepare_unlearn(self, unlearn_dataloaders: dict) -> None:
        self.unlearn_dataloaders = unlearn_dataloaders

    def get_unlearned_model(self) -> nn.Module:
        return self.model

    def get_params(self) -> dict:
        return self.params


class AdaptiveLoss(torch.nn.Module):
    """sfron.py:48-65 — stays in PyTorch (it is part of the loss, i.e. of the backward graph)."""

    def __init__(self, loss_function, lambd=1, reduction="mean"):
        super().__init__()
        self.loss_function = loss_function  # reduction=none
        self.lambd = lambd
        self.reduction = reduction

    def forward(self, predict, target):
        ori_loss = self.loss_function(predict, target)
        coef = 1 / (torch.pow(ori_loss.detach().clone(), self.lambd) + 1e-15)
        ad_loss = (coef / coef.sum()) * ori_loss * predict.shape[0]
        if self.reduction == "mean":
            ad_loss = ad_loss.mean()
        elif self.reduction == "sum":
            ad_loss = ad_loss.sum()
        return ad_loss


def _device_of(model):
    return next(model.parameters()).device


class SFRon(UnlearnMethod):
    """Classification/unlearn/sfron.py:67-354."""

    def __init__(self, model, loss_function, save_path, args) -> None:
        super().__init__(model, loss_function, save_path, args)
        self.num_classes = args.num_classes
        self.seed = args.seed
        self.eval = True
        self.forget_loss_function = None
        self.retain_loss_function = None
        self.weight_saliency_mask = None
        self.validate_fn = None
        # CIFAR10 10% defaults (sfron.py:100-123)
        self.opt = "sgd"
        self.momentum = 0.9
        self.weight_decay = 5e-4
        self.retain_lr = 0.01
        self.n_iters = 1500
        self.unlearn_loss = "adaga"
        self.forget_freq = 5
        self.forget_alpha = 25
        self.max_norm = 7.0
        self.ema_enabled = True
        self.ema_beta = 1.0
        self.sched = "cosine"
        self.lambd = 0.5
        self.mask = True
        self.th = 1
        self.log_freq = 500
        # extension: replay whole iterations (forward, backward, kernels) from CUDA graphs — the loop is launch-bound
        # at ResNet-18 size.  The per-iteration cosine learning rate then lives on the device (HotPath.set_lr_schedule).
        self.cuda_graph = bool(getattr(args, "cuda_graph", False))
        self._mhp: Optional[ModelHotPath] = None

    # ---- flat state, created on first use (hyper-parameters may be edited after __init__) ---------
    def _hot_path(self) -> ModelHotPath:
        if self._mhp is None:
            kind = "sgd" if self.opt == "sgd" else "adamw"
            cfg = OptConfig(kind=kind, lr=self.retain_lr, weight_decay=self.weight_decay,
                            momentum=self.momentum if kind == "sgd" else 0.0)
            self._mhp = ModelHotPath(self.model, cfg, ema_mode="slowfast" if self.ema_enabled else "none",
                                     ema_a=self.ema_beta, device=_device_of(self.model))
        return self._mhp

    def prepare_unlearn(self, unlearn_dataloaders: dict) -> None:
        self.unlearn_dataloaders = unlearn_dataloaders
        if self.unlearn_loss == "adaga":
            self.forget_loss_function = AdaptiveLoss(nn.CrossEntropyLoss(reduction="none"), lambd=self.lambd)
        elif self.unlearn_loss == "ga":
            self.forget_loss_function = self.loss_function
        self.retain_loss_function = self.loss_function
        if self.mask:
            self.weight_saliency_mask = self.get_weight_saliency_mask(
                forget_loader=self.unlearn_dataloaders["forget_train"],
                remain_loader=self.unlearn_dataloaders["retain_train"], threshold=self.th)
        else:
            self.weight_saliency_mask = None

    def _fisher(self, which: str, loader, path: str) -> None:
        """sfron.py:268-293 / 295-320: F += (batch-mean grad)**2 / len(loader), cached on disk."""
        mhp = self._hot_path()
        if os.path.exists(path):
            mhp.load_fisher(which, path)
            return
        dev = mhp.flat.device
        mhp.hp.buffer(f"{which}_fisher").zero_()
        self.model.eval()
        for image, target in loader:
            image, target = image.to(dev), target.to(dev)
            loss = self.loss_function(self.model(image), target)
            mhp.zero_grad()
            loss.backward()
            mhp.fisher_accumulate(which, float(len(loader)))
        mhp.zero_grad()
        mhp.save_fisher(which, path)

    def get_weight_saliency_mask(self, forget_loader, remain_loader, threshold):
        """sfron.py:262-336.  Returns the mask in the reference's {name: bool tensor} format; the
        device copy used by the loop stays in the flat `mask` buffer."""
        self._fisher("forget", forget_loader, os.path.join(self.save_path, "forget_fisher.pt"))
        self._fisher("remain", remain_loader, os.path.join(self.save_path, "remain_fisher.pt"))
        return self._hot_path().ratio_mask(threshold)

    def get_unlearned_model(self) -> nn.Module:
        """sfron.py:151-260."""
        mhp = self._hot_path()
        dev = mhp.flat.device
        retain_train_iter = cycle(self.unlearn_dataloaders["retain_train"])
        forget_train_iter = cycle(self.unlearn_dataloaders["forget_train"])
        lr_scheduler = {"cosine": cosine_lr_scheduler, "linear": linear_lr_scheduler,
                        "expdecay": expdecay_lr_scheduler}[self.sched]
        # the reference steps a CosineAnnealingLR every iteration (sfron.py:172-174,259): the same torch
        # scheduler on a one-element dummy optimizer yields bit-identical learning rates
        dummy = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], self.retain_lr)
        scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(dummy, T_max=self.n_iters)
        if self.ema_enabled:
            mhp.hp.init_slow(mhp.flat.p)              # ori_model = deepcopy(self.model)
        mhp.zero_grad()
        if self.cuda_graph:
            rates = []
            for _ in range(self.n_iters):             # torch's own scheduler -> the exact per-iteration rates
                rates.append(dummy.param_groups[0]["lr"])
                dummy.step()
                scheduler.step()
            return self._unlearn_graphed(mhp, rates, lr_scheduler, forget_train_iter, retain_train_iter)
        log_forget = log_remain = 0
        run_forget = run_remain = 0.0
        start_time = time.time()
        for step in range(0, self.n_iters):
            lr = dummy.param_groups[0]["lr"]
            self.model.train()
            if step % self.forget_freq == 0:
                cur_forget_alpha = lr_scheduler(self.forget_alpha, step, self.n_iters)
                x_forget, y_forget = next(forget_train_iter)
                x_forget, y_forget = x_forget.to(dev), y_forget.to(dev)
                outputs = self.model(x_forget)
                ori_forget_loss = -self.forget_loss_function(outputs, y_forget)
                (cur_forget_alpha * ori_forget_loss).backward()
                # grad *= mask ; clip_grad_norm(max_norm) ; optimizer.step()       sfron.py:201-206
                mhp.forget_step(use_mask=bool(self.mask), max_norm=self.max_norm, lr=lr)
                run_forget = run_forget + ori_forget_loss.detach()      # no host sync per step
                log_forget += 1
            self.model.train()
            x_retain, y_retain = next(retain_train_iter)
            x_retain, y_retain = x_retain.to(dev), y_retain.to(dev)
            ori_remain_loss = self.retain_loss_function(self.model(x_retain), y_retain)
            ori_remain_loss.backward()
            # optimizer.step() ; update_parameters(model, ori_model) ; ori_model = deepcopy(model)  :222,255-257
            mhp.remain_step(lr=lr, ema=self.ema_enabled)
            run_remain = run_remain + ori_remain_loss.detach()
            log_remain += 1
            if self.eval and (step + 1) % self.log_freq == 0:
                dt = time.time() - start_time
                print(f"step={step + 1} Forget L:{float(run_forget) / max(log_forget, 1):.4f} "
                      f"Remain L:{float(run_remain) / max(log_remain, 1):.4f} LR:{lr} steps/s:{log_remain / dt:.2f}")
                if self.validate_fn is not None:
                    self.validate_fn(self.model)
                log_forget = log_remain = 0
                run_forget = run_remain = 0.0
                start_time = time.time()
            dummy.step()          # no-op (the dummy has no gradient); keeps torch's step-order check quiet
            scheduler.step()
        return self.model

    def _unlearn_graphed(self, mhp, rates, alpha_scheduler, forget_iter, retain_iter) -> nn.Module:
        """The loop of get_unlearned_model with every iteration replayed from a CUDA graph: one graph for the
        iterations that start with a forget step (step % forget_freq == 0), one for the others.  Batches live in
        static tensors refilled before each replay; alpha_t is a device scalar; the learning rate of the iteration is
        read on the device from the scheduler's table and the optimizer step count lives on the device too.  A batch
        of another shape (a ragged last batch) runs that iteration eagerly — same kernels, same device-side state."""
        hp, flat, dev = mhp.hp, mhp.flat, mhp.flat.device
        if not flat.grads_as_views:
            raise RuntimeError("cuda_graph=True needs view-gradients (FlatParams(grads_as_views=True))")
        hp.set_lr_schedule(rates)
        hp.enable_graph_replay()
        alpha_dev = torch.zeros((), dtype=torch.float32, device=dev)
        xf, yf = (t.to(dev).clone() for t in next(forget_iter))
        xr, yr = (t.to(dev).clone() for t in next(retain_iter))
        pending = {"forget": (xf.clone(), yf.clone()), "retain": (xr.clone(), yr.clone())}   # already drawn: used first

        def forget_part():
            self.model.train()
            (alpha_dev * -self.forget_loss_function(self.model(xf), yf)).backward()
            mhp.forget_step(use_mask=bool(self.mask), max_norm=self.max_norm)            # sfron.py:201-206

        def retain_part():
            self.model.train()
            self.retain_loss_function(self.model(xr), yr).backward()
            mhp.remain_step(ema=self.ema_enabled)                                        # :222,255-257
            hp.advance_lr()                                                              # scheduler.step() :259

        bodies = {True: lambda: (forget_part(), retain_part()), False: retain_part}
        # one eager pass per body outside the capture (cuDNN / cuBLAS choose algorithms there), state put back after
        roles = [r for r in ("m", "slow") if hp.has(r)] + (["v"] if hp.has("v") else [])
        saved = {r: hp.buffer(r).clone() for r in roles}
        saved_p, saved_step, saved_idx = flat.p.clone(), hp.step_dev.clone(), hp.lr_index.clone()
        saved_buffers = {k: b.clone() for k, b in self.model.named_buffers()}
        saved_count = hp.step_count
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for with_forget in ((True, False) if self.forget_freq > 1 else (True,)):
                bodies[with_forget]()
        torch.cuda.current_stream(dev).wait_stream(side)

        def restore():
            for r in roles:
                hp.buffer(r).copy_(saved[r])
            flat.p.copy_(saved_p)
            hp.step_dev.copy_(saved_step)
            hp.lr_index.copy_(saved_idx)
            for k, b in self.model.named_buffers():
                b.copy_(saved_buffers[k])
            hp.step_count = saved_count
            mhp.zero_grad()

        restore()
        graphs = {}
        for with_forget in ((True, False) if self.forget_freq > 1 else (True,)):
            graphs[with_forget] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graphs[with_forget]):
                bodies[with_forget]()
        restore()                                          # capture records only, but host-side counters moved
        for step in range(self.n_iters):
            with_forget = step % self.forget_freq == 0
            static_ok = True
            if with_forget:
                alpha_dev.fill_(alpha_scheduler(self.forget_alpha, step, self.n_iters))
                bx, by = pending.pop("forget", None) or next(forget_iter)
                static_ok &= tuple(bx.shape) == tuple(xf.shape)
                if static_ok:
                    xf.copy_(bx, non_blocking=True)
                    yf.copy_(by, non_blocking=True)
            rx, ry = pending.pop("retain", None) or next(retain_iter)
            static_ok &= tuple(rx.shape) == tuple(xr.shape)
            if static_ok:
                xr.copy_(rx, non_blocking=True)
                yr.copy_(ry, non_blocking=True)
                graphs[with_forget].replay()
            else:                                          # ragged batch: the same iteration, launched eagerly
                self.model.train()
                if with_forget:
                    bx, by = bx.to(dev), by.to(dev)
                    (alpha_dev * -self.forget_loss_function(self.model(bx), by)).backward()
                    mhp.forget_step(use_mask=bool(self.mask), max_norm=self.max_norm)
                rx, ry = rx.to(dev), ry.to(dev)
                self.retain_loss_function(self.model(rx), ry).backward()
                mhp.remain_step(ema=self.ema_enabled)
                hp.advance_lr()
            if self.eval and (step + 1) % self.log_freq == 0 and self.validate_fn is not None:
                self.validate_fn(self.model)
        hp.step_count = int(hp.step_dev)
        return self.model

    def get_params(self) -> dict:
        self.params = {
            "opt": self.opt, "momentum": self.momentum, "weight_decay": self.weight_decay,
            "retain_lr": self.retain_lr, "n_iters": self.n_iters, "forget_freq": self.forget_freq,
            "forget_alpha": self.forget_alpha, "max_norm": self.max_norm, "ema_beta": self.ema_beta,
            "sched": self.sched, "lambd": self.lambd, "mask": self.mask, "threshold": self.th,
        }
        return self.params


class SalUn(UnlearnMethod):
    """The saliency-mask half of Classification/unlearn/salun.py (get_gradient_ratio, :140-195): the
    global top-k selection is on the hot path; SalUn's random-label fine-tuning loop is a comparison
    baseline and stays out of scope."""

    def __init__(self, model, loss_function, save_path, args) -> None:
        super().__init__(model, loss_function, save_path, args)
        self.th = 0.2
        self.mask = None

    def prepare_unlearn(self, unlearn_dataloaders: dict) -> None:
        self.unlearn_dataloaders = unlearn_dataloaders
        self.mask = self.get_gradient_ratio(unlearn_dataloaders["forget_train"])
        print(f"SalUn threshold: {self.th}")

    def get_gradient_ratio(self, forget_loader) -> Dict[str, torch.Tensor]:
        """Sum of the batch gradients of -loss, |.|, global top-`th` fraction -> int64 0/1 masks."""
        mhp = ModelHotPath(self.model, OptConfig(kind="sgd", lr=0.0), device=_device_of(self.model))
        dev = mhp.flat.device
        self.model.eval()
        mhp.zero_grad()
        for image, target in forget_loader:           # view-grads: backward accumulates the sum in place
            image, target = image.to(dev), target.to(dev)
            (-self.loss_function(self.model(image), target)).backward()
        k = int(mhp.layout.numel * self.th)
        mask = mhp.hp.topk_mask(mhp.grads(), k)
        out = formats.topk_mask_to_dict(mhp.layout, mask, all_names=mhp.flat.all_names)
        mhp.zero_grad()
        return out


_REGISTRY = {"SFRon": SFRon, "SalUn": SalUn}


def create_unlearn_method(unlearn_name):
    """Classification/unlearn/__init__.py:11-12 (an `eval` there; a table of the in-scope methods here)."""
    try:
        return _REGISTRY[unlearn_name]
    except KeyError:
        raise NotImplementedError(
            f"{unlearn_name}: only the SFR-on hot path ({sorted(_REGISTRY)}) is rebuilt; the other "
            "baselines of the reference are plain torch training loops and out of scope") from None
