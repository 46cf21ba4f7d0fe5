"""Hot-path halves of the diffusion entry points, over caller-supplied loss closures.

The reference's diffusion scripts interleave the path with dataset / VAE / sampler code that is
out of scope.  `DiffusionUnlearner` keeps the path itself — the order of
operations, hyper-parameter defaults, file names and dict formats of each family — and takes the
model plus closures that produce the (already alpha-weighted) losses:

  family "ddpm"  Diffusion.generate_fisher / generate_mask / save_fim / sfron_forget / saliency_unlearn
                 DDPM/runners/diffusion.py:1210-1364, 930-1036, 262-352, 1038-1208, 479-616
  family "dit"   DiT/generate_fisher.py:216-291, DiT/forget.py:256-355
  family "sd"    SD/train-scripts/generate_fisher.py:31-129, nsfw_removal.py:108-173
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Dict, Iterable, Optional

import torch

from .. import capi, formats
from ..engine import OptConfig
from .common import ModelHotPath, cosine_lr_scheduler


@dataclass
class FamilyPreset:
    opt: OptConfig
    ema_mode: str
    ema_a: float
    key_prefix: str            # prefix of the reference's file keys (DataParallel's "module.")
    clip_forget: Optional[float]
    clip_remain: Optional[float]
    clip_fisher: Optional[float]
    forget_fisher_name: str = "forget_fisher.pt"
    remain_fisher_name: str = "remain_fisher.pt"


def preset(family: str, **over) -> FamilyPreset:
    if family == "ddpm":      # configs/cifar10_sfron.yml: Adam lr 1e-4 beta1 .9 eps 1e-8 wd 0, grad_clip 1.0, ema_rate 1e-4
        p = FamilyPreset(OptConfig("adam", lr=over.pop("lr", 1e-4), beta1=0.9, beta2=0.999, eps=1e-8,
                                   weight_decay=0.0), "ddpm", 1e-4, "module.", 1.0, 1.0, 1.0)
    elif family == "dit":     # DiT/forget.py:199 AdamW(lr, wd=0); --grad-clip 1.0 on the forget step only; EMA 0.9999
        p = FamilyPreset(OptConfig("adamw", lr=over.pop("lr", 1e-4), weight_decay=0.0), "dit", 0.9999,
                         "module.", 1.0, None, None)
    elif family == "sd":      # nsfw_removal.py:81 Adam(lr); no clip, no EMA; U-Net-local key names
        p = FamilyPreset(OptConfig("adam", lr=over.pop("lr", 1e-5)), "none", 0.0, "", None, None, None,
                         forget_fisher_name="nude_forget.pt", remain_fisher_name="nude_remain.pt")
    else:
        raise ValueError(family)
    for k, v in over.items():
        setattr(p, k, v)
    return p


LossFn = Callable[[int], torch.Tensor]


def adaptive_loss(ori_loss: torch.Tensor, size: int, gamma: float = 1.0, eps: float = 1e-15,
                  keepdim: bool = False) -> torch.Tensor:
    """Adaptive forget-loss re-weighting: `coef = 1 / (loss**gamma + eps)`, normalised, times the batch
    size.  Stays in PyTorch (it is part of the loss graph).  eps differs per sub-project:
    1e-15 Classification sfron.py:57 and DiT/forget.py:44; 1e-8 DDPM/functions/losses.py:63."""
    coef = 1 / (torch.pow(ori_loss.detach().clone(), gamma) + eps)
    ad_loss = (coef / coef.sum()) * ori_loss * size
    return ad_loss if keepdim else ad_loss.mean(dim=0)


def per_sample_grad_rows(model: torch.nn.Module, per_sample_loss: Callable, batch: tuple) -> torch.Tensor:
    """[B, n] flat per-sample gradients in ONE vmapped backward pass — the producer for K1's `rows > 1`
    form.  Replaces DDPM save_fim's `for i in range(bs): loss[i].backward(retain_graph=True)` loop
    (runners/diffusion.py:326-333), which runs B backward passes over a retained graph.

    per_sample_loss(params_and_buffers: dict, *sample) -> scalar loss of ONE sample, written with
    torch.func.functional_call.  Rows follow the flat layout (trainable parameters, named_parameters order).
    """
    from torch.func import grad, vmap
    params = {n: p.detach() for n, p in model.named_parameters() if p.requires_grad}
    frozen = {n: p.detach() for n, p in model.named_parameters() if not p.requires_grad}
    buffers = {n: b for n, b in model.named_buffers()}

    def one(train_params, *sample):
        return per_sample_loss({**train_params, **frozen, **buffers}, *sample)

    grads = vmap(grad(one), in_dims=(None,) + (0,) * len(batch))(params, *batch)
    b = batch[0].shape[0]
    n = sum(p.numel() for p in params.values())
    stride = (n + 7) // 8 * 8                       # every row 16-byte aligned (fp32 and bf16)
    rows = torch.empty(b, stride, dtype=next(iter(grads.values())).dtype, device=batch[0].device)
    off = 0
    for name in params:
        g = grads[name].reshape(b, -1)
        rows[:, off:off + g.shape[1]].copy_(g)
        off += g.shape[1]
    return rows[:, :n]


class DiffusionUnlearner:
    def __init__(self, model: torch.nn.Module, family: str = "dit", *, device=None, **preset_overrides):
        self.family = family
        self.cfg = preset(family, **preset_overrides)
        self.model = model
        self.mhp = ModelHotPath(model, self.cfg.opt, ema_mode=self.cfg.ema_mode, ema_a=self.cfg.ema_a,
                                key_prefix=self.cfg.key_prefix, device=device)

    # ---- Fisher ------------------------------------------------------------------------------------
    def generate_fisher(self, which: str, n_batches: int, loss_fn: LossFn, out_dir: Optional[str] = None, *,
                        cuda_graph: bool = False, refill: Optional[Callable[[int], None]] = None) -> None:
        """`F += grad**2 / n_batches` over `n_batches` backward passes of `loss_fn(i)`; written in the
        reference's dict format to {out_dir}/{forget,remain}_fisher.pt (nude_*.pt for SD).

        cuda_graph=True: forward + backward + Fisher kernel of ONE batch captured once and replayed n_batches
        times (the reference runs this loop for 2000 batches of one sample, DiT/generate_fisher.py:220-239);
        `loss_fn` must then read its batch from static device tensors that `refill(i)` overwrites."""
        mhp = self.mhp
        acc = mhp.hp.buffer(f"{which}_fisher")
        acc.zero_()
        if cuda_graph:
            if not mhp.flat.grads_as_views:
                raise RuntimeError("cuda_graph=True needs view-gradients (FlatParams(grads_as_views=True))")
            dev = mhp.flat.device

            def body():
                loss_fn(0).backward()
                mhp.fisher_accumulate(which, float(n_batches), clip_max_norm=self.cfg.clip_fisher)
                mhp.zero_grad()

            if refill is not None:
                refill(0)
            mhp.zero_grad()
            side = torch.cuda.Stream(device=dev)          # one eager pass outside the capture, then undone
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                body()
            torch.cuda.current_stream(dev).wait_stream(side)
            acc.zero_()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                body()
            for i in range(n_batches):
                if refill is not None:
                    refill(i)
                graph.replay()
        else:
            for i in range(n_batches):
                mhp.zero_grad()
                loss_fn(i).backward()
                mhp.fisher_accumulate(which, float(n_batches), clip_max_norm=self.cfg.clip_fisher)
            mhp.zero_grad()
        if out_dir is not None:
            os.makedirs(out_dir, exist_ok=True)
            name = self.cfg.forget_fisher_name if which == "forget" else self.cfg.remain_fisher_name
            mhp.save_fisher(which, os.path.join(out_dir, name))

    def save_fim(self, per_sample_grads: Iterable[torch.Tensor], dataset_len: int, path: Optional[str] = None):
        """Per-sample FIM of DDPM save_fim: every item is a [B, n] (or [n]) block of per-sample
        gradients summed over the timesteps; F += row**2 / |D| row by row (runners/diffusion.py:337-344)."""
        acc = self.mhp.hp.buffer("fim")
        acc.zero_()
        for rows in per_sample_grads:
            capi.fisher_accum(acc, rows, float(dataset_len))
        if path is not None:
            formats.save_fim_pickle(path, self.mhp.layout, acc, prefix=self.cfg.key_prefix)
        return acc

    # ---- masks -------------------------------------------------------------------------------------
    def ratio_mask(self, threshold: float, out_dir: Optional[str] = None, name_fmt: str = "fisher_{th}.pt"):
        mask = self.mhp.ratio_mask(threshold)
        if out_dir is not None:
            torch.save(mask, os.path.join(out_dir, name_fmt.format(th=formats.threshold_tag(threshold))))
        return mask

    def generate_topk_mask(self, n_batches: int, loss_fn: LossFn, ratio: float = 0.5,
                           path: Optional[str] = None) -> Dict[str, object]:
        """SalUn mask of DDPM generate_mask: sum of the (clipped) batch gradients, |.|, global
        top-`ratio` -> int64 0/1 (runners/diffusion.py:955-1036)."""
        mhp = self.mhp
        acc = mhp.hp.buffer("grad_sum")
        acc.zero_()
        for i in range(n_batches):
            mhp.zero_grad()
            loss_fn(i).backward()
            # clip_grad_norm_ ; gradients[name] += grad  (:985-994) — norm pass + one fused accumulate pass
            mhp.hp.saliency_accumulate(mhp.grads(), clip_max_norm=self.cfg.clip_fisher)
        mhp.zero_grad()
        k = int(mhp.layout.numel * ratio)
        mask = mhp.hp.topk_mask(acc, k)
        out = formats.topk_mask_to_dict(mhp.layout, mask, all_names=mhp.flat.all_names, prefix=self.cfg.key_prefix)
        if path is not None:
            os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
            torch.save(out, path)
        return out

    def load_mask(self, path_or_dict) -> None:
        self.mhp.load_mask(path_or_dict)

    # ---- forget loops ----------------------------------------------------------------------------------
    def forget(self, n_iters: int, forget_loss_fn: LossFn, remain_loss_fn: LossFn, *, use_mask: bool = True,
               forget_alpha: float = 1.0, remain_alpha: float = 1.0, decay_forget_alpha: bool = False,
               mask_order: str = "mask_then_clip", log_every: int = 0, cuda_graph: bool = False,
               refill: Optional[Callable[[int], None]] = None) -> None:
        """method "ron": forget step (mask, clip, step) then remain step (clip?, step) then EMA, every
        iteration (runners/diffusion.py:1075-1180; DiT/forget.py:256-322; nsfw_removal.py:108-173).
        The closures return the UNWEIGHTED losses (-loss for gradient ascent).

        cuda_graph=True captures ONE whole iteration — both PyTorch forward/backward passes and the hot-path
        kernels — in a CUDA graph and replays it n_iters times (the small-batch loops are launch-bound).
        The closures must then read their batch from STATIC device tensors and `refill(step)` is called
        before each replay to copy the next batch into them; alpha_t reaches the graph as a device scalar."""
        mhp, cfg = self.mhp, self.cfg
        if cuda_graph:
            return self._forget_graphed(n_iters, forget_loss_fn, remain_loss_fn, use_mask, forget_alpha,
                                        remain_alpha, decay_forget_alpha, mask_order, refill)
        mhp.zero_grad()
        self.model.train()
        for step in range(n_iters):
            alpha = cosine_lr_scheduler(forget_alpha, step, n_iters) if decay_forget_alpha else forget_alpha
            (alpha * forget_loss_fn(step)).backward()
            mhp.forget_step(use_mask=use_mask, max_norm=cfg.clip_forget, mask_order=mask_order)
            (remain_alpha * remain_loss_fn(step)).backward()
            mhp.remain_step(max_norm=cfg.clip_remain, ema=True)
            if log_every and (step + 1) % log_every == 0:
                print(f"step:{step:04d} forget a:{alpha:.8f}")

    def _forget_graphed(self, n_iters, forget_loss_fn, remain_loss_fn, use_mask, forget_alpha, remain_alpha,
                        decay_forget_alpha, mask_order, refill) -> None:
        mhp, cfg, hp, flat = self.mhp, self.cfg, self.mhp.hp, self.mhp.flat
        if not flat.grads_as_views:
            raise RuntimeError("cuda_graph=True needs view-gradients (FlatParams(grads_as_views=True))")
        import time
        t_capture = time.perf_counter()
        hp.enable_graph_replay()                      # optimizer step counter on the device; buffers allocated
        self.model.train()
        alpha_dev = torch.zeros((), dtype=torch.float32, device=flat.device)

        def body():
            (alpha_dev * forget_loss_fn(0)).backward()
            mhp.forget_step(use_mask=use_mask, max_norm=cfg.clip_forget, mask_order=mask_order)
            (remain_alpha * remain_loss_fn(0)).backward()
            mhp.remain_step(max_norm=cfg.clip_remain, ema=True)

        # One eager iteration on a side stream (cuDNN / cuBLAS pick their algorithms and workspaces outside the
        # capture), with the state it advances put back afterwards: the run makes exactly n_iters steps.
        roles = [r for r in ("m", "v", "slow") if hp.has(r)]
        saved = {r: hp.buffer(r).clone() for r in roles}
        saved_p = flat.p.clone()
        saved_w = None if flat.p_work is None else flat.p_work.clone()
        saved_frozen = None if mhp.frozen_slow is None else mhp.frozen_slow.clone()
        saved_step, saved_count = hp.step_dev.clone(), hp.step_count
        if refill is not None:
            refill(0)
        alpha_dev.fill_(cosine_lr_scheduler(forget_alpha, 0, n_iters) if decay_forget_alpha else forget_alpha)
        mhp.zero_grad()
        side = torch.cuda.Stream(device=flat.device)
        side.wait_stream(torch.cuda.current_stream(flat.device))
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream(flat.device).wait_stream(side)
        for r in roles:
            hp.buffer(r).copy_(saved[r])
        flat.p.copy_(saved_p)
        if saved_w is not None:
            flat.p_work.copy_(saved_w)
        if saved_frozen is not None:
            mhp.frozen_slow.copy_(saved_frozen)
        hp.step_dev.copy_(saved_step)
        hp.step_count = saved_count
        del saved, saved_p, saved_w, saved_frozen
        mhp.zero_grad()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            body()
        # capture records, it does not execute: state is untouched, but the host-side counter moved
        hp.step_count = saved_count
        torch.cuda.synchronize(flat.device)
        self.graph_capture_s = time.perf_counter() - t_capture     # one-off cost, reported by the e2e tools
        for step in range(n_iters):
            alpha_dev.fill_(cosine_lr_scheduler(forget_alpha, step, n_iters) if decay_forget_alpha else forget_alpha)
            if refill is not None:
                refill(step)
            graph.replay()
        hp.step_count = saved_count + 2 * n_iters

    def saliency_unlearn(self, n_iters: int, joint_loss_fn: LossFn, *, use_mask: bool = True,
                         log_every: int = 0) -> None:
        """SalUn loop of DDPM `--mode saliency_unlearn` (runners/diffusion.py:518-594): ONE backward per
        iteration on `forget_alpha * forget_loss + remain_alpha * remain_loss` (the closure returns that
        sum), then clip -> mask -> optimizer.step -> EMA.  Note the order: unlike SFR-on, the reference
        clips the unmasked gradient and masks afterwards."""
        mhp, cfg = self.mhp, self.cfg
        mhp.zero_grad()
        self.model.train()
        for step in range(n_iters):
            loss = joint_loss_fn(step)
            loss.backward()
            mhp.joint_step(use_mask=use_mask, max_norm=cfg.clip_remain, mask_order="clip_then_mask", ema=True)
            if log_every and (step + 1) % log_every == 0:
                print(f"step: {step}, loss: {float(loss)}")

    # ---- consumers next to the path -------------------------------------------
    def snapshot_params(self, role: str = "params_mle") -> torch.Tensor:
        """`params_mle_dict[name] = param.data.clone()` (runners/diffusion.py:391-393) /
        `gpu1_init_params` (proximal_gradient.py): a flat copy of the current weights."""
        buf = self.mhp.hp.buffer(role)
        buf.copy_(self.mhp.flat.p)
        return buf

    def add_ewc_penalty(self, lmbda: float, fisher_role: str = "fim") -> torch.Tensor:
        """Selective-Amnesia term of sa_forget: call after `loss.backward()`; adds the gradient of
        `lmbda * sum(F * (p - p_mle)**2)` into the flat gradient and returns the penalty value."""
        hp = self.mhp.hp
        return hp.ewc_penalty(self.mhp.flat.p, hp.buffer("params_mle"), hp.buffer(fisher_role), self.mhp.grads(), lmbda)

    def proximal_shrink(self, k: int, init_role: str = "params_init") -> torch.Tensor:
        """proximal_gradient.py:151-183 after `optimizer.step()`: soft-threshold theta - theta0 at its
        k-th smallest magnitude; returns the threshold (device scalar)."""
        return self.mhp.hp.proximal_shrink(self.mhp.flat.p, self.mhp.hp.buffer(init_role), k)

    # ---- checkpoints -------------------------------------------------------------------------------
    def checkpoint(self, step: int = 0, args=None):
        """DDPM: [model_sd, opt_sd, step, ema_shadow]; DiT: {model, ema, opt, args}; SD: state_dict."""
        pre = self.cfg.key_prefix
        model_sd = {pre + k: v.detach().clone() for k, v in self.model.state_dict().items()}
        if self.family == "ddpm":
            shadow = {k[len(pre):]: v for k, v in self.mhp.slow_state_dict().items()}   # EMAHelper keys are module-local
            return formats.ddpm_checkpoint(model_sd, self.mhp.optimizer_state_dict(), step, shadow)
        if self.family == "dit":
            ema_sd = {k[len(pre):]: v for k, v in self.mhp.slow_state_dict().items()}     # ema = deepcopy(model): bare names
            return formats.dit_checkpoint(model_sd, ema_sd, self.mhp.optimizer_state_dict(), args)
        return {k: v.detach().clone() for k, v in self.model.state_dict().items()}
