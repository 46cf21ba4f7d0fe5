"""Drop-in for the reference's DDPM runner: `Diffusion(args, config)` with the hot-path modes of
DDPM/train.py:145-168 (`--mode generate_fisher | generate_mask | sfron | salun`) and DDPM/fim.py:84-89.

    runner = Diffusion(args, config)          # DDPM/runners/diffusion.py:69
    runner.generate_fisher()                  # :1210-1364  -> {ckpt_folder}/mask_{label}/{forget,remain}_fisher.pt
    runner.generate_mask()                    # :930-1036   -> results/cifar10/mask/{label}/with_0.5.pt   (cwd-relative)
    runner.save_fim()                         # :262-352    -> {ckpt_folder}/fisher_dict.pkl
    runner.sfron_forget()                     # :1038-1208  -> {config.ckpt_dir}/ckpt.pth  [model, opt, step, ema]
    runner.saliency_unlearn()                 # :479-616    -> {config.ckpt_dir}/ckpt.pth

Same `args` fields (ckpt_folder, label_to_forget, cond_scale, mask_path, forget_alpha, decay_forget_alpha,
remain_alpha, method, unlearn_loss, n_chunks), same `config` fields (diffusion.*, optim.*, model.ema / ema_rate /
type, training.n_iters / log_freq / snapshot_freq / save_freq / lambd, data.num_workers), same files, keys
(`module.` prefix of the DataParallel wrapper) and dtypes.  What the path does NOT own — the network, the datasets,
the sampler — comes in through `DDPMHooks`; by default they are imported from the reference's own modules
(`models.diffusion`, `dataset`, run from inside its DDPM/ directory), so that
`from sfron_b200.methods.ddpm import Diffusion` replaces `from runners.diffusion import Diffusion`.

Everything between `loss.backward()` and the next forward runs in the CUDA kernels (DiffusionUnlearner):
no per-parameter Python loop, no `.cpu()` of gradients, no mask upload per step.
Not reproduced: `--method joint` of sfron_forget (:1160-1167), which multiplies the PREVIOUS step's gradients by
the mask before `zero_grad()` — i.e. applies no mask at all; it raises instead of silently doing something else.
"""
from __future__ import annotations

import logging
import os
import time
from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import numpy as np
import torch

from .. import capi, formats
from ..engine import OptConfig
from .common import cosine_lr_scheduler, cycle
from .diffusion import DiffusionUnlearner, adaptive_loss


# ---- what the runner needs from outside the path ----------------------------------------------------
@dataclass
class DDPMHooks:
    model_factory: Callable                      # config -> nn.Module                 (models.diffusion.Conditional_Model)
    forget_dataset: Callable                     # (args, config, label) -> (remain_loader, forget_loader)
    data_transform: Callable                     # (config, x) -> x                    (dataset.data_transform)
    fim_loader: Optional[Callable] = None        # (args, config, batch_size) -> DataLoader over class_samples (save_fim)
    sample_visualization: Optional[Callable] = None   # (runner, model, step, cond_scale) after a snapshot


def reference_hooks() -> DDPMHooks:
    """The reference's own model / data code (needs its DDPM/ directory on sys.path)."""
    try:
        from models.diffusion import Conditional_Model
        from dataset import data_transform, get_forget_dataset
    except Exception as e:                                       # pragma: no cover - depends on the deployment
        raise ImportError("Diffusion(args, config) without `hooks=` imports the reference's models.diffusion and "
                          "dataset modules: run it from the reference's DDPM/ directory or pass DDPMHooks") from e

    def fim_loader(args, config, bs):
        from torch.utils.data import DataLoader
        from torchvision import transforms
        from torchvision.datasets import ImageFolder
        ds = ImageFolder(os.path.join(args.ckpt_folder, "class_samples"), transform=transforms.ToTensor())
        return DataLoader(ds, batch_size=bs, num_workers=config.data.num_workers, shuffle=True)

    return DDPMHooks(Conditional_Model, get_forget_dataset, data_transform, fim_loader)


def beta_schedule(kind: str, beta_start: float, beta_end: float, steps: int) -> np.ndarray:
    """The schedules DDPM/runners/diffusion.py:36-66 offers (float64, as there)."""
    if kind == "quad":
        return np.linspace(beta_start ** 0.5, beta_end ** 0.5, steps, dtype=np.float64) ** 2
    if kind == "linear":
        return np.linspace(beta_start, beta_end, steps, dtype=np.float64)
    if kind == "const":
        return beta_end * np.ones(steps, dtype=np.float64)
    if kind == "jsd":
        return 1.0 / np.linspace(steps, 1, steps, dtype=np.float64)
    if kind == "sigmoid":
        x = np.linspace(-6, 6, steps)
        return 1 / (np.exp(-x) + 1) * (beta_end - beta_start) + beta_start
    raise NotImplementedError(kind)


def eps_loss(model, x0, t, c, e, b, cond_drop_prob=0.1, keepdim=False):
    """Conditional noise-estimation loss (DDPM/functions/losses.py:21-37)."""
    a = (1 - b).cumprod(dim=0).index_select(0, t).view(-1, 1, 1, 1)
    x = x0 * a.sqrt() + e * (1.0 - a).sqrt()
    out = model(x, t.float(), c, cond_drop_prob=cond_drop_prob, mode="train")
    per = (e - out).square().sum(dim=(1, 2, 3))
    return per if keepdim else per.mean(dim=0)


LOSSES = {"simple": eps_loss}                                     # loss_registry_conditional


class Diffusion:
    def __init__(self, args, config, hooks: Optional[DDPMHooks] = None, device=None):
        self.args, self.config = args, config
        self.hooks = hooks or reference_hooks()
        if device is None:
            if not torch.cuda.is_available():
                raise capi.SfrError(capi.ERR_NO_DEVICE, "Diffusion", "the SFR-on hot path runs on CUDA only")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        d = config.diffusion
        betas = beta_schedule(d.beta_schedule, d.beta_start, d.beta_end, d.num_diffusion_timesteps)
        self.betas = torch.from_numpy(betas).float().to(self.device)
        self.num_timesteps = self.betas.shape[0]
        # test / measurement tap: called with (kind, flat gradient) right before the kernels consume it
        self.gradient_tap: Optional[Callable[[str, torch.Tensor], None]] = None

    # ---- shared pieces ---------------------------------------------------------------------------------
    def _load(self) -> Tuple[torch.nn.Module, list]:
        """Conditional_Model(config) + states[0] of {ckpt_folder}/ckpts/ckpt.pth (keys carry DataParallel's prefix)."""
        print("Loading checkpoints {}".format(self.args.ckpt_folder))
        model = self.hooks.model_factory(self.config).to(self.device)
        states = torch.load(os.path.join(self.args.ckpt_folder, "ckpts/ckpt.pth"), map_location=self.device,
                            weights_only=False)
        model.load_state_dict({k[len("module."):] if k.startswith("module.") else k: v for k, v in states[0].items()},
                              strict=True)
        return model, states

    def _unlearner(self, model, states=None) -> DiffusionUnlearner:
        o, cfg = self.config.optim, self.config
        opt = OptConfig("adam", lr=o.lr, beta1=o.beta1, beta2=0.999, eps=o.eps, weight_decay=o.weight_decay)
        if getattr(o, "optimizer", "Adam") != "Adam" or getattr(o, "amsgrad", False):
            raise NotImplementedError("the path implements the optimizer of the reference's configs: Adam, amsgrad off")
        ema = bool(cfg.model.ema)
        un = DiffusionUnlearner(model, "ddpm", device=self.device, opt=opt, ema_mode="ddpm" if ema else "none",
                                ema_a=float(cfg.model.ema_rate) if ema else 0.0, clip_forget=o.grad_clip,
                                clip_remain=o.grad_clip, clip_fisher=o.grad_clip)
        if ema and states is not None:                       # ema_helper.register(model); load_state_dict(states[-1])
            shadow = formats.dict_to_flat(un.mhp.layout, states[-1], device=self.device)
            un.mhp.hp.slow.copy_(shadow)
        return un

    def _t(self, n: int) -> torch.Tensor:
        """antithetic timestep sampling (:1252-1255 and every loop of the runner)."""
        t = torch.randint(low=0, high=self.num_timesteps, size=(n // 2 + 1,)).to(self.device)
        return torch.cat([t, self.num_timesteps - t - 1], dim=0)[:n]

    def _tap(self, kind: str, un: DiffusionUnlearner) -> None:
        if self.gradient_tap is not None:
            self.gradient_tap(kind, un.mhp.grads())

    def _test_mode_loss(self, model, x, c):
        """`mode="test"` classifier-free-guided prediction loss of the Fisher / mask generators (:1243-1265)."""
        x = self.hooks.data_transform(self.config, x.to(self.device))
        c = c.to(self.device)
        e = torch.randn_like(x)
        t = self._t(x.size(0))
        a = (1 - self.betas).cumprod(dim=0).index_select(0, t).view(-1, 1, 1, 1)
        x = x * a.sqrt() + e * (1.0 - a).sqrt()
        out = model(x, t.float(), c, cond_scale=self.args.cond_scale, mode="test")
        return (e - out).square().sum(dim=(1, 2, 3)).mean(dim=0)

    def _snapshot(self, un: DiffusionUnlearner, step: int) -> None:
        torch.save(un.checkpoint(step), os.path.join(self.config.ckpt_dir, "ckpt.pth"))
        if self.hooks.sample_visualization is not None:
            self.hooks.sample_visualization(self, un.model, step, self.args.cond_scale)

    # ---- --mode generate_fisher ------------------------------------------------------------------------
    def generate_fisher(self):
        args, config = self.args, self.config
        logging.info("Generating fisher of diffusion to achieve gradient sparsity.")
        remain_loader, forget_loader = self.hooks.forget_dataset(args, config, args.label_to_forget)
        model, _ = self._load()
        mask_path = os.path.join(args.ckpt_folder, f"mask_{args.label_to_forget}")
        os.makedirs(mask_path, exist_ok=True)
        un = self._unlearner(model)
        model.eval()
        for which, loader in (("forget", forget_loader), ("remain", remain_loader)):
            batches = iter(loader)
            running, log_steps, start = 0.0, 0, time.time()

            def loss_fn(i):
                x, c = next(batches)
                loss = self._test_mode_loss(model, x, c)
                nonlocal running, log_steps, start
                running, log_steps = running + float(loss.detach()), log_steps + 1
                if (i + 1) % config.training.log_freq == 0:
                    logging.info(f"{which.capitalize()} (step={i + 1:07d}) Loss: {running / log_steps:.4f} "
                                 f"Train Steps/Sec: {log_steps / (time.time() - start):.2f}")
                    running, log_steps, start = 0.0, 0, time.time()
                return _Tapped(loss, lambda: self._tap(which, un))

            # F += clip(grad)**2 / len(loader), on the device; written as {forget,remain}_fisher.pt (:1277-1299)
            un.generate_fisher(which, len(loader), loss_fn, mask_path)

    # ---- --mode generate_mask (SalUn top-k) ------------------------------------------------------------
    def generate_mask(self):
        args, config = self.args, self.config
        logging.info(f"Generating mask of diffusion to achieve gradient sparsity. Gamma: {config.training.gamma}, "
                     f"lambda: {config.training.lmbda}")
        _, forget_loader = self.hooks.forget_dataset(args, config, args.label_to_forget)
        model, _ = self._load()
        un = self._unlearner(model)
        model.eval()
        batches = iter(forget_loader)

        def loss_fn(i):
            x, c = next(batches)
            return _Tapped(self._test_mode_loss(model, x, c), lambda: self._tap("forget", un))

        mask_dir = os.path.join("results/cifar10/mask", str(args.label_to_forget))      # cwd-relative, as :1000
        for ratio in [0.5]:
            print(ratio)
            un.generate_topk_mask(len(forget_loader), loss_fn, ratio, os.path.join(mask_dir, f"with_{ratio}.pt"))

    # ---- DDPM/fim.py: per-sample FIM -------------------------------------------------------------------
    def save_fim(self, batch_size: Optional[int] = None):
        """F += (sum over all timesteps of the per-sample gradient)**2 / |D| (:262-352).  The reference processes
        one sample per GPU of its DataParallel wrapper (bs = torch.cuda.device_count()); one process drives one GPU
        here, so `batch_size` defaults to that count and the per-sample rows go through K1's row form."""
        args, config = self.args, self.config
        bs = batch_size or max(1, torch.cuda.device_count())
        if self.hooks.fim_loader is None:
            raise ValueError("save_fim needs DDPMHooks.fim_loader")
        loader = self.hooks.fim_loader(args, config, bs)
        model, _ = self._load()
        model.eval()
        un = self._unlearner(model)
        flat = un.mhp.flat
        stride = (flat.n + 7) // 8 * 8
        rows = torch.zeros(bs, stride, dtype=torch.float32, device=self.device)
        out = os.path.join(args.ckpt_folder, "fisher_dict.pkl")
        n_data = len(loader.dataset)

        def blocks():
            for step, (x, c) in enumerate(loader):
                x, c = x.to(self.device), c.to(self.device)
                rows.zero_()
                for chunk in torch.chunk(torch.arange(0, self.num_timesteps), args.n_chunks):
                    loss = 0
                    for ti in chunk:
                        e = torch.randn_like(x)
                        t = torch.tensor([int(ti)]).expand(x.size(0)).to(self.device)
                        loss = loss + LOSSES[config.model.type](model, x, t, c, e, self.betas, keepdim=True)
                    for i in range(x.size(0)):              # first-order gradient of every sample, separately
                        un.mhp.zero_grad()
                        loss[i].backward(retain_graph=i != x.size(0) - 1)
                        rows[i, :flat.n] += un.mhp.grads()
                    del loss
                un.mhp.zero_grad()
                yield rows[:x.size(0), :flat.n]
                if (step + 1) % config.training.save_freq == 0:
                    formats.save_fim_pickle(out, un.mhp.layout, un.mhp.hp.buffer("fim"), prefix="module.")

        un.save_fim(blocks(), n_data, out)

    # ---- --mode sfron ----------------------------------------------------------------------------------
    def sfron_forget(self):
        args, config = self.args, self.config
        if args.method != "ron":
            raise NotImplementedError(f"--method {args.method}: only 'ron' is on the path (see the module docstring)")
        remain_loader, forget_loader = self.hooks.forget_dataset(args, config, args.label_to_forget)
        remain_iter, forget_iter = cycle(remain_loader), cycle(forget_loader)
        model, states = self._load()
        un = self._unlearner(model, states)
        use_mask = bool(args.mask_path)
        if use_mask:
            un.load_mask(args.mask_path)                    # uploaded ONCE (the reference re-uploads it every step)
        loss_fn = LOSSES[config.model.type]
        criteria = torch.nn.MSELoss()
        n_iters = config.training.n_iters
        mhp = un.mhp
        mhp.zero_grad()
        model.train()
        start = time.time()
        for step in range(n_iters):
            alpha = cosine_lr_scheduler(args.forget_alpha, step, n_iters) if args.decay_forget_alpha else args.forget_alpha
            # forget stage (:1080-1138)
            x, c = next(forget_iter)
            x = self.hooks.data_transform(config, x.to(self.device))
            c = c.to(self.device)
            e = torch.randn_like(x)
            t = self._t(x.size(0))
            if args.unlearn_loss == "ga":
                ori_forget = -loss_fn(model, x, t, c, e, self.betas)
            elif args.unlearn_loss == "rl":
                a = (1 - self.betas).cumprod(dim=0).index_select(0, t).view(-1, 1, 1, 1)
                xt = x * a.sqrt() + e * (1.0 - a).sqrt()
                out = model(xt, t.float(), c, mode="train")
                pseudo_c = torch.full(c.shape, (args.label_to_forget + 1) % 10, device=c.device)
                ori_forget = criteria(model(xt, t.float(), pseudo_c, mode="train").detach(), out)
            elif args.unlearn_loss == "adaga":
                per = loss_fn(model, x, t, c, e, self.betas, 0.1, keepdim=True)
                ori_forget = -adaptive_loss(per, x.shape[0], gamma=config.training.lambd, eps=1e-8)
            else:
                raise NotImplementedError
            (alpha * ori_forget).backward()
            self._tap("forget", un)
            mhp.forget_step(use_mask=use_mask, max_norm=config.optim.grad_clip)     # mask, clip, Adam — one pass
            # remain stage (:1140-1180)
            x, c = next(remain_iter)
            x = self.hooks.data_transform(config, x.to(self.device))
            c = c.to(self.device)
            e = torch.randn_like(x)
            t = self._t(x.size(0))
            ori_remain = loss_fn(model, x, t, c, e, self.betas)
            (args.remain_alpha * ori_remain).backward()
            self._tap("remain", un)
            mhp.remain_step(max_norm=config.optim.grad_clip, ema=bool(config.model.ema))    # clip, Adam, EMA — one pass
            if (step + 1) % config.training.log_freq == 0:
                end = time.time()
                logging.info(f"step:{step:04d}, remain L:{ori_remain.item():.4f}, remain a:{args.remain_alpha}, "
                             f"forget L:{ori_forget.item():.4f}, forget a:{alpha:.8f}, time:{end - start:.2f}")
                start = time.time()
            if (step + 1) % config.training.snapshot_freq == 0:
                self._snapshot(un, step)

    # ---- --mode salun ----------------------------------------------------------------------------------
    def saliency_unlearn(self):
        args, config = self.args, self.config
        remain_loader, forget_loader = self.hooks.forget_dataset(args, config, args.label_to_forget)
        remain_iter, forget_iter = cycle(remain_loader), cycle(forget_loader)
        model, states = self._load()
        un = self._unlearner(model, states)
        use_mask = bool(args.mask_path)
        if use_mask:
            un.load_mask(args.mask_path)
        loss_fn = LOSSES[config.model.type]
        criteria = torch.nn.MSELoss()
        mhp = un.mhp
        mhp.zero_grad()
        model.train()
        start = time.time()
        for step in range(config.training.n_iters):
            x, c = next(remain_iter)
            x = self.hooks.data_transform(config, x.to(self.device))
            c = c.to(self.device)
            e = torch.randn_like(x)
            remain_loss = loss_fn(model, x, self._t(x.size(0)), c, e, self.betas)
            x, c = next(forget_iter)
            x = self.hooks.data_transform(config, x.to(self.device))
            c = c.to(self.device)
            e = torch.randn_like(x)
            t = self._t(x.size(0))
            if args.unlearn_loss == "ga":
                forget_loss = -loss_fn(model, x, t, c, e, self.betas)
            elif args.unlearn_loss == "rl":
                a = (1 - self.betas).cumprod(dim=0).index_select(0, t).view(-1, 1, 1, 1)
                xt = x * a.sqrt() + e * (1.0 - a).sqrt()
                out = model(xt, t.float(), c, mode="train")
                pseudo_c = torch.full(c.shape, (args.label_to_forget + 1) % 10, device=c.device)
                forget_loss = criteria(model(xt, t.float(), pseudo_c, mode="train").detach(), out)
            else:
                raise NotImplementedError
            loss = args.forget_alpha * forget_loss + args.remain_alpha * remain_loss
            if (step + 1) % config.training.log_freq == 0:
                end = time.time()
                logging.info(f"step: {step}, loss: {loss.item()}, time: {end - start}")
                start = time.time()
            loss.backward()
            self._tap("joint", un)
            # clip the UNMASKED gradient, then mask, Adam, EMA (:576-593) — one pass
            mhp.joint_step(use_mask=use_mask, max_norm=config.optim.grad_clip, mask_order="clip_then_mask",
                           ema=bool(config.model.ema))
            if (step + 1) % config.training.snapshot_freq == 0:
                self._snapshot(un, step)


class _Tapped:
    """A loss whose `.backward()` also fires the runner's gradient tap (the generators hand their losses to
    DiffusionUnlearner, which calls backward itself)."""

    def __init__(self, loss, after):
        self.loss, self.after = loss, after

    def backward(self):
        self.loss.backward()
        self.after()


def main(argv=None):
    """`python -m sfron_b200.methods.ddpm --mode ...` from the reference's DDPM/ directory: its own argument and
    config parsing (train.py:21-137), this runner in place of runners.diffusion.Diffusion."""
    import sys
    import traceback
    sys.argv = [sys.argv[0]] + list(argv if argv is not None else sys.argv[1:])
    from train import parse_args_and_config                  # the reference's parser (needs DDPM/ on sys.path)
    args, config = parse_args_and_config()
    modes = {"sfron": "sfron_forget", "salun": "saliency_unlearn", "generate_mask": "generate_mask",
             "generate_fisher": "generate_fisher"}
    if args.mode not in modes:
        raise SystemExit(f"--mode {args.mode} is outside the SFR-on hot path; use the reference's train.py for it")
    try:
        getattr(Diffusion(args, config), modes[args.mode])()
    except Exception:
        logging.error(traceback.format_exc())
        return 1
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
