"""Drop-in entry points of the DiT sub-project: `forget.py` and `generate_fisher.py` with their argparse flags.

    python -m sfron_b200.methods.dit forget --data-path ... --forget-class 207 --method ron --mask-path ...
    python -m sfron_b200.methods.dit generate_fisher --data-path ... --forget-class 207 --mask-path ...

Flags, defaults, experiment-directory naming, output files and formats are the reference's
(DiT/forget.py:150-183,366-394; DiT/generate_fisher.py:132-151,296-316):
  forget           {results_dir}/{idx:03d}-{model}-forget-{class}-{method}-{loss}-lr{lr}-f{fa}-r{ra}/checkpoints/{steps:07d}.pt
                   = {"model": state_dict with `module.` keys, "ema": bare keys, "opt": AdamW state_dict, "args": args}
  generate_fisher  {mask_path}/{class}/forget_fisher.pt, remain_fisher.pt   (dict name -> fp32 CPU tensor | int 0)
What the path does not own comes through `DiTHooks`: the DiT network, the Gaussian-diffusion loss object, the
forget / remain datasets and the VAE encoder.  The default hooks import them from the reference's DiT/ directory
(`models`, `diffusion`, `diffusers`); a harness passes its own (synthetic latents, an identity encoder).
Between `loss.backward()` and the next forward everything runs in the CUDA kernels (DiffusionUnlearner "dit":
AdamW wd 0, clip on the forget step only, EMA 0.9999 including the frozen `pos_embed`).
`--method joint` (one backward on remain + alpha * forget) maps to ONE unmasked AdamW step + EMA, as in the
reference (its mask multiply sits under `if args.method == "ron"`).
"""
from __future__ import annotations

import argparse
import logging
import os
from dataclasses import dataclass
from glob import glob
from time import time
from typing import Callable, Optional

import torch
from torch.utils.data import DataLoader

from .common import cosine_lr_scheduler, cycle
from .diffusion import DiffusionUnlearner


@dataclass
class DiTHooks:
    model_factory: Callable          # args -> nn.Module (DiT_models[args.model](input_size=, num_classes=)), weights loaded
    diffusion_factory: Callable      # () -> object with .num_timesteps and .training_losses(model, x, t, model_kwargs)
    unlearn_dataset: Callable        # args -> (forget_dataset, remain_dataset)
    encode: Callable                 # (x, device) -> latents   (vae.encode(x).latent_dist.sample().mul_(0.18215))
    sample_visualization: Optional[Callable] = None


def reference_hooks() -> DiTHooks:
    """The reference's own DiT / diffusion / VAE / dataset code (needs its DiT/ directory on sys.path)."""
    try:
        from diffusers.models import AutoencoderKL
        from diffusion import create_diffusion
        from download import find_model
        from models import DiT_models
        from torchvision import transforms
        from forget import center_crop_arr, get_unlearn_dataset
    except Exception as e:                                       # pragma: no cover - depends on the deployment
        raise ImportError("the default DiT hooks import the reference's models / diffusion / forget modules and "
                          "diffusers: run from the reference's DiT/ directory or pass DiTHooks") from e
    vae = {}

    def model_factory(args):
        model = DiT_models[args.model](input_size=args.image_size // 8, num_classes=args.num_classes)
        if args.ckpt:
            model.load_state_dict(find_model(args.ckpt))
        return model

    def dataset(args):
        tf = transforms.Compose([
            transforms.Lambda(lambda im: center_crop_arr(im, args.image_size)), transforms.RandomHorizontalFlip(),
            transforms.ToTensor(), transforms.Normalize(mean=[0.5] * 3, std=[0.5] * 3, inplace=True)])
        return get_unlearn_dataset(args.data_path, args.forget_class, tf)

    def encode(x, device, args=None):
        if "m" not in vae:
            vae["m"] = AutoencoderKL.from_pretrained(f"stabilityai/sd-vae-ft-{getattr(args, 'vae', 'ema')}").to(device)
        return vae["m"].encode(x).latent_dist.sample().mul_(0.18215)

    return DiTHooks(model_factory, lambda: create_diffusion(timestep_respacing=""), dataset, encode)


def _common_flags(p: argparse.ArgumentParser) -> None:
    p.add_argument("--data-path", type=str, required=True)
    p.add_argument("--results-dir", type=str, default="results")
    p.add_argument("--model", type=str, default="DiT-XL/2")
    p.add_argument("--image-size", type=int, choices=[256, 512], default=256)
    p.add_argument("--num-classes", type=int, default=1000)
    p.add_argument("--n-iters", type=int, default=2000)
    p.add_argument("--batch-size", type=int, default=1)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--vae", type=str, choices=["ema", "mse"], default="ema")
    p.add_argument("--num-workers", type=int, default=4)
    p.add_argument("--log-every", type=int, default=100)
    p.add_argument("--ckpt", type=str, default=None)
    p.add_argument("--forget-class", type=int, required=True, nargs="?", help="class to forget")
    p.add_argument("--mask-path", type=str, default=None, help="the path to store mask")


def forget_parser() -> argparse.ArgumentParser:
    """DiT/forget.py:366-394."""
    p = argparse.ArgumentParser(prog="forget")
    _common_flags(p)
    p.add_argument("--lr", type=float, default=1e-5)
    p.add_argument("--ckpt-every", type=int, default=1000)
    p.add_argument("--snapshot-every", type=int, default=500)
    p.add_argument("--method", type=str, required=True, nargs="?", help="unlearning method")
    p.add_argument("--unlearn-loss", type=str, default="ga", help="unlearning loss")
    p.add_argument("--grad-clip", type=float, default=1.0, help="clip gradient")
    p.add_argument("--forget-alpha", type=float, default=1.0, help="forget loss alpha")
    p.add_argument("--decay-forget-alpha", action="store_true", help="whether decay forget loss alpha")
    p.add_argument("--remain-alpha", type=float, default=1.0, help="remain loss alpha")
    return p


def generate_fisher_parser() -> argparse.ArgumentParser:
    """DiT/generate_fisher.py:296-316."""
    p = argparse.ArgumentParser(prog="generate_fisher")
    _common_flags(p)
    return p


def _experiment_dir(args, suffix: str) -> str:
    os.makedirs(args.results_dir, exist_ok=True)
    index = len(glob(f"{args.results_dir}/*"))
    name = args.model.replace("/", "-")
    exp = f"{args.results_dir}/{index:03d}-{name}-{suffix}"
    os.makedirs(f"{exp}/checkpoints", exist_ok=True)
    return exp


def _loaders(args, hooks):
    forget_ds, remain_ds = hooks.unlearn_dataset(args)
    kw = dict(batch_size=args.batch_size, shuffle=True, num_workers=args.num_workers, pin_memory=True)
    return forget_ds, remain_ds, cycle(DataLoader(forget_ds, **kw)), cycle(DataLoader(remain_ds, **kw))


def forget_main(args, hooks: Optional[DiTHooks] = None, *, device=None, gradient_tap: Optional[Callable] = None) -> str:
    """DiT/forget.py:150-355.  Returns the checkpoint path."""
    hooks = hooks or reference_hooks()
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(args.seed)
    print(f"Starting seed={args.seed}")
    exp = _experiment_dir(args, f"forget-{args.forget_class}-{args.method}-{args.unlearn_loss}-lr{args.lr}"
                                f"-f{args.forget_alpha}-r{args.remain_alpha}")
    logger = logging.getLogger(__name__)
    logger.info(f"Experiment directory created at {exp}")
    assert args.image_size % 8 == 0, "Image size must be divisible by 8 (for the VAE encoder)."
    if args.method not in ("ron", "joint"):
        raise NotImplementedError(args.method)
    model = hooks.model_factory(args).to(device)
    diffusion = hooks.diffusion_factory()
    # AdamW(lr, wd 0) :199 ; ema = deepcopy(model), update_ema(decay 0) :187,225 -> the slow weights start as a copy
    un = DiffusionUnlearner(model, "dit", device=device, lr=args.lr, clip_forget=args.grad_clip)
    logger.info(f"DiT Parameters: {sum(p.numel() for p in model.parameters()):,}")
    forget_ds, remain_ds, forget_iter, remain_iter = _loaders(args, hooks)
    logger.info(f"Forget Dataset contains {len(forget_ds):,} images ({args.data_path})")
    logger.info(f"Remain Dataset contains {len(remain_ds):,} images ({args.data_path})")
    model.train()                      # enables embedding dropout for classifier-free guidance
    use_mask = bool(args.mask_path)
    if use_mask:
        logger.info(f"Load mask from {args.mask_path}")
        un.load_mask(args.mask_path)   # uploaded once; the reference moves it to the GPU every iteration (:289-292)
    mhp = un.mhp
    mhp.zero_grad()
    train_steps = log_steps = 0
    running_f = running_r = 0.0
    start = time()
    logger.info(f"Training for {args.n_iters} iterations...")
    for step in range(args.n_iters):
        model.train()
        x, y = next(forget_iter)
        x, y = x.to(device), y.to(device)
        t = torch.randint(0, diffusion.num_timesteps, (x.shape[0],), device=device)
        with torch.no_grad():
            x = hooks.encode(x, device)
        if args.unlearn_loss == "ga":
            ori_forget = -diffusion.training_losses(model, x, t, dict(y=y))["loss"].mean()
        elif args.unlearn_loss == "rl":
            pseudo = torch.full(y.shape, (args.forget_class + 100) % 1000, device=y.device)
            ori_forget = diffusion.training_losses(model, x, t, dict(y=pseudo))["loss"].mean()
        else:
            raise NotImplementedError(args.unlearn_loss)
        if args.method == "ron":
            alpha = cosine_lr_scheduler(args.forget_alpha, step, args.n_iters) if args.decay_forget_alpha \
                else args.forget_alpha
            (alpha * ori_forget).backward()
            if gradient_tap:
                gradient_tap("forget", mhp.grads())
            mhp.forget_step(use_mask=use_mask, max_norm=args.grad_clip)                 # :285-299 in one pass
        x, y = next(remain_iter)
        x, y = x.to(device), y.to(device)
        with torch.no_grad():
            x = hooks.encode(x, device)
        t = torch.randint(0, diffusion.num_timesteps, (x.shape[0],), device=device)
        ori_remain = diffusion.training_losses(model, x, t, dict(y=y))["loss"].mean()
        loss = ori_remain + args.forget_alpha * ori_forget if args.method == "joint" else ori_remain
        loss.backward()
        if gradient_tap:
            gradient_tap("remain", mhp.grads())
        mhp.remain_step(ema=True)                                                       # opt.step ; update_ema :320-322
        running_f += float(ori_forget.detach())
        running_r += float(ori_remain.detach())
        log_steps += 1
        train_steps += 1
        if train_steps % args.log_every == 0:
            logger.info(f"(step={train_steps:07d}) Forget Loss: {running_f / log_steps:.4f}, Remain Loss: "
                        f"{running_r / log_steps:.4f}, Train Steps/Sec: {log_steps / (time() - start):.2f}")
            running_f = running_r = 0.0
            log_steps, start = 0, time()
        if hooks.sample_visualization and train_steps % args.snapshot_every == 0 and train_steps > 0:
            hooks.sample_visualization(model, diffusion, train_steps, f"{exp}/checkpoints")
    path = f"{exp}/checkpoints/{train_steps:07d}.pt"
    torch.save(un.checkpoint(args=args), path)
    logger.info(f"Saved checkpoint to {path}")
    model.eval()
    logger.info("Done!")
    return path


def generate_fisher_main(args, hooks: Optional[DiTHooks] = None, *, device=None,
                         gradient_tap: Optional[Callable] = None) -> str:
    """DiT/generate_fisher.py:132-294.  Returns the directory the two Fisher files were written to."""
    hooks = hooks or reference_hooks()
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(args.seed)
    print(f"Starting seed={args.seed}")
    exp = _experiment_dir(args, "fisher")
    logger = logging.getLogger(__name__)
    logger.info(f"Experiment directory created at {exp}")
    model = hooks.model_factory(args).to(device)
    diffusion = hooks.diffusion_factory()
    un = DiffusionUnlearner(model, "dit", device=device, lr=0.0)
    _, _, forget_iter, remain_iter = _loaders(args, hooks)
    model.eval()
    out_dir = os.path.join(args.mask_path, str(args.forget_class))
    os.makedirs(out_dir, exist_ok=True)
    logger.info(f"Save in {out_dir}")
    for which, it in (("forget", forget_iter), ("remain", remain_iter)):
        logger.info(f"Getting {which} fisher...")
        state = dict(start=time(), log=0)

        def loss_fn(i, it=it, which=which, state=state):
            x, y = next(it)
            x, y = x.to(device), y.to(device)
            with torch.no_grad():
                x = hooks.encode(x, device)
            t = torch.randint(0, diffusion.num_timesteps, (x.shape[0],), device=device)
            loss = diffusion.training_losses(model, x, t, dict(y=y))["loss"].mean()
            state["log"] += 1
            if (i + 1) % args.log_every == 0:
                logger.info(f"(step={i + 1:07d}) Train Steps/Sec: {state['log'] / (time() - state['start']):.2f}")
                state.update(start=time(), log=0)
            return _Tapped(loss, (lambda: gradient_tap(which, un.mhp.grads())) if gradient_tap else None)

        un.generate_fisher(which, args.n_iters, loss_fn, out_dir)      # F += grad**2 / n_iters ; *_fisher.pt
    logger.info("Done!")
    return out_dir


class _Tapped:
    def __init__(self, loss, after):
        self.loss, self.after = loss, after

    def backward(self):
        self.loss.backward()
        if self.after:
            self.after()


def main(argv=None) -> int:
    import sys
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ("forget", "generate_fisher"):
        raise SystemExit("usage: python -m sfron_b200.methods.dit {forget|generate_fisher} [reference flags]")
    logging.basicConfig(level=logging.INFO, format="[%(asctime)s] %(message)s")
    if argv[0] == "forget":
        forget_main(forget_parser().parse_args(argv[1:]))
    else:
        generate_fisher_main(generate_fisher_parser().parse_args(argv[1:]))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
