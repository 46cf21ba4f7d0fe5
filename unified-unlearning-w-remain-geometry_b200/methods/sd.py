"""Drop-in entry points of the SD sub-project: `train-scripts/generate_fisher.py` and `train-scripts/nsfw_removal.py`.

    generate_nsfw_fisher(c_guidance, batch_size, epochs, lr, config_path, ckpt_path, diffusers_config_path, device,
                         image_size=512, num_timesteps=1000)            SD/train-scripts/generate_fisher.py:8-129
        -> fisher/nude_forget.pt, fisher/nude_remain.pt   (cwd-relative; U-Net-local keys, fp32 CPU tensors)
    nsfw_removal(train_method, forget_alpha, remain_alpha, batch_size, n_iters, lr, config_path, ckpt_path, mask_path,
                 diffusers_config_path, device, mask_threshold=None, image_size=512, ddim_steps=50)
                                                                        SD/train-scripts/nsfw_removal.py:38-220
        -> models/{name}/{name}.pt  (model.state_dict(); name as the reference builds it)
    parsers with the reference's flags: generate_fisher_parser() (:133-203), nsfw_removal_parser() (:256-351)

The LatentDiffusion wrapper (CLIP text encoder, VAE, q_sample, apply_model), the NSFW / not-NSFW image loaders and the
CompVis -> diffusers export are outside the path and come through `SDHooks`; by default they are the reference's own
(`train-scripts/convertModels.py`, `dataset.py`, `ldm/`).  The path itself — Fisher accumulation of the U-Net
gradients, the saliency mask, both Adam steps — runs in the CUDA kernels on the flat U-Net parameter vector.

Reference quirk kept visible: nsfw_removal.py:157-160 guards its mask multiply with `n in parameters`, a string looked
up in a list of tensors, so the reference NEVER applies its mask.  `apply_mask=False` (default) reproduces that run
bit for bit; `apply_mask=True` is what the sibling scripts (gradient_ascent.py:94-99, random_label.py:129-134) do.
"""
from __future__ import annotations

import argparse
import os
from dataclasses import dataclass
from time import time
from typing import Callable, Optional

import torch

from .common import cycle
from .diffusion import DiffusionUnlearner

WORD_NUDE = "a photo of a nude person"
WORD_WEAR = "a photo of a person wearing clothes"


@dataclass
class SDHooks:
    setup_model: Callable            # (config_path, ckpt_path, device) -> LatentDiffusion-like `model`
    setup_data: Callable             # (batch_size, image_size) -> (forget_dl, remain_dl) of image batches
    save_model: Optional[Callable] = None     # (model, name, num, config_path, diffusers_config_path) beyond the .pt


def reference_hooks() -> SDHooks:
    try:
        from convertModels import savemodelDiffusers
        from dataset import setup_forget_nsfw_data, setup_model
    except Exception as e:                                       # pragma: no cover - depends on the deployment
        raise ImportError("the default SD hooks import the reference's train-scripts/dataset.py and convertModels.py: "
                          "run from the reference's SD/ directory or pass SDHooks") from e

    def export(model, name, num, config_path, diffusers_config_path):
        file_name = f"{name}-step_{num}" if num is not None else name
        print("Saving Model in Diffusers Format")
        savemodelDiffusers(f"models/{name}", file_name, config_path, diffusers_config_path, device="cpu")

    return SDHooks(setup_model, setup_forget_nsfw_data, export)


def _unet(model) -> torch.nn.Module:
    return model.model.diffusion_model


def _save_compvis(model, name, num) -> str:
    """save_model(..., save_compvis=True) of nsfw_removal.py:222-246."""
    folder = f"models/{name}"
    os.makedirs(folder, exist_ok=True)
    path = f"{folder}/{name}-step_{num}.pt" if num is not None else f"{folder}/{name}.pt"
    torch.save(model.state_dict(), path)
    print(f"Saving Model in compvis Format at {path}")
    return path


def generate_nsfw_fisher(c_guidance, batch_size, epochs, lr, config_path, ckpt_path, diffusers_config_path, device,
                         image_size=512, num_timesteps=1000, *, hooks: Optional[SDHooks] = None,
                         gradient_tap: Optional[Callable] = None) -> None:
    hooks = hooks or reference_hooks()
    model = hooks.setup_model(config_path, ckpt_path, device)
    forget_dl, remain_dl = hooks.setup_data(batch_size, image_size)
    print(len(forget_dl), len(remain_dl))
    model.eval()
    criteria = torch.nn.MSELoss()
    un = DiffusionUnlearner(_unet(model), "sd", device=device, lr=lr)
    os.makedirs("fisher", exist_ok=True)
    for which, dl, word in (("forget", forget_dl, WORD_NUDE), ("remain", remain_dl, WORD_WEAR)):
        batches = iter(dl)

        def loss_fn(i, batches=batches, word=word, which=which):
            images = next(batches).to(device)
            prompts, null_prompts = [word] * batch_size, [""] * batch_size
            print(prompts)
            x, emb = model.get_input({"jpg": images.permute(0, 2, 3, 1), "txt": prompts}, model.first_stage_key)
            _, null_emb = model.get_input({"jpg": images.permute(0, 2, 3, 1), "txt": null_prompts}, model.first_stage_key)
            t = torch.randint(0, model.num_timesteps, (x.shape[0],), device=device).long()
            noise = torch.randn_like(x, device=device)
            noisy = model.q_sample(x_start=x, t=t, noise=noise)
            preds = (1 + c_guidance) * model.apply_model(noisy, t, emb) - c_guidance * model.apply_model(noisy, t, null_emb)
            loss = -criteria(noise, preds)                                       # generate_fisher.py:66-70
            return _Tapped(loss, (lambda: gradient_tap(which, un.mhp.grads())) if gradient_tap else None)

        un.generate_fisher(which, len(dl), loss_fn, "fisher")                    # F += grad**2 / len(dl) ; nude_*.pt


def nsfw_removal(train_method, forget_alpha, remain_alpha, batch_size, n_iters, lr, config_path, ckpt_path, mask_path,
                 diffusers_config_path, device, mask_threshold=None, image_size=512, ddim_steps=50, *,
                 hooks: Optional[SDHooks] = None, apply_mask: bool = False,
                 gradient_tap: Optional[Callable] = None) -> str:
    hooks = hooks or reference_hooks()
    model = hooks.setup_model(config_path, ckpt_path, device)
    criteria = torch.nn.MSELoss()
    forget_dl, remain_dl = hooks.setup_data(batch_size, image_size)
    unet = _unet(model)
    if train_method not in ("xattn", "full"):
        raise ValueError(train_method)
    for name, param in unet.named_parameters():          # train only the cross-attention layers, or everything
        param.requires_grad_(train_method == "full" or "attn2" in name)
        if train_method == "xattn" and "attn2" in name:
            print(name)
    model.train()
    # torch.optim.Adam(parameters, lr) :81 — no clip, no EMA; frozen tensors keep their names in the files
    un = DiffusionUnlearner(unet, "sd", device=device, lr=lr)
    if mask_path:
        print(f"Load saliency mask from {mask_path} with threshold={mask_threshold}")
        if apply_mask:
            un.load_mask(os.path.join(mask_path, f"nude_mask_{mask_threshold}.pt"))
        name = f"compvis-nsfw-mask{mask_threshold}-method_sfron-lr{lr}_fa{forget_alpha}_ra{remain_alpha}"
    else:
        name = f"compvis-nsfw-method_sfron-lr{lr}_fa{forget_alpha}_ra{remain_alpha}"
    print(f"prompt of NSFW: {WORD_NUDE}")
    print(f"prompt of non-NSFW: {WORD_WEAR}")
    forget_iter, remain_iter = cycle(forget_dl), cycle(remain_dl)
    mhp = un.mhp
    mhp.zero_grad()
    train_steps = log_steps = 0
    running_f = running_r = 0.0
    start = time()
    for step in range(n_iters):
        model.train()
        forget_images, remain_images = next(forget_iter), next(remain_iter)
        jpg = forget_images.permute(0, 2, 3, 1)
        x, emb = model.get_input({"jpg": jpg, "txt": [WORD_NUDE] * batch_size}, model.first_stage_key)
        px, pemb = model.get_input({"jpg": jpg, "txt": [WORD_WEAR] * batch_size}, model.first_stage_key)
        t = torch.randint(0, model.num_timesteps, (x.shape[0],), device=model.device).long()
        noise = torch.randn_like(x, device=model.device)
        out = model.apply_model(model.q_sample(x_start=x, t=t, noise=noise), t, emb)
        pseudo = model.apply_model(model.q_sample(x_start=px, t=t, noise=noise), t, pemb).detach()
        ori_forget = criteria(out, pseudo)
        (forget_alpha * ori_forget).backward()
        if gradient_tap:
            gradient_tap("forget", mhp.grads())
        mhp.forget_step(use_mask=bool(mask_path) and apply_mask, max_norm=None)      # [mask ;] Adam step
        ori_remain = model.shared_step({"jpg": remain_images.permute(0, 2, 3, 1), "txt": [WORD_WEAR] * batch_size})[0]
        (remain_alpha * ori_remain).backward()
        if gradient_tap:
            gradient_tap("remain", mhp.grads())
        mhp.remain_step(max_norm=None, ema=False)                                    # Adam step
        running_f += float(ori_forget.detach())
        running_r += float(ori_remain.detach())
        log_steps += 1
        train_steps += 1
        if train_steps % 10 == 0:
            print(f"(step={train_steps:07d}) Forget Loss: {running_f / log_steps:.6f}, Remain Loss: "
                  f"{running_r / log_steps:.6f}, Train Steps/Sec: {log_steps / (time() - start):.2f}")
            running_f = running_r = 0.0
            log_steps, start = 0, time()
        if (train_steps + 1) % 200 == 0:
            _save_compvis(model, name, train_steps + 1)
            if hooks.save_model:
                hooks.save_model(model, name, train_steps + 1, config_path, diffusers_config_path)
    model.eval()
    path = _save_compvis(model, name, None)
    if hooks.save_model:
        hooks.save_model(model, name, None, config_path, diffusers_config_path)
    return path


class _Tapped:
    def __init__(self, loss, after):
        self.loss, self.after = loss, after

    def backward(self):
        self.loss.backward()
        if self.after:
            self.after()


def generate_fisher_parser() -> argparse.ArgumentParser:
    """SD/train-scripts/generate_fisher.py:133-203."""
    p = argparse.ArgumentParser(prog="generate_fisher")
    p.add_argument("--c_guidance", type=float, required=False, default=7.5)
    p.add_argument("--batch_size", type=int, required=False, default=8)
    p.add_argument("--epochs", type=int, required=False, default=1)
    p.add_argument("--lr", type=float, required=False, default=1e-5)
    p.add_argument("--ckpt_path", type=str, required=False, default="models/ldm/stable-diffusion-v1/sd-v1-4-full-ema.ckpt")
    p.add_argument("--config_path", type=str, required=False, default="configs/stable-diffusion/v1-inference.yaml")
    p.add_argument("--diffusers_config_path", type=str, required=False, default="diffusers_unet_config.json")
    p.add_argument("--device", type=str, required=False, default="4")      # the reference's default GPU index
    p.add_argument("--image_size", type=int, required=False, default=512)
    p.add_argument("--num_timesteps", type=int, required=False, default=1000)
    return p


def nsfw_removal_parser() -> argparse.ArgumentParser:
    """SD/train-scripts/nsfw_removal.py:256-351."""
    p = argparse.ArgumentParser(prog="SFR-on for SD")
    p.add_argument("--train_method", type=str, required=True)
    p.add_argument("--batch_size", type=int, required=False, default=8)
    p.add_argument("--n_iters", type=int, default=1000)
    p.add_argument("--lr", type=float, required=False, default=1e-5)     # (`type=int` in the reference: a typo there)
    p.add_argument("--config_path", type=str, required=False, default="configs/stable-diffusion/v1-inference.yaml")
    p.add_argument("--ckpt_path", type=str, required=False, default="models/ldm/stable-diffusion-v1/sd-v1-4-full-ema.ckpt")
    p.add_argument("--diffusers_config_path", type=str, required=False, default="diffusers_unet_config.json")
    p.add_argument("--device", type=str, required=False, default="0,0")
    p.add_argument("--image_size", type=int, required=False, default=512)
    p.add_argument("--ddim_steps", type=int, required=False, default=50)
    p.add_argument("--forget_alpha", type=float, required=False, default=1.0)
    p.add_argument("--remain_alpha", type=float, required=False, default=1.0)
    p.add_argument("--mask_path", type=str, required=False, default=None)
    p.add_argument("--mask_threshold", type=float, default=None)
    return p


def main(argv=None) -> int:
    import sys
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ("generate_fisher", "nsfw_removal"):
        raise SystemExit("usage: python -m sfron_b200.methods.sd {generate_fisher|nsfw_removal} [reference flags]")
    if argv[0] == "generate_fisher":
        a = generate_fisher_parser().parse_args(argv[1:])
        generate_nsfw_fisher(a.c_guidance, a.batch_size, a.epochs, a.lr, a.config_path, a.ckpt_path,
                             a.diffusers_config_path, f"cuda:{int(a.device)}", a.image_size, a.num_timesteps)
    else:
        a = nsfw_removal_parser().parse_args(argv[1:])
        nsfw_removal(a.train_method, a.forget_alpha, a.remain_alpha, a.batch_size, a.n_iters, a.lr, a.config_path,
                     a.ckpt_path, a.mask_path, a.diffusers_config_path, f"cuda:{int(a.device.split(',')[0])}",
                     a.mask_threshold, a.image_size, a.ddim_steps)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
