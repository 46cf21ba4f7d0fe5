#!/usr/bin/env python
"""End-to-end BASELINE configs 2 and 4 on one B200: Fisher diagonals -> saliency mask -> SFR-on forget loop.

  --family ddpm  config 2: class-conditional DDPM CIFAR-10 U-Net (tools/ddpm_unet.py, 38,632,323 params), batches of
                 128 synthetic 32x32 images, Fisher of the clipped guided loss, ratio mask, 50-step `ron` loop
                 (adaptive gradient ascent, cosine-decayed forget alpha, Adam, clip 1.0 on both steps, EMA 1e-4).
  --family sd    config 4: Stable Diffusion v1.x latent U-Net (tools/sd_unet.py, 859,520,964 params), synthetic
                 4x64x64 latents and 77x768 text embeddings, Fisher of the guided loss, ratio mask, concept-erasure
                 loop (forget: match a pseudo-prompt prediction; remain: noise MSE; Adam, no clip, no EMA).

Two arms on the SAME GPU, same model, same synthetic batches:
  stock : the reference's stages as written, restated here because /root/reference does not exist on the GPU box —
          DDPM/runners/diffusion.py:1244-1299,1309-1364 + DDPM/generate_fisher_mask.py:31-48 + :1075-1180;
          SD/train-scripts/generate_fisher.py:36-79,86-129 + generate_fisher_mask.py:36-48 + nsfw_removal.py:108-173
          (Fisher accumulated on the CPU per tensor, mask on the CPU, mask uploaded per tensor every step,
          clip_grad_norm_, torch Adam, per-tensor EMA).  For SD the mask multiply follows the sibling scripts'
          intent (gradient_ascent.py:94-99); nsfw_removal.py's `n in parameters` test never fires.
  ours  : `sfron_b200.methods.diffusion.DiffusionUnlearner` / `methods.masks.generate_fisher_mask` — the same stages,
          files and formats with everything after `backward()` on the flat-vector kernels.
Model forward/backward runs in PyTorch in both arms.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def cosine(base, step, total):
    return base * (1 + math.cos(math.pi * step / total)) / 2


class Family:
    """Model + synthetic batches + the losses of one sub-project; identical for both arms."""

    def __init__(self, name, dev, batch_size, seed=0):
        self.name, self.dev, self.bs = name, dev, batch_size
        self.gen = torch.Generator(device=dev)
        self.seed = seed
        if name == "ddpm":
            from ddpm_unet import DDPMCondUNet, ddpm_alphas_cumprod
            torch.manual_seed(seed)
            self.model = DDPMCondUNet().to(dev)
            self.ac = ddpm_alphas_cumprod(dev)
            self.clip_fisher = self.clip_forget = self.clip_remain = 1.0
            self.lr, self.ema_mu, self.prefix = 1e-4, 1e-4, "module."
            self.fisher_names = ("forget_fisher.pt", "remain_fisher.pt", "fisher_{th}.pt")
            self.forget_alpha, self.decay = 5.0, True
        else:
            from sd_unet import SDUNet, sd_alphas_cumprod
            torch.manual_seed(seed)
            self.model = SDUNet().to(dev)
            self.model.randomise_zero_layers()
            self.ac = sd_alphas_cumprod(dev)
            self.clip_fisher = self.clip_forget = self.clip_remain = None
            self.lr, self.ema_mu, self.prefix = 1e-5, None, ""
            self.fisher_names = ("nude_forget.pt", "nude_remain.pt", "nude_mask_{th}.pt")
            self.forget_alpha, self.decay = 1.0, False
            g = torch.Generator(device=dev).manual_seed(seed + 7)
            self.ctx_forget = torch.randn(1, 77, 768, device=dev, generator=g)     # "a photo of a nude person"
            self.ctx_pseudo = torch.randn(1, 77, 768, device=dev, generator=g)     # "... wearing clothes"
            self.ctx_null = torch.randn(1, 77, 768, device=dev, generator=g)       # ""
        self.n = sum(p.numel() for p in self.model.parameters())

    def reseed(self, stream):
        self.gen.manual_seed(self.seed * 1000 + stream)

    # ---- losses (the closures handed to either arm) ------------------------------------------------
    def _ddpm_batch(self, forget):
        n, dev, g = self.bs, self.dev, self.gen
        x = 2 * torch.rand(n, 3, 32, 32, device=dev, generator=g) - 1
        c = torch.zeros(n, dtype=torch.long, device=dev) if forget else torch.randint(1, 10, (n,), device=dev, generator=g)
        e = torch.randn(n, 3, 32, 32, device=dev, generator=g)
        t = torch.randint(0, 1000, (n // 2 + 1,), device=dev, generator=g)
        return x, torch.cat([t, 1000 - t - 1])[:n], c, e                  # antithetic timesteps

    def _sd_batch(self):
        n, dev, g = self.bs, self.dev, self.gen
        z = torch.randn(n, 4, 64, 64, device=dev, generator=g)
        return z, torch.randint(0, 1000, (n,), device=dev, generator=g), torch.randn(n, 4, 64, 64, device=dev, generator=g)

    def draw(self, forget):
        """One synthetic batch (fresh tensors) for the forget / remain losses."""
        return self._ddpm_batch(forget) if self.name == "ddpm" else self._sd_batch()

    def fisher_loss(self, forget, batch=None):
        batch = self.draw(forget) if batch is None else batch
        if self.name == "ddpm":
            from ddpm_unet import eps_loss
            x, t, c, e = batch
            return eps_loss(self.model, x, t, c, e, self.ac, mode="test", cond_scale=2.0)
        from sd_unet import guided_eps_loss
        z, t, e = batch
        ctx = (self.ctx_forget if forget else self.ctx_pseudo).expand(self.bs, -1, -1)
        return guided_eps_loss(self.model, z, t, ctx, self.ctx_null.expand(self.bs, -1, -1), e, self.ac, cond_scale=7.5)

    def forget_loss(self, batch=None):
        batch = self.draw(True) if batch is None else batch
        if self.name == "ddpm":                       # adaga: -adaptive_loss(...), DDPM/functions/losses.py:49-69
            from ddpm_unet import eps_loss
            x, t, c, e = batch
            per = eps_loss(self.model, x, t, c, e, self.ac, mode="train", keepdim=True)
            coef = 1 / (torch.pow(per.detach().clone(), 0.5) + 1e-8)
            return -((coef / coef.sum()) * per * self.bs).mean(dim=0)
        z, t, e = batch                               # nsfw_removal.py:141-155
        a = self.ac.index_select(0, t).view(-1, 1, 1, 1)
        zt = a.sqrt() * z + (1 - a).sqrt() * e
        out = self.model(zt, t, self.ctx_forget.expand(self.bs, -1, -1))
        with torch.no_grad():
            pseudo = self.model(zt, t, self.ctx_pseudo.expand(self.bs, -1, -1))
        return torch.nn.functional.mse_loss(out, pseudo)

    def remain_loss(self, batch=None):
        batch = self.draw(False) if batch is None else batch
        if self.name == "ddpm":
            from ddpm_unet import eps_loss
            x, t, c, e = batch
            return eps_loss(self.model, x, t, c, e, self.ac, mode="train")
        from sd_unet import eps_mse_loss
        z, t, e = batch
        return eps_mse_loss(self.model, z, t, self.ctx_pseudo.expand(self.bs, -1, -1), e, self.ac)


def warmup(fam: Family):
    """cuDNN autotune / allocator growth for every loss shape, before anything is timed."""
    for fn in (lambda: fam.fisher_loss(True), fam.forget_loss, fam.remain_loss):
        for _ in range(2):
            fn().backward()
    for p in fam.model.parameters():
        p.grad = None


def _timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


def stock(fam: Family, n_fisher, n_iters, tmp):
    model, dev = fam.model, fam.dev
    named = [(fam.prefix + n, p) for n, p in model.named_parameters()]
    opt0 = torch.optim.Adam(model.parameters(), lr=fam.lr)

    def fisher_stage():
        for which, fname in ((True, fam.fisher_names[0]), (False, fam.fisher_names[1])):
            fam.reseed(1 if which else 2)
            acc = {n: 0 for n, _ in named}
            model.eval()
            for _ in range(n_fisher):
                loss = fam.fisher_loss(which)
                opt0.zero_grad()
                loss.backward()
                if fam.clip_fisher is not None:
                    torch.nn.utils.clip_grad_norm_(model.parameters(), fam.clip_fisher)
                with torch.no_grad():
                    for n, p in named:
                        if p.grad is not None:
                            acc[n] += p.grad.data.cpu() ** 2 / n_fisher
            torch.save(acc, os.path.join(tmp, fname))

    def mask_stage():
        ff = torch.load(os.path.join(tmp, fam.fisher_names[0]))
        rf = torch.load(os.path.join(tmp, fam.fisher_names[1]))
        mask, zeros, total = {}, 0, 0
        for n in ff:
            w = (ff[n] + 1e-15) / (rf[n] + 1e-15) >= 1.0
            zeros += w.numel() - int(w.count_nonzero())
            total += w.numel()
            mask[n] = w
        print(f"Total sparsity th:1.0 weight:{zeros / total * 100}", file=sys.stderr)
        torch.save(mask, os.path.join(tmp, fam.fisher_names[2].format(th="1.0")))
        return mask

    def loop_stage(n_steps):
        mask = torch.load(os.path.join(tmp, fam.fisher_names[2].format(th="1.0")))
        opt = torch.optim.Adam(model.parameters(), lr=fam.lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
        shadow = {n: p.data.clone() for n, p in named} if fam.ema_mu is not None else None
        model.train()
        fam.reseed(3)
        for step in range(n_steps):
            alpha = cosine(fam.forget_alpha, step, n_steps) if fam.decay else fam.forget_alpha
            loss = alpha * fam.forget_loss()
            opt.zero_grad()
            loss.backward()
            for n, p in named:
                if p.grad is not None:
                    p.grad *= mask[n].to(p.grad.device)
            if fam.clip_forget is not None:
                torch.nn.utils.clip_grad_norm_(model.parameters(), fam.clip_forget)
            opt.step()
            loss = fam.remain_loss()
            opt.zero_grad()
            loss.backward()
            if fam.clip_remain is not None:
                torch.nn.utils.clip_grad_norm_(model.parameters(), fam.clip_remain)
            opt.step()
            if shadow is not None:
                for n, p in named:
                    shadow[n].data = (1.0 - fam.ema_mu) * p.data + fam.ema_mu * shadow[n].data

    t_f, _ = _timed(fisher_stage)
    t_m, mask = _timed(mask_stage)
    _timed(lambda: loop_stage(2))
    t_l, _ = _timed(lambda: loop_stage(n_iters))
    return t_f, t_m, t_l, mask


def ours(fam: Family, n_fisher, n_iters, tmp, cuda_graph=False, graph_fisher=False):
    from sfron_b200.methods.diffusion import DiffusionUnlearner
    from sfron_b200.methods.masks import generate_fisher_mask
    un = DiffusionUnlearner(fam.model, "ddpm" if fam.name == "ddpm" else "sd", lr=fam.lr)

    def fisher_stage():
        fam.model.eval()
        for which, forget, stream in (("forget", True, 1), ("remain", False, 2)):
            fam.reseed(stream)
            if not graph_fisher:
                un.generate_fisher(which, n_fisher, lambda i: fam.fisher_loss(forget), out_dir=tmp)
                continue
            sb = fam.draw(forget)

            def refill(i):
                for dst, src in zip(sb, fam.draw(forget)):
                    dst.copy_(src)

            un.generate_fisher(which, n_fisher, lambda i: fam.fisher_loss(forget, sb), out_dir=tmp, cuda_graph=True,
                               refill=refill)

    def mask_stage():
        return generate_fisher_mask(tmp, 1.0, forget_name=fam.fisher_names[0], remain_name=fam.fisher_names[1],
                                    out_fmt=fam.fisher_names[2])

    def loop_stage(n_steps):
        un.load_mask(os.path.join(tmp, fam.fisher_names[2].format(th="1.0")))
        fam.reseed(3)
        if not cuda_graph:
            un.forget(n_steps, lambda i: fam.forget_loss(), lambda i: fam.remain_loss(), forget_alpha=fam.forget_alpha,
                      decay_forget_alpha=fam.decay)
            return
        # whole iteration in one CUDA graph: the batches live in static tensors, refilled before every replay
        fb, rb = fam.draw(True), fam.draw(False)

        def refill(step):
            for dst, src in zip(fb, fam.draw(True)):
                dst.copy_(src)
            for dst, src in zip(rb, fam.draw(False)):
                dst.copy_(src)

        un.forget(n_steps, lambda i: fam.forget_loss(fb), lambda i: fam.remain_loss(rb), forget_alpha=fam.forget_alpha,
                  decay_forget_alpha=fam.decay, cuda_graph=True, refill=refill)

    t_f, _ = _timed(fisher_stage)
    t_m, path = _timed(mask_stage)
    _timed(lambda: loop_stage(2))
    t_l, _ = _timed(lambda: loop_stage(n_iters))
    if cuda_graph:
        fam.graph_capture_s = un.graph_capture_s
        t_l -= un.graph_capture_s               # one-off capture reported separately
    return t_f, t_m, t_l, torch.load(path, weights_only=False)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--family", required=True, choices=["ddpm", "sd"])
    ap.add_argument("--arm", default="both", choices=["both", "stock", "ours"])
    ap.add_argument("--iters", type=int, default=None, help="forget-loop iterations (config 2: 50)")
    ap.add_argument("--fisher-batches", type=int, default=None, help="batches per Fisher (forget and remain each)")
    ap.add_argument("--batch-size", type=int, default=None)
    ap.add_argument("--cuda-graph", action="store_true",
                    help="ours arm: DiffusionUnlearner.forget(cuda_graph=True) — the whole iteration captured once and replayed")
    ap.add_argument("--cuda-graph-fisher", action="store_true",
                    help="ours arm: also replay the Fisher stage from a CUDA graph.  Pays off only on long Fisher runs (the "
                         "reference's 2000 batches): each of the two captures costs 1-5 s, more than the handful of batches "
                         "this tool times (measured: DDPM 1.7 -> 4.9 s, SD 7.6 -> 19.0 s for 8 / 3 batches each)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    iters = args.iters or (50 if args.family == "ddpm" else 10)
    n_fisher = args.fisher_batches or (8 if args.family == "ddpm" else 3)
    bs = args.batch_size or (128 if args.family == "ddpm" else 2)
    res = {"family": args.family, "batch_size": bs, "fisher_batches_each": n_fisher, "loop_iters": iters, "n_gpus": 1,
           "data": "synthetic", "dtype": "fp32"}
    masks = {}
    for name, fn in (("stock", stock), ("ours", ours)):
        if args.arm not in ("both", name):
            continue
        fam = Family(args.family, dev, bs)
        res["params"] = fam.n
        warmup(fam)
        with tempfile.TemporaryDirectory() as tmp:
            if name == "ours":
                t_f, t_m, t_l, masks[name] = ours(fam, n_fisher, iters, tmp, cuda_graph=args.cuda_graph,
                                                  graph_fisher=args.cuda_graph_fisher)
            else:
                t_f, t_m, t_l, masks[name] = stock(fam, n_fisher, iters, tmp)
        res[name] = {"fisher_s": round(t_f, 4), "fisher_batches_per_s": round(2 * n_fisher / t_f, 3),
                     "mask_s": round(t_m, 4), "loop_s": round(t_l, 4), "loop_steps_per_s": round(iters / t_l, 3)}
        if name == "ours" and args.cuda_graph:
            res[name]["cuda_graph"] = True
            res[name]["graph_capture_s"] = round(fam.graph_capture_s, 4)
        del fam
        torch.cuda.empty_cache()
    if len(masks) == 2:
        # sanity, not parity (the arms' backward passes are separate non-deterministic GPU runs): mask agreement
        same = total = 0
        for n, m in masks["stock"].items():
            same += int((m == masks["ours"][n]).sum())
            total += m.numel()
        res["mask_agreement"] = same / total
        for k in ("fisher_s", "mask_s", "loop_s"):
            res.setdefault("speedup", {})[k[:-2]] = round(res["stock"][k] / res["ours"][k], 2)
    line = json.dumps(res)
    print(line, flush=True)
    if args.out:
        with open(args.out, "a") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
