#!/usr/bin/env python
"""End-to-end BASELINE config 1 on one B200: ResNet-18 (CIFAR-shaped synthetic data), SFR-on.

  stock : the reference's Classification SFR-on method AS WRITTEN, restated here because /root/reference
          does not exist on the GPU box — sfron.py:262-336 (Fisher on the CPU per tensor, ratio mask) and
          sfron.py:189-259 (per-parameter `grad *= mask[name].to(device)`, clip_grad_norm, SGD.step,
          update_parameters + deepcopy(model) every iteration, CosineAnnealingLR).
  ours  : `sfron_b200.methods.create_unlearn_method("SFRon")` — the same class surface with the hot path on
          the flat-vector kernels.

Model forward/backward in PyTorch on the GPU in both arms (tools/resnet18_cifar.py harness, 11,173,962 params).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import tempfile
import time
from copy import deepcopy

import torch
import torch.nn as nn
from torch.utils.data import DataLoader, TensorDataset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from resnet18_cifar import ResNet18Harness  # noqa: E402


def loaders(bs, n_forget, n_retain, seed=0):
    g = torch.Generator().manual_seed(seed)
    fx, fy = torch.randn(n_forget, 3, 32, 32, generator=g), torch.randint(0, 10, (n_forget,), generator=g)
    rx, ry = torch.randn(n_retain, 3, 32, 32, generator=g), torch.randint(0, 10, (n_retain,), generator=g)
    mk = lambda x, y: DataLoader(TensorDataset(x, y), batch_size=bs, shuffle=False)
    return dict(forget_train=mk(fx, fy), retain_train=mk(rx, ry), forget_valid=None, retain_valid=None)


def cycle(dl):
    while True:
        for d in dl:
            yield d


def stock(model, dls, n_iters, dev):
    """Reference form (hyper-parameters of sfron.py:100-123)."""
    crit = nn.CrossEntropyLoss()
    ce_none = nn.CrossEntropyLoss(reduction="none")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fish = {}
    for which in ("forget_train", "retain_train"):
        opt0 = torch.optim.SGD(model.parameters(), lr=0)
        acc = {n: 0 for n, _ in model.named_parameters()}
        model.eval()
        for x, y in dls[which]:
            x, y = x.to(dev), y.to(dev)
            loss = crit(model(x), y)
            opt0.zero_grad()
            loss.backward()
            with torch.no_grad():
                for n, p in model.named_parameters():
                    if p.grad is not None:
                        acc[n] += p.grad.data.cpu() ** 2 / len(dls[which])
        fish[which] = acc
    mask = {n: ((fish["forget_train"][n] + 1e-15) / (fish["retain_train"][n] + 1e-15)) >= 1 for n in fish["forget_train"]}
    torch.cuda.synchronize()
    t_prepare = time.perf_counter() - t0

    opt = torch.optim.SGD(model.parameters(), 0.01, momentum=0.9, weight_decay=5e-4)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=n_iters)
    ori = deepcopy(model)
    beta = 1.0
    fi, ri = cycle(dls["forget_train"]), cycle(dls["retain_train"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for step in range(n_iters):
        model.train()
        if step % 5 == 0:
            alpha = 25 * (1 + math.cos(math.pi * step / n_iters)) / 2
            x, y = next(fi)
            x, y = x.to(dev), y.to(dev)
            opt.zero_grad()
            l = ce_none(model(x), y)
            coef = 1 / (torch.pow(l.detach().clone(), 0.5) + 1e-15)
            loss = -alpha * ((coef / coef.sum()) * l * x.shape[0]).mean()
            loss.backward()
            for n, p in model.named_parameters():
                if p.grad is not None:
                    p.grad *= mask[n].to(p.grad.device)
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=7.0)
            opt.step()
        x, y = next(ri)
        x, y = x.to(dev), y.to(dev)
        opt.zero_grad()
        crit(model(x), y).backward()
        opt.step()
        with torch.no_grad():
            for ps, pm in zip(model.parameters(), ori.parameters()):
                ps.detach().copy_((1 - beta) * pm.detach().to(ps.device) + beta * ps.detach())
        ori = deepcopy(model)
        sched.step()
    torch.cuda.synchronize()
    return t_prepare, (time.perf_counter() - t0) / n_iters


def ours(model, dls, n_iters, dev, cuda_graph=False):
    from sfron_b200.methods import create_unlearn_method
    with tempfile.TemporaryDirectory() as tmp:
        m = create_unlearn_method("SFRon")(model, nn.CrossEntropyLoss(), tmp,
                                           argparse.Namespace(num_classes=10, seed=0, cuda_graph=cuda_graph))
        m.n_iters, m.log_freq = n_iters, 10 ** 9
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.prepare_unlearn(dls)
        torch.cuda.synchronize()
        t_prepare = time.perf_counter() - t0
        t0 = time.perf_counter()
        m.get_unlearned_model()
        torch.cuda.synchronize()
        return t_prepare, (time.perf_counter() - t0) / n_iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--batch-size", type=int, default=128)      # Classification/scripts/unlearn.sh
    ap.add_argument("--cuda-graph", action="store_true",
                    help="also time the loop with every iteration replayed from a CUDA graph (device-side cosine LR)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dls = loaders(args.batch_size, 512, 2048)
    res = {"model": "ResNet-18 harness, 11,173,962 params", "batch_size": args.batch_size, "iters": args.iters,
           "forget_batches": len(dls["forget_train"]), "retain_batches": len(dls["retain_train"])}
    arms = [("stock", stock), ("ours", ours)]
    if args.cuda_graph:
        arms.append(("ours_cudagraph", lambda m, d, n, dv: ours(m, d, n, dv, cuda_graph=True)))
    for name, fn in arms:
        torch.manual_seed(0)
        model = ResNet18Harness().to(dev)
        fn(deepcopy(model), dls, 10, dev)                        # warm-up (cuDNN autotune, allocator)
        tp, ts = fn(model, dls, args.iters, dev)
        res[name] = {"prepare_s": round(tp, 4), "loop_steps_per_s": round(1 / ts, 2)}
    line = json.dumps(res)
    print(line, flush=True)
    if args.out:
        with open(args.out, "a") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
