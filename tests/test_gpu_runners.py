"""Runner-level drop-in: `sfron_b200.methods.ddpm.Diffusion(args, config)` driven through the reference's own
flag and config field names (DDPM/train.py:21-102,145-168), end to end on the GPU — real forward / backward passes
of a small conditional network on synthetic loaders, the hot path in the CUDA kernels, files written where and how
the reference writes them.

Two independent checks per mode:
  * structure against the golden `ddpm_runner.pt` (the reference's `Diffusion` executed whole): file names, dict
    keys with the `module.` prefix, dtypes, the `[model, optimizer, step, ema]` checkpoint, optimizer param_groups
    and per-parameter step counts;
  * arithmetic against the CPU oracle on IDENTICAL inputs: the runner's gradient tap records the flat gradient of
    every backward pass exactly as the kernels receive it, and the oracle replays those gradients — Fisher and
    masks bit-exact / 1e-6 where a clip norm is involved, weights, moments and EMA within 1e-6.
"""
import argparse
import os

import pytest
import torch
from torch.utils.data import DataLoader, TensorDataset

from conftest import load_golden, unflat
from helpers_models import TinyCondNet
from oracle import sfron_oracle as O

pytestmark = pytest.mark.gpu


def close(a, b, rtol=1e-6):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    rms = b.pow(2).mean().sqrt().item() if b.numel() else 0.0
    return not bool(((a - b).abs() > rtol * (b.abs() + rms)).any())


def ns(**kw):
    return argparse.Namespace(**kw)


@pytest.fixture()
def setup(tmp_path):
    import sfron_b200  # noqa: F401
    from sfron_b200.methods.ddpm import DDPMHooks
    fx = load_golden("ddpm_runner.pt")
    h, pnames = fx["hyper"], fx["names"]
    n_iters = h["n_iters"]
    config = ns(
        diffusion=ns(beta_schedule="linear", beta_start=1e-4, beta_end=0.02, num_diffusion_timesteps=1000),
        optim=ns(lr=h["lr"], beta1=h["beta1"], eps=h["eps"], weight_decay=h["weight_decay"], grad_clip=h["grad_clip"],
                 optimizer="Adam", amsgrad=False),
        model=ns(ema=True, ema_rate=h["ema_rate"], type="simple"),
        training=ns(n_iters=n_iters, log_freq=10 ** 9, snapshot_freq=n_iters, save_freq=10 ** 9, lambd=1.0, gamma=1.0,
                    lmbda=1.0),
        data=ns(num_workers=0, n_classes=10), ckpt_dir=str(tmp_path / "out"))
    os.makedirs(config.ckpt_dir)
    os.makedirs(tmp_path / "ckpts")
    theta = unflat(fx["theta0"], pnames, fx["shapes"])
    ema0 = {k[len("module."):]: v.clone() for k, v in theta.items()}
    torch.save([theta, {}, 0, ema0], tmp_path / "ckpts" / "ckpt.pth")          # the reference's pretrain checkpoint
    g = torch.Generator().manual_seed(21)
    fxs, fys = torch.rand(6, 3, 8, 8, generator=g), torch.zeros(6, dtype=torch.long)
    rxs, rys = torch.rand(9, 3, 8, 8, generator=g), torch.randint(1, 10, (9,), generator=g)
    forget = DataLoader(TensorDataset(fxs, fys), batch_size=3, shuffle=False)
    remain = DataLoader(TensorDataset(rxs, rys), batch_size=3, shuffle=False)
    hooks = DDPMHooks(model_factory=TinyCondNet, forget_dataset=lambda a, c, label: (remain, forget),
                      data_transform=lambda c, x: 2 * x - 1.0,
                      fim_loader=lambda a, c, bs: DataLoader(TensorDataset(fxs, fys), batch_size=bs, shuffle=False))

    def args(**over):
        a = ns(ckpt_folder=str(tmp_path), label_to_forget=0, cond_scale=2.0, mask_path=None, forget_alpha=1.0,
               decay_forget_alpha=False, remain_alpha=1.0, method="ron", unlearn_loss="ga", n_chunks=250)
        for k, v in over.items():
            setattr(a, k, v)
        return a

    return fx, config, hooks, args, tmp_path, (len(forget), len(remain))


def run(config, hooks, args, method, seed):
    from sfron_b200.methods.ddpm import Diffusion
    torch.manual_seed(seed)
    runner = Diffusion(args, config, hooks=hooks)
    taps = []
    runner.gradient_tap = lambda kind, g: taps.append((kind, g.detach().float().cpu().clone()))
    getattr(runner, method)()
    torch.cuda.synchronize()
    return taps


def flat_of(d, names):
    return torch.cat([d[n].reshape(-1) for n in names])


def test_ddpm_runner_modes_end_to_end(setup, monkeypatch):
    from sfron_b200.methods.masks import generate_fisher_mask
    fx, config, hooks, args, tmp, (nf, nr) = setup
    pnames, shapes, h = fx["names"], fx["shapes"], fx["hyper"]
    names = [n[len("module."):] for n in pnames]
    n = fx["theta0"].numel()
    monkeypatch.chdir(tmp)                                   # generate_mask writes cwd-relative, as the reference

    # ---- --mode generate_fisher ------------------------------------------------------------------------
    taps = run(config, hooks, args(), "generate_fisher", 31)
    assert [k for k, _ in taps] == ["forget"] * nf + ["remain"] * nr
    mdir = tmp / "mask_0"
    ff = torch.load(mdir / "forget_fisher.pt", weights_only=False)
    rf = torch.load(mdir / "remain_fisher.pt", weights_only=False)
    assert list(ff.keys()) == pnames and list(rf.keys()) == pnames
    assert all(t.dtype == torch.float32 and t.device.type == "cpu" and list(t.shape) == shapes[k] for k, t in ff.items())
    for which, got, grads, count in (("forget", ff, taps[:nf], nf), ("remain", rf, taps[nf:], nr)):
        acc = {"w": torch.zeros(n)}
        for _, g in grads:
            O.fisher_accumulate_clipped(acc, {"w": g}, count, h["grad_clip"])          # :1270-1281
        assert close(flat_of(got, pnames), acc["w"], 2e-6), which
    # ---- DDPM/generate_fisher_mask.py on the files the runner wrote --------------------------------------
    ratio_path = generate_fisher_mask(str(mdir), 1.0)
    assert os.path.basename(ratio_path) == "fisher_1.0.pt"
    rmask = torch.load(ratio_path, weights_only=False)
    want, _, _ = O.ratio_mask(ff, rf, 1.0)
    assert all(rmask[k].dtype == torch.bool and torch.equal(rmask[k], want[k]) for k in pnames)
    # ---- --mode generate_mask (SalUn top-k) --------------------------------------------------------------
    taps = run(config, hooks, args(), "generate_mask", 32)
    tpath = tmp / "results" / "cifar10" / "mask" / "0" / "with_0.5.pt"
    tmask = torch.load(tpath, weights_only=False)
    assert list(tmask.keys()) == pnames and tmask[pnames[0]].dtype == torch.int64
    acc = torch.zeros(n)
    for _, g in taps:
        y = g.clone()
        O.clip_grad_norm([y], h["grad_clip"])                                           # :985-994
        acc += y
    assert torch.equal(flat_of(tmask, pnames), O.topk_mask_flat(acc, int(n * 0.5)).long())
    assert flat_of(tmask, pnames).shape == fx["topk"]["mask"].shape

    def check_ckpt(rec, ref, n_opt_steps):
        model_sd, opt_sd, step, ema_sd = torch.load(os.path.join(config.ckpt_dir, "ckpt.pth"), weights_only=False)
        assert list(model_sd.keys()) == rec["ckpt_model_keys"] and list(ema_sd.keys()) == rec["ckpt_ema_keys"]
        assert step == rec["step"] and opt_sd["param_groups"] == rec["ckpt_opt_param_groups"]
        st = opt_sd["state"]
        assert [float(st[i]["step"]) for i in range(len(names))] == rec["opt_steps"] == [float(n_opt_steps)] * len(names)
        assert close(flat_of(model_sd, pnames), ref.flat("p"))
        assert close(flat_of(ema_sd, names), ref.flat("slow"))
        assert close(torch.cat([st[i]["exp_avg"].reshape(-1) for i in range(len(names))]), ref.flat("m"))
        assert close(torch.cat([st[i]["exp_avg_sq"].reshape(-1) for i in range(len(names))]), ref.flat("v"))

    def oracle():
        return O.FlatReferenceLoop({"w": (n,)}, {"w": fx["theta0"]}, "adam",
                                   dict(lr=h["lr"], beta1=h["beta1"], eps=h["eps"], weight_decay=h["weight_decay"]),
                                   ema_mode="ddpm", ema_a=h["ema_rate"])

    # ---- --mode sfron: ron, adaptive gradient ascent, cosine-decayed forget alpha, the ratio mask ----------
    taps = run(config, hooks, args(mask_path=ratio_path, unlearn_loss="adaga", decay_forget_alpha=True,
                                   forget_alpha=5.0), "sfron_forget", 33)
    assert [k for k, _ in taps] == ["forget", "remain"] * h["n_iters"]
    ref, mask_flat = oracle(), flat_of(rmask, pnames)
    for (_, gf), (_, gr) in zip(taps[0::2], taps[1::2]):
        ref.forget_step({"w": gf}, mask={"w": mask_flat}, max_norm=h["grad_clip"])       # :1122-1138
        ref.remain_step({"w": gr}, max_norm=h["grad_clip"], ema=True)                    # :1156-1180
    check_ckpt(fx["sfron"], ref, 2 * h["n_iters"])
    # ---- --mode salun: joint loss, clip BEFORE the int64 top-k mask ---------------------------------------
    taps = run(config, hooks, args(mask_path=str(tpath), unlearn_loss="rl", forget_alpha=0.3), "saliency_unlearn", 34)
    assert [k for k, _ in taps] == ["joint"] * h["n_iters"]
    ref, mask_flat = oracle(), flat_of(tmask, pnames)
    for _, g in taps:
        ref.forget_step({"w": g}, mask={"w": mask_flat}, max_norm=h["grad_clip"], order="clip_then_mask")   # :576-590
        ref.slow_update()
    check_ckpt(fx["salun"], ref, h["n_iters"])


def test_ddpm_runner_save_fim_and_unsupported_modes(setup):
    """DDPM/fim.py: per-sample FIM through K1's row form == the oracle on the tapped per-sample gradients is covered
    by test_gpu_parity (golden ddpm_fim.pt); here the runner's file, keys and loop structure."""
    import pickle
    from sfron_b200.methods.ddpm import Diffusion
    fx, config, hooks, args, tmp, _ = setup
    config.diffusion.num_diffusion_timesteps = 8             # 8 timesteps in 2 chunks keeps the loop short
    torch.manual_seed(35)
    runner = Diffusion(args(n_chunks=2), config, hooks=hooks)
    runner.save_fim(batch_size=2)
    with open(tmp / "fisher_dict.pkl", "rb") as f:
        fim = pickle.load(f)
    assert list(fim.keys()) == fx["names"]
    total = torch.cat([t.reshape(-1) for t in fim.values()])
    assert bool((total >= 0).all()) and float(total.sum()) > 0
    with pytest.raises(NotImplementedError):
        Diffusion(args(method="joint"), config, hooks=hooks).sfron_forget()


# =========================================================================== DiT: forget.py / generate_fisher.py
class _TinyDiffusion:
    """Stand-in for create_diffusion(): the two members the DiT scripts use."""
    num_timesteps = 1000

    def __init__(self, device):
        betas = torch.linspace(1e-4, 0.02, 1000, dtype=torch.float64)
        self.ac = torch.cumprod(1 - betas, 0).float().to(device)

    def training_losses(self, model, x, t, model_kwargs):
        noise = torch.randn_like(x)
        a = self.ac[t].view(-1, 1, 1, 1)
        out = model(a.sqrt() * x + (1 - a).sqrt() * noise, t, **model_kwargs)
        return {"loss": (out[:, :x.shape[1]] - noise).square().mean(dim=(1, 2, 3))}


def _dit_setup(tmp_path):
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from dit_xl2 import DiTXL2Harness
    from sfron_b200.methods.dit import DiTHooks
    fx = load_golden("dit_scripts.pt")
    names = [n[len("module."):] for n in fx["names"]]
    shapes = {k[len("module."):]: v for k, v in fx["shapes"].items()}
    theta = unflat(fx["theta0"], names, shapes)

    def model_factory(args):
        model = DiTXL2Harness(input_size=8, patch=8, width=8, depth=1, heads=2, classes=10)
        model.load_state_dict(theta, strict=True)
        return model

    g = torch.Generator().manual_seed(5)
    forget = TensorDataset(torch.randn(4, 4, 8, 8, generator=g), torch.full((4,), 3))
    remain = TensorDataset(torch.randn(6, 4, 8, 8, generator=g), torch.randint(0, 10, (6,), generator=g))
    hooks = DiTHooks(model_factory=model_factory, diffusion_factory=lambda: _TinyDiffusion("cuda"),
                     unlearn_dataset=lambda args: (forget, remain), encode=lambda x, device: x.mul(0.18215))
    return fx, names, shapes, hooks


def test_dit_cli_generate_fisher_and_forget(tmp_path):
    """`generate_fisher` and `forget` through the reference's flags: files where DiT/generate_fisher.py:250,290 and
    DiT/forget.py:346-353 put them, structure == golden dit_scripts.pt, arithmetic == oracle on the tapped gradients."""
    from sfron_b200.methods import dit
    from sfron_b200.methods.masks import generate_mask_dit
    fx, names, shapes, hooks = _dit_setup(tmp_path)
    pnames, tnames = fx["names"], fx["train_names"]
    common = ["--data-path", "unused", "--results-dir", str(tmp_path / "results"), "--num-workers", "0",
              "--forget-class", "3", "--image-size", "256", "--num-classes", "10"]
    # ---- generate_fisher.py ----------------------------------------------------------------------------
    args = dit.generate_fisher_parser().parse_args(common + ["--n-iters", "3", "--mask-path", str(tmp_path / "mask")])
    taps = []
    out_dir = dit.generate_fisher_main(args, hooks, gradient_tap=lambda k, g: taps.append((k, g.float().cpu().clone())))
    assert out_dir == os.path.join(str(tmp_path / "mask"), "3") and [k for k, _ in taps] == ["forget"] * 3 + ["remain"] * 3
    assert os.path.isdir(tmp_path / "results" / "000-DiT-XL-2-fisher" / "checkpoints")
    n = taps[0][1].numel()
    for which, grads in (("forget", taps[:3]), ("remain", taps[3:])):
        d = torch.load(os.path.join(out_dir, f"{which}_fisher.pt"), weights_only=False)
        assert list(d.keys()) == pnames and d["module.pos_embed"] == 0               # frozen: the int-0 placeholder
        acc = torch.zeros(n)
        for _, g in grads:
            O.flat_fisher_accum(acc, g, 3)
        got = torch.cat([d[k].reshape(-1) for k in tnames])
        assert torch.equal(got.view(torch.int32), acc.view(torch.int32)), which        # K1 is bit-exact
    (mask_path,) = generate_mask_dit(str(tmp_path / "mask"), [3], [1.0])
    mask = torch.load(mask_path, weights_only=False)
    mask_flat = torch.cat([mask[k].reshape(-1) for k in tnames])
    # ---- forget.py ---------------------------------------------------------------------------------------
    args = dit.forget_parser().parse_args(common + ["--n-iters", "4", "--method", "ron", "--lr", "1e-4",
                                                    "--forget-alpha", "0.5", "--decay-forget-alpha",
                                                    "--mask-path", mask_path, "--log-every", "2"])
    taps = []
    path = dit.forget_main(args, hooks, gradient_tap=lambda k, g: taps.append((k, g.float().cpu().clone())))
    assert path == str(tmp_path / "results" / "001-DiT-XL-2-forget-3-ron-ga-lr0.0001-f0.5-r1.0" / "checkpoints" / "0000004.pt")
    assert [k for k, _ in taps] == ["forget", "remain"] * 4
    ck = torch.load(path, weights_only=False)
    rec = fx["forget"]
    assert list(ck.keys()) == rec["ckpt_keys"] and ck["args"].forget_class == 3
    assert list(ck["model"].keys()) == rec["model_keys"] and list(ck["ema"].keys()) == rec["ema_keys"]
    assert sorted(ck["opt"]["state"].keys()) == rec["opt_state_keys"]                # no slot for the frozen pos_embed
    theta0_train = torch.cat([unflat(fx["theta0"], names, shapes)[k[len("module."):]].reshape(-1) for k in tnames])
    ref = O.FlatReferenceLoop({"w": (n,)}, {"w": theta0_train}, "adamw", dict(lr=1e-4, weight_decay=0.0),
                              ema_mode="dit", ema_a=0.9999)
    for (_, gf), (_, gr) in zip(taps[0::2], taps[1::2]):
        ref.forget_step({"w": gf}, mask={"w": mask_flat}, max_norm=1.0)                 # forget.py:285-299
        ref.remain_step({"w": gr}, ema=True)                                            # :310-322
    got = torch.cat([ck["model"][k].reshape(-1) for k in tnames])
    assert close(got, ref.flat("p"))
    assert close(torch.cat([ck["ema"][k[len("module."):]].reshape(-1) for k in tnames]), ref.flat("slow"))
    assert [float(ck["opt"]["state"][i]["step"]) for i in sorted(ck["opt"]["state"])] == [8.0] * len(tnames)


# =========================================================================== SD: generate_fisher.py / nsfw_removal.py
class _TinyLatentDiffusion(torch.nn.Module):
    """The five members of LatentDiffusion the SD scripts call, around a 586-parameter U-Net with an `attn2`."""
    first_stage_key = "jpg"
    num_timesteps = 1000

    def __init__(self, theta, device):
        super().__init__()
        from helpers_models import TinyLatentUNet
        self.model = torch.nn.Module()
        self.model.diffusion_model = TinyLatentUNet()
        self.model.diffusion_model.load_state_dict(theta, strict=True)
        self.device = torch.device(device)
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
        self.register_buffer("ac", torch.cumprod(1 - betas, 0).float())
        self.to(self.device)

    def get_input(self, batch, key):
        x = batch[key].permute(0, 3, 1, 2).to(self.device)[:, :, :8, :8].mean(dim=1, keepdim=True).repeat(1, 4, 1, 1)
        seed = float(sum(map(ord, batch["txt"][0]))) if batch["txt"][0] else 0.0
        ctx = torch.full((x.shape[0], 8), seed * 1e-3, device=self.device) + torch.arange(8, device=self.device) * 0.1
        return x, ctx

    def q_sample(self, x_start, t, noise):
        a = self.ac[t].view(-1, 1, 1, 1)
        return a.sqrt() * x_start + (1 - a).sqrt() * noise

    def apply_model(self, x, t, ctx):
        u = self.model.diffusion_model
        h = u.conv_in(x) + u.temb(t.float()[:, None] / 1000.0)[:, :, None, None]
        q = u.attn2.to_q(h.mean(dim=(2, 3)))
        h = h + (q * u.attn2.to_k(ctx) * u.attn2.to_v(ctx))[:, :, None, None]
        return u.conv_out(torch.tanh(h))

    def shared_step(self, batch):
        x, ctx = self.get_input(batch, self.first_stage_key)
        t = torch.randint(0, self.num_timesteps, (x.shape[0],), device=self.device).long()
        noise = torch.randn_like(x)
        loss = (self.apply_model(self.q_sample(x, t, noise), t, ctx) - noise).square().mean()
        return loss, {}


def test_sd_cli_generate_fisher_and_nsfw_removal(tmp_path, monkeypatch):
    from sfron_b200.methods import sd
    from sfron_b200.methods.masks import generate_fisher_mask
    fx = load_golden("sd_scripts.pt")
    names, shapes = fx["names"], fx["shapes"]
    theta = unflat(fx["theta0"], names, shapes)
    monkeypatch.chdir(tmp_path)                               # both scripts write cwd-relative (fisher/, models/)
    g = torch.Generator().manual_seed(8)
    forget_dl = DataLoader(torch.rand(4, 3, 16, 16, generator=g), batch_size=2)
    remain_dl = DataLoader(torch.rand(6, 3, 16, 16, generator=g), batch_size=2)
    hooks = sd.SDHooks(setup_model=lambda cfg, ckpt, device: _TinyLatentDiffusion(theta, device),
                       setup_data=lambda bs, size: (forget_dl, remain_dl))
    a = sd.generate_fisher_parser().parse_args(["--batch_size", "2", "--c_guidance", "7.5"])
    taps = []
    torch.manual_seed(1)
    sd.generate_nsfw_fisher(a.c_guidance, a.batch_size, a.epochs, a.lr, a.config_path, a.ckpt_path, a.diffusers_config_path,
                            "cuda:0", a.image_size, a.num_timesteps, hooks=hooks,
                            gradient_tap=lambda k, gr: taps.append((k, gr.float().cpu().clone())))
    assert [k for k, _ in taps] == ["forget"] * 2 + ["remain"] * 3
    n = taps[0][1].numel()
    for which, fname, grads in (("forget", "nude_forget.pt", taps[:2]), ("remain", "nude_remain.pt", taps[2:])):
        d = torch.load(tmp_path / "fisher" / fname, weights_only=False)
        assert list(d.keys()) == names                        # U-Net-local keys, no prefix
        acc = torch.zeros(n)
        for _, gr in grads:
            O.flat_fisher_accum(acc, gr, len(grads))
        assert torch.equal(torch.cat([d[k].reshape(-1) for k in names]).view(torch.int32), acc.view(torch.int32)), which
    mpath = generate_fisher_mask(str(tmp_path / "fisher"), 1.0, forget_name="nude_forget.pt",
                                 remain_name="nude_remain.pt", out_fmt="nude_mask_{th}.pt")
    mask = torch.load(mpath, weights_only=False)
    # ---- nsfw_removal.py: xattn (only attn2 trains), as the reference behaves (mask never applied) and as intended ----
    a = sd.nsfw_removal_parser().parse_args(["--train_method", "xattn", "--batch_size", "2", "--n_iters", "3", "--lr", "1e-3",
                                             "--forget_alpha", "0.7", "--mask_path", str(tmp_path / "fisher"),
                                             "--mask_threshold", "1.0"])
    train = [k for k in names if "attn2" in k]
    theta0_train = torch.cat([theta[k].reshape(-1) for k in train])
    for apply_mask in (False, True):
        taps = []
        torch.manual_seed(2)
        path = sd.nsfw_removal(a.train_method, a.forget_alpha, a.remain_alpha, a.batch_size, a.n_iters, a.lr, a.config_path,
                               a.ckpt_path, a.mask_path, a.diffusers_config_path, "cuda:0", a.mask_threshold, a.image_size,
                               a.ddim_steps, hooks=hooks, apply_mask=apply_mask,
                               gradient_tap=lambda k, gr: taps.append((k, gr.float().cpu().clone())))
        assert path == "models/compvis-nsfw-mask1.0-method_sfron-lr0.001_fa0.7_ra1.0/compvis-nsfw-mask1.0-method_sfron-lr0.001_fa0.7_ra1.0.pt"
        sdict = torch.load(tmp_path / path, weights_only=False)
        assert [k for k in sdict if k.startswith("model.diffusion_model.")] == ["model.diffusion_model." + k for k in names]
        ref = O.FlatReferenceLoop({"w": (theta0_train.numel(),)}, {"w": theta0_train}, "adam", dict(lr=1e-3))
        mflat = torch.cat([mask[k].reshape(-1) for k in train])
        for (_, gf), (_, gr) in zip(taps[0::2], taps[1::2]):
            ref.forget_step({"w": gf}, mask={"w": mflat} if apply_mask else None)       # nsfw_removal.py:157-162
            ref.remain_step({"w": gr}, ema=False)                                       # :165-173
        got = torch.cat([sdict["model.diffusion_model." + k].reshape(-1) for k in train])
        assert close(got, ref.flat("p")), apply_mask
        for k in names:                                       # everything outside the cross-attention is untouched
            if k not in train:
                assert torch.equal(sdict["model.diffusion_model." + k].cpu(), theta[k])
