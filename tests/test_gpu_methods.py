"""GPU tests of the drop-in entry points (sfron_b200.methods) against the reference-generated
golden fixtures: same class / CLI surface, same files, same results."""
import argparse
import os

import pytest
import torch
import torch.nn as nn

from conftest import load_golden, unflat
from helpers_models import TinyCond, TinyDiT, TinyLatentUNet, TinyNet, inject, loaders

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def close(a, b, rtol=1e-6):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    rms = b.pow(2).mean().sqrt().item()
    return bool(((a - b).abs() <= rtol * (b.abs() + rms)).all())


def set_flat(model, flat, names, shapes):
    vals = unflat(flat, names, shapes)
    with torch.no_grad():
        for n, p in model.named_parameters():
            p.copy_(vals[n])


@pytest.mark.parametrize("tag", ["default", "beta09"])
def test_classification_sfron_dropin(dev, tmp_path, tag):
    """The whole method (Fisher -> mask -> loop) with the real model forward/backward on the GPU,
    against the reference's CPU run: GPU conv/matmul rounding differs from CPU's, so the comparison is
    looser than the kernel-level parity tests (which replay identical gradients)."""
    from sfron_b200.methods import create_unlearn_method
    fx = load_golden(f"cls_sfron_{tag}.pt")
    h = fx["hyper"]
    torch.manual_seed(0)
    model = TinyNet()
    set_flat(model, fx["theta0"], fx["names"], fx["shapes"])
    model = model.to(dev)
    args = argparse.Namespace(num_classes=10, seed=0)
    method = create_unlearn_method("SFRon")(model, nn.CrossEntropyLoss(), str(tmp_path), args)
    method.n_iters, method.forget_freq, method.log_freq = h["n_iters"], h["forget_freq"], 10 ** 9
    method.ema_beta = h["ema_beta"]
    method.prepare_unlearn(loaders(1))
    # files in the reference's format
    for which in ("forget", "remain"):
        d = torch.load(tmp_path / f"{which}_fisher.pt", weights_only=False)
        assert list(d.keys()) == fx["names"]
        got = torch.cat([d[n].reshape(-1) for n in fx["names"]])
        assert got.dtype == torch.float32 and got.device.type == "cpu"
        assert torch.allclose(got, fx[f"{which}_fisher"], rtol=1e-3, atol=1e-9)
    mask = method.weight_saliency_mask
    got_mask = torch.cat([mask[n].reshape(-1) for n in fx["names"]])
    assert got_mask.dtype == torch.bool
    assert (got_mask.to(torch.uint8) == fx["mask"]).float().mean() > 0.99
    out = method.get_unlearned_model()
    assert out is model
    final = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu()
    assert torch.allclose(final, fx["theta_final"], rtol=0, atol=5e-3)
    assert (final - fx["theta0"]).abs().max() > 1e-3          # it did move
    assert set(method.get_params()) >= {"opt", "momentum", "weight_decay", "retain_lr", "n_iters", "threshold"}
    # second construction reuses the cached Fisher files (sfron.py:269-271,296-298)
    m2 = create_unlearn_method("SFRon")(model, nn.CrossEntropyLoss(), str(tmp_path), args)
    m2.prepare_unlearn(loaders(1))
    got2 = torch.cat([m2.weight_saliency_mask[n].reshape(-1) for n in fx["names"]])
    assert torch.equal(got2, got_mask)
    with pytest.raises(NotImplementedError):
        create_unlearn_method("SCRUB")


@pytest.mark.parametrize("forget_freq", [1, 3])
def test_classification_loop_replayed_from_cuda_graphs(dev, tmp_path, forget_freq):
    """`args.cuda_graph`: every iteration (forward, backward, kernels) replayed from a CUDA graph, the per-iteration
    cosine learning rate read on the device from torch's own scheduler table, the optimizer step counted on the
    device.  Same model, data and seeds as the eager loop -> the same weights (sfron.py:151-260)."""
    from sfron_b200.methods import create_unlearn_method
    fx = load_golden("cls_sfron_default.pt")
    finals = []
    for graphed in (False, True):
        torch.manual_seed(0)
        model = TinyNet()
        set_flat(model, fx["theta0"], fx["names"], fx["shapes"])
        model = model.to(dev)
        args = argparse.Namespace(num_classes=10, seed=0, cuda_graph=graphed)
        out_dir = tmp_path / f"g{int(graphed)}"
        out_dir.mkdir()
        method = create_unlearn_method("SFRon")(model, nn.CrossEntropyLoss(), str(out_dir), args)
        method.n_iters, method.forget_freq, method.log_freq, method.ema_beta = 12, forget_freq, 10 ** 9, 0.9
        method.prepare_unlearn(loaders(1))
        method.get_unlearned_model()
        torch.cuda.synchronize()
        finals.append(torch.cat([p.detach().reshape(-1) for p in model.parameters()]).cpu())
        if graphed:
            hp = method._hot_path().hp
            steps = 12 + len([s for s in range(12) if s % forget_freq == 0])
            assert int(hp.step_dev) == steps and int(hp.lr_index) == 12
    assert (finals[0] - fx["theta0"]).abs().max() > 1e-3
    assert close(finals[1], finals[0], 1e-6)


def test_salun_topk_dropin(dev, tmp_path):
    from sfron_b200.methods import create_unlearn_method
    fx = load_golden("salun_topk.pt")
    for th in ("0.2", "0.5"):
        torch.manual_seed(3)
        model = TinyNet().to(dev)
        args = argparse.Namespace(num_classes=10, seed=0, batch_size=4)
        method = create_unlearn_method("SalUn")(model, nn.CrossEntropyLoss(), str(tmp_path), args)
        method.th = float(th)
        hard = method.get_gradient_ratio(loaders(2)["forget_train"])
        got = torch.cat([hard[n].reshape(-1) for n in fx["names"]])
        assert got.dtype == torch.int64 and int(got.sum()) == int(got.numel() * float(th))
        assert (got == fx[th]["mask"]).float().mean() > 0.99     # GPU vs CPU backward rounding


def test_mask_clis_match_reference_scripts(dev, tmp_path):
    from sfron_b200.methods import masks
    # DDPM / SD: --ckpt_folder / --threshold, outputs fisher_{th}.pt / nude_mask_{th}.pt
    for fam, fixture, fn, rn, out in (("ddpm", "ddpm_ratio_mask.pt", "forget_fisher.pt", "remain_fisher.pt", "fisher_{}.pt"),
                                      ("sd", "sd_ratio_mask.pt", "nude_forget.pt", "nude_remain.pt", "nude_mask_{}.pt")):
        fx = load_golden(fixture)
        folder = tmp_path / fam
        folder.mkdir()
        torch.save(fx["forget"], folder / fn)
        torch.save(fx["remain"], folder / rn)
        for th_s, ref in fx["masks"].items():
            masks.main([fam, "--ckpt_folder", str(folder), "--threshold", th_s])
            got = torch.load(folder / out.format(float(th_s)), weights_only=False)
            assert list(got.keys()) == list(ref.keys())
            assert all(got[n].dtype == torch.bool and torch.equal(got[n], ref[n]) for n in ref)
    # DiT: --mask-path / --forget-class / --thresholds, int-0 placeholder preserved
    fx = load_golden("dit_ratio_mask.pt")
    folder = tmp_path / "dit" / "7"
    folder.mkdir(parents=True)
    torch.save(fx["forget"], folder / "forget_fisher.pt")
    torch.save(fx["remain"], folder / "remain_fisher.pt")
    masks.generate_mask_dit(str(tmp_path / "dit"), [7], [0.5, 1, 3, 5, 10])
    for th_s, ref in fx["masks"].items():
        got = torch.load(folder / f"fisher_{th_s}.pt", weights_only=False)
        assert list(got.keys()) == list(ref.keys())
        for n, r in ref.items():
            if torch.is_tensor(r):
                assert torch.equal(got[n], r)
            else:
                assert got[n] == 0 and not torch.is_tensor(got[n])


def test_dit_family_loop_and_checkpoint(dev):
    from sfron_b200.methods.diffusion import DiffusionUnlearner
    fx = load_golden("dit_adamw_ema_loop.pt")
    names, shapes = [n[len("module."):] for n in fx["names"]], {k[len("module."):]: v for k, v in fx["shapes"].items()}
    model = TinyDiT()
    set_flat(model, fx["theta0"], names, shapes)
    model = model.to(dev)
    un = DiffusionUnlearner(model, "dit", lr=fx["hyper"]["lr"])
    tnames = fx["train_names"]
    tshapes = {n: fx["shapes"][n] for n in tnames}
    un.load_mask(unflat(fx["mask"], tnames, tshapes))          # keys carry "module.", like the reference file
    gf, gr = fx["forget_grads"], fx["remain_grads"]
    un.forget(len(gf), lambda i: inject(model, gf[i]), lambda i: inject(model, gr[i]))
    final = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert close(final, fx["theta_final"])
    ckpt = un.checkpoint(step=len(gf), args=argparse.Namespace(lr=1e-4))
    assert set(ckpt) == {"model", "ema", "opt", "args"}
    assert list(ckpt["model"].keys()) == fx["names"]                # DataParallel-prefixed, as model.state_dict()
    ema = torch.cat([ckpt["ema"][n].reshape(-1) for n in names])     # ema = deepcopy(model): bare names
    assert close(ema, fx["ema_final"])
    # the optimizer slot loads into the reference's optimizer
    ref_model = nn.DataParallel(TinyDiT())
    opt = torch.optim.AdamW(ref_model.parameters(), lr=1e-4, weight_decay=0)
    opt.load_state_dict(ckpt["opt"])
    assert float(opt.state[ref_model.module.fc.weight]["step"]) == 2 * len(gf)


def test_ddpm_family_loop_fisher_and_topk(dev, tmp_path):
    from sfron_b200.methods.diffusion import DiffusionUnlearner
    fx = load_golden("ddpm_adam_ema_loop.pt")
    h = fx["hyper"]
    names = [n[len("module."):] for n in fx["names"]]
    shapes = {k[len("module."):]: v for k, v in fx["shapes"].items()}
    model = TinyNet()
    set_flat(model, fx["theta0"], names, shapes)
    model = model.to(dev)
    un = DiffusionUnlearner(model, "ddpm", lr=h["lr"])
    un.load_mask({n: m for n, m in unflat(fx["mask"], fx["names"], fx["shapes"]).items()})
    gf, gr = fx["forget_grads"], fx["remain_grads"]
    un.forget(len(gf), lambda i: inject(model, gf[i]), lambda i: inject(model, gr[i]))
    final = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert close(final, fx["theta_final"])
    states = un.checkpoint(step=5)
    assert isinstance(states, list) and len(states) == 4 and states[2] == 5
    assert close(torch.cat([states[3][n].reshape(-1) for n in names]), fx["ema_final"])
    assert close(torch.cat([states[1]["state"][i]["exp_avg_sq"].reshape(-1) for i in range(4)]), fx["exp_avg_sq"])
    # DDPM Fisher: clipped gradient, files with module.-prefixed keys
    from oracle import sfron_oracle as O
    acc = O.fisher_init(fx["names"])
    for g in gf:
        O.fisher_accumulate_clipped(acc, unflat(g, fx["names"], fx["shapes"]), len(gf), h["grad_clip"])
    un.generate_fisher("forget", len(gf), lambda i: inject(model, gf[i]), out_dir=str(tmp_path))
    d = torch.load(tmp_path / "forget_fisher.pt", weights_only=False)
    assert list(d.keys()) == fx["names"]
    assert close(torch.cat([d[n].reshape(-1) for n in fx["names"]]), torch.cat([acc[n].reshape(-1) for n in fx["names"]]))
    # SalUn top-k mask of the DDPM runner on the golden gradients (no clip: family preset overridden)
    sx = load_golden("salun_topk.pt")
    un2 = DiffusionUnlearner(TinyNet().to(dev), "ddpm", clip_fisher=None)
    grads = sx["0.5"]["grads"]
    hard = un2.generate_topk_mask(len(grads), lambda i: inject(un2.model, grads[i]), ratio=0.5,
                                  path=str(tmp_path / "mask" / "with_0.5.pt"))
    saved = torch.load(tmp_path / "mask" / "with_0.5.pt", weights_only=False)
    got = torch.cat([saved["module." + n].reshape(-1) for n in sx["names"]])
    assert got.dtype == torch.int64 and torch.equal(got, sx["0.5"]["mask"])
    # per-sample FIM rows
    rows = torch.randn(3, un2.mhp.layout.numel + 2, device=dev)[:, :un2.mhp.layout.numel]
    fim = un2.save_fim([rows[:2], rows[2:]], dataset_len=10, path=str(tmp_path / "fisher_dict.pkl"))
    want = torch.zeros(un2.mhp.layout.numel)
    for r in rows.cpu():
        want += r ** 2 / 10
    assert torch.equal(fim.cpu(), want)
    import pickle
    with open(tmp_path / "fisher_dict.pkl", "rb") as f:
        assert list(pickle.load(f).keys()) == ["module." + n for n in sx["names"]]


def test_ddpm_runner_dropin_against_whole_reference_methods(dev, tmp_path):
    """The DDPM family of DiffusionUnlearner against `Diffusion.generate_fisher`, `generate_mask`,
    `sfron_forget` and `saliency_unlearn` EXECUTED WHOLE from the reference (fixture ddpm_runner.pt):
    the recorded raw gradient of every reference backward pass is replayed through a real backward."""
    from sfron_b200.methods.diffusion import DiffusionUnlearner
    from sfron_b200.methods.masks import generate_fisher_mask
    fx = load_golden("ddpm_runner.pt")
    h, pnames = fx["hyper"], fx["names"]
    names = [n[len("module."):] for n in pnames]
    shapes = {k[len("module."):]: v for k, v in fx["shapes"].items()}

    def fresh():
        model = TinyCond()
        assert [n for n, _ in model.named_parameters()] == names
        set_flat(model, fx["theta0"], names, shapes)
        return DiffusionUnlearner(model.to(dev), "ddpm", lr=h["lr"])

    def flat_of(d):
        return torch.cat([d[n].reshape(-1) for n in pnames])

    # --mode generate_fisher: Fisher of the clipped batch gradient, reference file names and keys
    un, fi = fresh(), fx["fisher"]
    un.generate_fisher("forget", len(fi["forget_grads"]), lambda i: inject(un.model, fi["forget_grads"][i]), out_dir=str(tmp_path))
    un.generate_fisher("remain", len(fi["remain_grads"]), lambda i: inject(un.model, fi["remain_grads"][i]), out_dir=str(tmp_path))
    ff = torch.load(tmp_path / "forget_fisher.pt", weights_only=False)
    rf = torch.load(tmp_path / "remain_fisher.pt", weights_only=False)
    assert list(ff.keys()) == pnames and list(rf.keys()) == pnames
    assert close(flat_of(ff), fi["forget_fisher"]) and close(flat_of(rf), fi["remain_fisher"])
    # generate_fisher_mask.py on the REFERENCE's Fisher files -> bit-exact bool mask, same file name
    torch.save(unflat(fi["forget_fisher"], pnames, fx["shapes"]), tmp_path / "forget_fisher.pt")
    torch.save(unflat(fi["remain_fisher"], pnames, fx["shapes"]), tmp_path / "remain_fisher.pt")
    path = generate_fisher_mask(str(tmp_path), 1.0)
    assert os.path.basename(path) == "fisher_1.0.pt"
    rmask = torch.load(path, weights_only=False)
    assert rmask[pnames[0]].dtype == torch.bool
    assert torch.equal(flat_of(rmask).to(torch.uint8), fx["ratio_mask"])
    # --mode generate_mask: SalUn top-k of |sum of clipped gradients| -> int64 0/1, bit-exact
    un, tk = fresh(), fx["topk"]
    hard = un.generate_topk_mask(len(tk["grads"]), lambda i: inject(un.model, tk["grads"][i]), ratio=tk["ratio"],
                                 path=str(tmp_path / "results" / "with_0.5.pt"))
    got = flat_of(hard)
    assert got.dtype == torch.int64 and torch.equal(got.cpu(), tk["mask"])

    def check_final(un, rec, n_steps):
        final = torch.cat([p.detach().reshape(-1) for p in un.model.parameters()])
        assert close(final, rec["theta"])
        states = un.checkpoint(step=rec["step"])
        assert list(states[0].keys()) == rec["ckpt_model_keys"] and list(states[3].keys()) == rec["ckpt_ema_keys"]
        assert close(torch.cat([states[3][n].reshape(-1) for n in names]), rec["ema"])
        st = states[1]["state"]
        assert close(torch.cat([st[i]["exp_avg"].reshape(-1) for i in range(len(names))]), rec["exp_avg"])
        assert close(torch.cat([st[i]["exp_avg_sq"].reshape(-1) for i in range(len(names))]), rec["exp_avg_sq"])
        assert [float(st[i]["step"]) for i in range(len(names))] == rec["opt_steps"] == [float(n_steps)] * len(names)
        assert states[1]["param_groups"] == rec["ckpt_opt_param_groups"]      # loadable by the reference's optimizer

    # --mode sfron (ron): alpha_t is already inside the recorded forget gradients
    un, rec = fresh(), fx["sfron"]
    un.load_mask(path)
    gf, gr = rec["grads"][0::2], rec["grads"][1::2]
    assert rec["kinds"][0::2] == ["forget"] * len(gf)
    un.forget(len(gf), lambda i: inject(un.model, gf[i]), lambda i: inject(un.model, gr[i]))
    check_final(un, rec, 2 * len(gf))
    # the same loop with the WHOLE iteration captured in a CUDA graph: static gradient buffers refilled per replay
    un = fresh()
    un.load_mask(path)
    gf_s, gr_s = torch.zeros_like(gf[0], device=dev), torch.zeros_like(gr[0], device=dev)

    def refill(i):
        gf_s.copy_(gf[i], non_blocking=False)
        gr_s.copy_(gr[i], non_blocking=False)

    un.forget(len(gf), lambda i: inject(un.model, gf_s), lambda i: inject(un.model, gr_s), cuda_graph=True, refill=refill)
    check_final(un, rec, 2 * len(gf))
    # --mode saliency_unlearn: joint loss, clip BEFORE the int64 top-k mask
    un, rec = fresh(), fx["salun"]
    un.load_mask(unflat(fx["topk"]["mask"], pnames, fx["shapes"]))
    un.saliency_unlearn(len(rec["grads"]), lambda i: inject(un.model, rec["grads"][i]))
    check_final(un, rec, len(rec["grads"]))


def test_dit_scripts_dropin_against_whole_reference_scripts(dev, tmp_path):
    """The DiT family against DiT/generate_fisher.py, generate_mask.py and forget.py EXECUTED WHOLE (fixture
    dit_scripts.pt): Fisher files, the mask CLI, the ron loop (eager and from one CUDA graph) and the checkpoint."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from dit_xl2 import DiTXL2Harness
    from sfron_b200.methods.diffusion import DiffusionUnlearner
    from sfron_b200.methods.masks import generate_mask_dit
    fx = load_golden("dit_scripts.pt")
    pnames, tnames = fx["names"], fx["train_names"]
    names = [n[len("module."):] for n in pnames]
    shapes = {k[len("module."):]: v for k, v in fx["shapes"].items()}

    def fresh():
        model = DiTXL2Harness(input_size=8, patch=8, width=8, depth=1, heads=2, classes=10)
        assert [(n, list(p.shape)) for n, p in model.named_parameters()] == [(n, shapes[n]) for n in names]
        set_flat(model, fx["theta0"], names, shapes)
        return DiffusionUnlearner(model.to(dev), "dit", lr=fx["forget"]["hyper"]["lr"])

    un, fi = fresh(), fx["fisher"]
    fdir = tmp_path / "mask" / "3"
    for which in ("forget", "remain"):
        grads = fi[f"{which}_grads"]
        un.generate_fisher(which, len(grads), lambda i: inject(un.model, grads[i]), out_dir=str(fdir))
        d = torch.load(fdir / f"{which}_fisher.pt", weights_only=False)
        assert list(d.keys()) == pnames and d["module.pos_embed"] == 0
        got = torch.cat([d[n].reshape(-1) for n in tnames])
        assert torch.equal(got.view(torch.int32), fi[f"{which}_fisher"].view(torch.int32)), which     # K1 is bit-exact
    # the same Fisher loop with forward + backward + K1 replayed from one CUDA graph: still bit-exact
    ung = fresh()
    g_s = torch.zeros_like(fi["forget_grads"][0], device=dev)
    ung.generate_fisher("forget", len(fi["forget_grads"]), lambda i: inject(ung.model, g_s), cuda_graph=True,
                        refill=lambda i: g_s.copy_(fi["forget_grads"][i]))
    assert torch.equal(ung.mhp.hp.forget_fisher.cpu().view(torch.int32), fi["forget_fisher"].view(torch.int32))
    (path,) = generate_mask_dit(str(tmp_path / "mask"), [3], [1.0])
    assert os.path.basename(path) == "fisher_1.0.pt"
    mask = torch.load(path, weights_only=False)
    assert mask["module.pos_embed"] == 0 and mask[tnames[0]].dtype == torch.bool
    assert torch.equal(torch.cat([mask[n].reshape(-1) for n in tnames]).to(torch.uint8), fx["ratio_mask"])

    rec = fx["forget"]
    gf, gr = rec["grads"][0::2], rec["grads"][1::2]
    for graphed in (False, True):
        un = fresh()
        un.load_mask(path)
        if graphed:
            gf_s, gr_s = torch.zeros_like(gf[0], device=dev), torch.zeros_like(gr[0], device=dev)

            def refill(i):
                gf_s.copy_(gf[i])
                gr_s.copy_(gr[i])
            un.forget(len(gf), lambda i: inject(un.model, gf_s), lambda i: inject(un.model, gr_s), cuda_graph=True, refill=refill)
        else:
            un.forget(len(gf), lambda i: inject(un.model, gf[i]), lambda i: inject(un.model, gr[i]))
        final = torch.cat([p.detach().reshape(-1) for p in un.model.parameters()])          # frozen pos_embed included
        assert close(final, rec["theta"]), graphed
        ck = un.checkpoint(step=len(gf), args=argparse.Namespace(lr=1e-4))
        assert list(ck.keys()) == rec["ckpt_keys"]
        assert list(ck["model"].keys()) == rec["model_keys"] and list(ck["ema"].keys()) == rec["ema_keys"]
        assert close(torch.cat([ck["ema"][n].reshape(-1) for n in names]), rec["ema"]), graphed
        st = ck["opt"]["state"]
        assert sorted(st.keys()) == rec["opt_state_keys"]                                    # no slot for frozen pos_embed
        assert [float(st[i]["step"]) for i in sorted(st)] == rec["opt_steps"]
        assert close(torch.cat([st[i]["exp_avg"].reshape(-1) for i in sorted(st)]), rec["exp_avg"])
        assert close(torch.cat([st[i]["exp_avg_sq"].reshape(-1) for i in sorted(st)]), rec["exp_avg_sq"])


def test_sd_scripts_dropin_against_whole_reference_scripts(dev, tmp_path):
    """The SD family against generate_fisher.py / generate_fisher_mask.py / nsfw_removal.py / gradient_ascent.py
    EXECUTED WHOLE (fixture sd_scripts.pt): U-Net-local keys, Adam, no clip, no EMA."""
    from sfron_b200.methods.diffusion import DiffusionUnlearner
    from sfron_b200.methods.masks import generate_fisher_mask
    fx = load_golden("sd_scripts.pt")
    names, shapes, fi = fx["names"], fx["shapes"], fx["fisher"]

    def fresh(lr):
        model = TinyLatentUNet()
        assert [n for n, _ in model.named_parameters()] == names
        set_flat(model, fx["theta0"], names, shapes)
        return DiffusionUnlearner(model.to(dev), "sd", lr=lr)

    un = fresh(1e-5)
    fdir = tmp_path / "fisher"
    for which, fname in (("forget", "nude_forget.pt"), ("remain", "nude_remain.pt")):
        grads = fi[f"{which}_grads"]
        un.generate_fisher(which, len(grads), lambda i: inject(un.model, grads[i]), out_dir=str(fdir))
        d = torch.load(fdir / fname, weights_only=False)
        assert list(d.keys()) == names                                         # no "module." / "model.diffusion_model." prefix
        got = torch.cat([d[n].reshape(-1) for n in names])
        assert torch.equal(got.view(torch.int32), fi[f"{which}_fisher"].view(torch.int32)), which
    path = generate_fisher_mask(str(fdir), 1.0, forget_name="nude_forget.pt", remain_name="nude_remain.pt",
                                out_fmt="nude_mask_{th}.pt")
    assert os.path.basename(path) == "nude_mask_1.0.pt"
    mask = torch.load(path, weights_only=False)
    assert torch.equal(torch.cat([mask[n].reshape(-1) for n in names]).to(torch.uint8), fx["ratio_mask"])
    # nsfw_removal.py as it actually behaves: its `n in parameters` guard never fires, so no mask is applied
    rec = fx["nsfw_removal"]
    un = fresh(rec["lr"])
    un.load_mask(path)
    gf, gr = rec["grads"][0::2], rec["grads"][1::2]
    un.forget(len(gf), lambda i: inject(un.model, gf[i]), lambda i: inject(un.model, gr[i]), use_mask=rec["mask_applied"])
    assert close(torch.cat([p.detach().reshape(-1) for p in un.model.parameters()]), rec["theta"])
    # gradient_ascent.py: one masked Adam step per iteration on the joint loss (saliency_unlearn without clip / EMA)
    rec = fx["gradient_ascent"]
    un = fresh(rec["lr"])
    un.load_mask(path)
    un.saliency_unlearn(len(rec["grads"]), lambda i: inject(un.model, rec["grads"][i]), use_mask=rec["mask_applied"])
    assert close(torch.cat([p.detach().reshape(-1) for p in un.model.parameters()]), rec["theta"])
    sd = un.checkpoint()
    assert list(sd.keys()) == names                                            # plain state_dict, as save_model writes


def test_ddpm_sa_forget_ewc_dropin(dev):
    """Selective-Amnesia step of the DDPM runner executed whole (fixture ddpm_sa_forget.pt): base-loss backward, then
    `add_ewc_penalty` (fused penalty + gradient kernel) instead of the per-tensor penalty graph, clip, Adam, EMA."""
    from sfron_b200.methods.diffusion import DiffusionUnlearner
    fx = load_golden("ddpm_sa_forget.pt")
    h, pnames = fx["hyper"], fx["names"]
    names = [n[len("module."):] for n in pnames]
    shapes = {k[len("module."):]: v for k, v in fx["shapes"].items()}
    model = TinyCond()
    set_flat(model, fx["theta0"], names, shapes)
    un = DiffusionUnlearner(model.to(dev), "ddpm", lr=h["lr"])
    un.snapshot_params("params_mle")                                   # params_mle_dict[name] = param.data.clone()
    un.mhp.hp.set_buffer("fim", fx["fisher"].to(dev))                  # fisher_dict.pkl
    un.mhp.zero_grad()
    for step, g_ref in enumerate(fx["grads"]):
        inject(un.model, fx["base_grad"]).backward()                   # (1 + gamma) * c: the stand-in diffusion loss
        penalty = un.add_ewc_penalty(h["lmbda"])
        assert (float(penalty) == 0.0) == (step == 0)
        assert close(un.mhp.grads(), g_ref), step                      # == what the reference's single backward produced
        un.mhp.joint_step(use_mask=False, max_norm=h["grad_clip"], ema=True)
    assert close(torch.cat([p.detach().reshape(-1) for p in un.model.parameters()]), fx["theta"])
    assert close(torch.cat([un.mhp.slow_state_dict()["module." + n].reshape(-1) for n in names]), fx["ema"])


def test_bf16_model_mixed_precision_flat_params(dev):
    """BASELINE config 3 (bf16): module weights / grads are bf16 views, the kernels keep an fp32 master."""
    import sfron_b200 as sfr
    from sfron_b200.methods.common import ModelHotPath
    from oracle import sfron_oracle as O
    torch.manual_seed(0)
    model = TinyNet().to(dev).to(torch.bfloat16)
    theta0 = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()]).cpu()
    mhp = ModelHotPath(model, sfr.OptConfig(kind="adamw", lr=1e-3), ema_mode="dit", ema_a=0.999)
    fp = mhp.flat
    assert fp.p.dtype == torch.float32 and fp.g.dtype == torch.bfloat16 and fp.p_work.dtype == torch.bfloat16
    assert next(model.parameters()).data_ptr() == fp.p_work.data_ptr()
    n = fp.n
    ref = O.FlatReferenceLoop({"w": (n,)}, {"w": theta0}, "adamw", dict(lr=1e-3), ema_mode="dit", ema_a=0.999)
    mhp.hp.mask.fill_(1)
    mhp.hp.mark_mask_ready()
    g = torch.Generator().manual_seed(1)
    for _ in range(3):
        grad = (torch.randn(n, generator=g) * 0.1).bfloat16()
        inject(model, grad.float()).backward()           # bf16 params -> bf16 grads land in the flat g
        assert torch.equal(fp.g.cpu(), grad)
        mhp.remain_step(max_norm=1.0)
        ref.remain_step({"w": grad.float()}, max_norm=1.0)
        assert int(fp.g.count_nonzero()) == 0             # fused zero_grad
    assert close(fp.p, ref.flat("p"))                     # fp32 master follows the fp32 oracle
    assert torch.equal(fp.p_work, fp.p.bfloat16())        # working copy = rounded master
    assert torch.equal(next(model.parameters()).reshape(-1), fp.p_work[:216])


def test_padded_flat_params_views_and_length(dev):
    import sfron_b200 as sfr
    model = TinyNet().to(dev)
    fp = sfr.FlatParams(model, dev, pad_multiple=128)
    assert fp.n == 554 and fp.n_padded == 640 and fp.p.numel() == 554
    assert fp.p.data_ptr() == fp.p_padded.data_ptr() and int(fp.p_padded[554:].abs().sum()) == 0


def test_per_sample_fim_vmap_producer_matches_retain_graph_loop(dev):
    """the vmapped per-sample-gradient producer + K1 rows against the reference's
    `loss[i].backward(retain_graph=True)` loop and `F += tmp_i**2 / |D|` (runners/diffusion.py:326-344)."""
    from torch.func import functional_call
    from sfron_b200.methods.diffusion import DiffusionUnlearner, adaptive_loss, per_sample_grad_rows
    torch.manual_seed(0)
    model = TinyNet().to(dev)
    x = torch.randn(6, 3, 8, 8, device=dev)
    y = torch.randint(0, 10, (6,), device=dev)
    # reference form: per-sample losses, one backward per sample over a retained graph
    loss = torch.nn.functional.cross_entropy(model(x), y, reduction="none")
    ref_rows = []
    for i in range(6):
        model.zero_grad()
        loss[i].backward(retain_graph=i != 5)
        ref_rows.append(torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone())
    model.zero_grad()
    ref_rows = torch.stack(ref_rows)
    want = torch.zeros(ref_rows.shape[1], device=dev)
    for r in ref_rows:
        want += r ** 2 / 50

    def one_sample_loss(pb, xi, yi):
        out = functional_call(model, pb, (xi.unsqueeze(0),))
        return torch.nn.functional.cross_entropy(out, yi.unsqueeze(0))

    un = DiffusionUnlearner(model, "ddpm")
    rows = per_sample_grad_rows(model, one_sample_loss, (x, y))
    assert rows.shape == ref_rows.shape and torch.allclose(rows, ref_rows, rtol=1e-4, atol=1e-7)
    fim = un.save_fim([rows[:4], rows[4:]], dataset_len=50)
    assert torch.allclose(fim, want, rtol=1e-4, atol=1e-12)
    # adaptive re-weighting keeps the batch mean scale and up-weights low-loss samples
    l = torch.tensor([0.5, 1.0, 2.0], device=dev)
    w = adaptive_loss(l, 3, gamma=1.0, eps=1e-8, keepdim=True) / l
    assert torch.allclose(w.sum(), torch.tensor(3.0, device=dev)) and w[0] > w[1] > w[2]


def test_host_gradient_feeder_pipeline(dev):
    """Host-buffer entry of the path: pinned host gradients -> device, double-buffered; every step must see
    exactly the gradients submitted for it, in order, and refuse pageable memory or an over-full pipe."""
    import sfron_b200 as sfr
    n = 100_003
    feeder = sfr.HostGradientFeeder(n, dev, slots=("forget", "remain"))
    hp = sfr.HotPath(n, dev, sfr.OptConfig())
    host = [(torch.full((n,), float(i)).pin_memory(), torch.full((n,), float(-i)).pin_memory()) for i in range(5)]
    feeder.submit(forget=host[0][0], remain=host[0][1])
    want = torch.zeros(n)
    for i in range(5):
        g = feeder.acquire()
        if i + 1 < 5:
            feeder.submit(forget=host[i + 1][0], remain=host[i + 1][1])      # overlaps this step's kernels
        assert float(g["forget"][0]) == float(i) and float(g["remain"][-1]) == float(-i)
        hp.fisher_accumulate("forget", g["forget"], 1.0)
        feeder.release()
        want += float(i) ** 2
    assert torch.equal(hp.forget_fisher.cpu(), want)
    assert feeder.bytes_per_step == 2 * 4 * n
    with pytest.raises(ValueError):
        feeder.submit(forget=torch.zeros(n), remain=host[0][1])              # pageable host memory
    feeder.submit(forget=host[0][0], remain=host[0][1])
    feeder.submit(forget=host[1][0], remain=host[1][1])
    with pytest.raises(RuntimeError):
        feeder.submit(forget=host[2][0], remain=host[2][1])                  # depth 2: pipe is full
