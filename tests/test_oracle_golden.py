"""Pins the CPU oracle against outputs of the REFERENCE ITSELF (tests/golden/*.pt, produced by
tests/golden/make_golden.py from /root/reference).  CPU only."""
import pytest
import torch

import os

from conftest import ROOT, bits_equal, load_golden, max_rel_err, unflat
from oracle import sfron_oracle as O

RTOL = 1e-6  # north-star tolerance for fp32 values; masks must be bit-exact


def _close(a, b):
    return max_rel_err(a, b, floor=1e-12) <= RTOL


@pytest.mark.parametrize("tag", ["default", "beta09"])
def test_classification_sfron_replay(tag):
    fx = load_golden(f"cls_sfron_{tag}.pt")
    names, shapes, hp = fx["names"], fx["shapes"], fx["hyper"]
    # Fisher: F += grad**2 / len(loader)  (sfron.py:288-291,315-318)
    for which in ("forget", "remain"):
        grads = fx[f"fisher_{which}_grads"]
        acc = O.fisher_init(names)
        for g in grads:
            O.fisher_accumulate(acc, unflat(g, names, shapes), len(grads))
        got = torch.cat([acc[n].reshape(-1) for n in names])
        assert bits_equal(got, fx[f"{which}_fisher"]), f"{which} Fisher differs from the reference run"
    # ratio mask (sfron.py:322-336)
    ff, rf = unflat(fx["forget_fisher"], names, shapes), unflat(fx["remain_fisher"], names, shapes)
    masks, zeros, total = O.ratio_mask(ff, rf, fx["threshold"])
    got = torch.cat([masks[n].reshape(-1) for n in names]).to(torch.uint8)
    assert torch.equal(got, fx["mask"])
    assert total == got.numel() and zeros == int((got == 0).sum())
    # forget / remain loop with slow-fast interpolation (sfron.py:189-259)
    loop = O.FlatReferenceLoop(shapes, unflat(fx["theta0"], names, shapes), "sgd",
                               dict(lr=hp["retain_lr"], momentum=hp["momentum"], weight_decay=hp["weight_decay"]),
                               ema_mode="slowfast", ema_a=hp["ema_beta"])
    for kind, lr, g in zip(fx["loop_kinds"], fx["loop_lrs"], fx["loop_grads"]):
        loop.set_lr(lr)
        gd = unflat(g, names, shapes)
        if kind == "forget":
            loop.forget_step(gd, mask=masks, max_norm=hp["max_norm"])
        else:
            loop.remain_step(gd, ema=True)
    assert _close(loop.flat("p"), fx["theta_final"]), max_rel_err(loop.flat("p"), fx["theta_final"])


def test_salun_topk_mask():
    fx = load_golden("salun_topk.pt")
    names, shapes = fx["names"], fx["shapes"]
    for th in ("0.2", "0.5"):
        rec = fx[th]
        # `gradients[name] += param.grad.data` per batch, then abs_ (salun.py:158-165)
        tot = torch.zeros_like(rec["grads"][0])
        for g in rec["grads"]:
            tot += g
        grads = unflat(tot.abs(), names, shapes)
        assert rec["mask_dtype"] == "torch.int64"
        for stable in (True, False):
            hard = O.topk_mask(grads, float(th), stable=stable)
            got = torch.cat([hard[n].reshape(-1) for n in names])
            assert got.dtype == torch.int64
            assert torch.equal(got, rec["mask"]), f"th={th} stable={stable}"
        flat = O.topk_mask_flat(tot, int(tot.numel() * float(th)))
        assert torch.equal(flat.to(torch.int64), rec["mask"])


@pytest.mark.parametrize("fixture", ["ddpm_ratio_mask.pt", "sd_ratio_mask.pt", "dit_ratio_mask.pt"])
def test_ratio_mask_scripts(fixture):
    fx = load_golden(fixture)
    for th_s, ref_masks in fx["masks"].items():
        th = float(th_s)
        masks, zeros, total = O.ratio_mask(fx["forget"], fx["remain"], th)
        assert list(masks.keys()) == list(ref_masks.keys())
        for name, ref in ref_masks.items():
            if torch.is_tensor(ref):
                assert masks[name].dtype == torch.bool and torch.equal(masks[name], ref), (th, name)
            else:
                assert masks[name] == 0 and ref == 0      # int placeholder for grad-less params
        assert total == sum(m.numel() for m in ref_masks.values() if torch.is_tensor(m))


def test_ddpm_adam_ema_loop():
    fx = load_golden("ddpm_adam_ema_loop.pt")
    names, shapes, hp = fx["names"], fx["shapes"], fx["hyper"]
    loop = O.FlatReferenceLoop(shapes, unflat(fx["theta0"], names, shapes), "adam",
                               dict(lr=hp["lr"], beta1=hp["beta1"], beta2=hp["beta2"], eps=hp["eps"],
                                    weight_decay=hp["weight_decay"]), ema_mode="ddpm", ema_a=hp["ema_rate"])
    mask = {n: m.bool() for n, m in unflat(fx["mask"], names, shapes).items()}
    for gf, gr in zip(fx["forget_grads"], fx["remain_grads"]):
        loop.forget_step(unflat(gf, names, shapes), mask=mask, max_norm=hp["grad_clip"])
        loop.remain_step(unflat(gr, names, shapes), max_norm=hp["grad_clip"], ema=True)
    assert _close(loop.flat("p"), fx["theta_final"])
    assert _close(loop.flat("m"), fx["exp_avg"])
    assert _close(loop.flat("v"), fx["exp_avg_sq"])
    assert _close(loop.flat("slow"), fx["ema_final"])


def test_dit_adamw_ema_loop():
    fx = load_golden("dit_adamw_ema_loop.pt")
    names, tnames, shapes, hp = fx["names"], fx["train_names"], fx["shapes"], fx["hyper"]
    theta0 = unflat(fx["theta0"], names, shapes)
    loop = O.FlatReferenceLoop({n: shapes[n] for n in tnames}, theta0, "adamw",
                               dict(lr=hp["lr"], weight_decay=hp["weight_decay"]), ema_mode="dit", ema_a=hp["decay"])
    tshapes = {n: shapes[n] for n in tnames}
    mask = {n: m.bool() for n, m in unflat(fx["mask"], tnames, tshapes).items()}
    frozen = {n: theta0[n] for n in names if n not in tnames}
    frozen_ema = {n: t.clone() for n, t in frozen.items()}
    for gf, gr in zip(fx["forget_grads"], fx["remain_grads"]):
        loop.forget_step(unflat(gf, tnames, tshapes), mask=mask, max_norm=hp["grad_clip"])
        loop.remain_step(unflat(gr, tnames, tshapes), max_norm=None, ema=True)
        O.ema_dit_(frozen_ema, frozen, hp["decay"])      # update_ema walks ALL parameters
    got_p = torch.cat([(loop.params[n].detach() if n in tnames else frozen[n]).reshape(-1) for n in names])
    got_e = torch.cat([(loop.slow[n] if n in tnames else frozen_ema[n]).reshape(-1) for n in names])
    assert _close(got_p, fx["theta_final"])
    assert _close(got_e, fx["ema_final"])
    for t, ref in enumerate(fx["cosine"]):
        assert O.cosine_lr_scheduler(25.0, t, 10) == ref


def _runner_fixture():
    fx = load_golden("ddpm_runner.pt")
    return fx, fx["names"], fx["shapes"], fx["hyper"]


def test_ddpm_runner_fisher_and_masks():
    """Diffusion.generate_fisher / generate_fisher_mask.py / Diffusion.generate_mask executed whole
    (DDPM/runners/diffusion.py:1210-1364, 930-1036) — Fisher of the CLIPPED batch gradient."""
    fx, names, shapes, hp = _runner_fixture()
    fi = fx["fisher"]
    fishers = {}
    for which, n_batches in (("forget", fx["n_forget_batches"]), ("remain", fx["n_remain_batches"])):
        acc = O.fisher_init(names)
        for g in fi[f"{which}_grads"]:
            O.fisher_accumulate_clipped(acc, unflat(g, names, shapes), n_batches, hp["grad_clip"])
        fishers[which] = acc
        got = torch.cat([acc[n].reshape(-1) for n in names])
        assert bits_equal(got, fi[f"{which}_fisher"]), which
    masks, _, _ = O.ratio_mask(unflat(fi["forget_fisher"], names, shapes), unflat(fi["remain_fisher"], names, shapes), 1.0)
    assert torch.equal(torch.cat([masks[n].reshape(-1) for n in names]).to(torch.uint8), fx["ratio_mask"])
    # SalUn top-k: sum of the clipped gradients, abs, global rank (:985-1034)
    tk = fx["topk"]
    tot = {n: 0 for n in names}
    for g in tk["grads"]:
        gd = unflat(g, names, shapes)
        O.clip_grad_norm(list(gd.values()), hp["grad_clip"])
        for n in names:
            tot[n] = tot[n] + gd[n]
    hard = O.topk_mask({n: t.abs() for n, t in tot.items()}, tk["ratio"])
    got = torch.cat([hard[n].reshape(-1) for n in names])
    assert got.dtype == torch.int64 and torch.equal(got, tk["mask"])


def _runner_loop(fx, names, shapes, hp):
    return O.FlatReferenceLoop(shapes, unflat(fx["theta0"], names, shapes), "adam",
                               dict(lr=hp["lr"], beta1=hp["beta1"], beta2=hp["beta2"], eps=hp["eps"],
                                    weight_decay=hp["weight_decay"]), ema_mode="ddpm", ema_a=hp["ema_rate"])


def _check_runner_final(loop, rec):
    assert _close(loop.flat("p"), rec["theta"])
    assert _close(loop.flat("m"), rec["exp_avg"])
    assert _close(loop.flat("v"), rec["exp_avg_sq"])
    assert _close(loop.flat("slow"), rec["ema"])


def test_ddpm_runner_sfron_forget():
    """Diffusion.sfron_forget executed whole (method ron, adaga, decayed alpha, ratio mask; :1038-1208)."""
    fx, names, shapes, hp = _runner_fixture()
    rec = fx["sfron"]
    loop = _runner_loop(fx, names, shapes, hp)
    mask = {n: m.bool() for n, m in unflat(fx["ratio_mask"], names, shapes).items()}
    for kind, g in zip(rec["kinds"], rec["grads"]):
        gd = unflat(g, names, shapes)
        if kind == "forget":
            loop.forget_step(gd, mask=mask, max_norm=hp["grad_clip"])
        else:
            loop.remain_step(gd, max_norm=hp["grad_clip"], ema=True)
    _check_runner_final(loop, rec)
    assert rec["step"] == hp["n_iters"] - 1 and set(rec["opt_steps"]) == {2.0 * hp["n_iters"]}
    assert rec["ckpt_model_keys"] == names and rec["ckpt_ema_keys"] == [n[len("module."):] for n in names]


def test_ddpm_runner_saliency_unlearn():
    """Diffusion.saliency_unlearn executed whole: one joint step per iteration, clip BEFORE the int64
    top-k mask, then EMA (:479-616)."""
    fx, names, shapes, hp = _runner_fixture()
    rec = fx["salun"]
    loop = _runner_loop(fx, names, shapes, hp)
    mask = unflat(fx["topk"]["mask"], names, shapes)
    for g in rec["grads"]:
        loop.forget_step(unflat(g, names, shapes), mask=mask, max_norm=hp["grad_clip"], order="clip_then_mask")
        loop.slow_update()
    _check_runner_final(loop, rec)


def test_dit_scripts_executed_whole():
    """DiT/generate_fisher.py, generate_mask.py and forget.py run as scripts (fixture dit_scripts.pt): Fisher bit-exact,
    mask bit-exact (frozen pos_embed keeps the int-0 placeholder), AdamW + clip + EMA trajectory within 1e-6."""
    fx = load_golden("dit_scripts.pt")
    names, tnames, shapes = fx["names"], fx["train_names"], fx["shapes"]
    tshapes = {n: shapes[n] for n in tnames}
    fi = fx["fisher"]
    acc = {}
    for which in ("forget", "remain"):
        acc[which] = O.fisher_init(names)
        for g in fi[f"{which}_grads"]:
            O.fisher_accumulate(acc[which], {**{n: None for n in names}, **unflat(g, tnames, tshapes)}, fx["n_fisher"])
        assert acc[which]["module.pos_embed"] == 0                    # never received a gradient
        got = torch.cat([acc[which][n].reshape(-1) for n in tnames])
        assert bits_equal(got, fi[f"{which}_fisher"]), which
    masks, _, _ = O.ratio_mask(acc["forget"], acc["remain"], 1.0)
    assert masks["module.pos_embed"] == 0
    assert torch.equal(torch.cat([masks[n].reshape(-1) for n in tnames]).to(torch.uint8), fx["ratio_mask"])
    rec, hp = fx["forget"], fx["forget"]["hyper"]
    theta0 = unflat(fx["theta0"], names, shapes)
    loop = O.FlatReferenceLoop(tshapes, theta0, "adamw", dict(lr=hp["lr"], weight_decay=0.0), ema_mode="dit", ema_a=hp["decay"])
    frozen = {n: theta0[n] for n in names if n not in tnames}
    frozen_ema = {n: t.clone() for n, t in frozen.items()}
    mask = {n: masks[n] for n in tnames}
    for kind, g in zip(rec["kinds"], rec["grads"]):
        gd = unflat(g, tnames, tshapes)
        if kind == "forget":
            loop.forget_step(gd, mask=mask, max_norm=hp["grad_clip"])
        else:
            loop.remain_step(gd, max_norm=None, ema=True)
            O.ema_dit_(frozen_ema, frozen, hp["decay"])
    got_p = torch.cat([(loop.params[n].detach() if n in tnames else frozen[n]).reshape(-1) for n in names])
    got_e = torch.cat([(loop.slow[n] if n in tnames else frozen_ema[n]).reshape(-1) for n in names])
    assert _close(got_p, rec["theta"]) and _close(got_e, rec["ema"])
    assert _close(loop.flat("m"), rec["exp_avg"]) and _close(loop.flat("v"), rec["exp_avg_sq"])
    assert rec["opt_state_keys"] == list(range(1, len(names))) and set(rec["opt_steps"]) == {2.0 * hp["n_iters"]}


def test_sd_scripts_executed_whole():
    """SD/train-scripts generate_fisher.py, generate_fisher_mask.py, nsfw_removal.py and gradient_ascent.py run whole
    (fixture sd_scripts.pt).  nsfw_removal.py never applies its mask (`n in parameters`, :157-160); the sibling
    gradient_ascent.py does (:94-99) — both behaviours are pinned."""
    fx = load_golden("sd_scripts.pt")
    names, shapes, fi = fx["names"], fx["shapes"], fx["fisher"]
    assert fx["full_names"] == ["model.diffusion_model." + n for n in names]       # what `n.split(...)[-1]` strips
    acc = {}
    for which in ("forget", "remain"):
        grads = fi[f"{which}_grads"]
        acc[which] = O.fisher_init(names)
        for g in grads:
            O.fisher_accumulate(acc[which], unflat(g, names, shapes), len(grads))
        assert bits_equal(torch.cat([acc[which][n].reshape(-1) for n in names]), fi[f"{which}_fisher"]), which
    masks, _, _ = O.ratio_mask(acc["forget"], acc["remain"], 1.0)
    assert torch.equal(torch.cat([masks[n].reshape(-1) for n in names]).to(torch.uint8), fx["ratio_mask"])
    theta0 = unflat(fx["theta0"], names, shapes)
    # nsfw_removal: forget step then remain step, one Adam, no mask in effect, no clip, no EMA
    rec = fx["nsfw_removal"]
    loop = O.FlatReferenceLoop(shapes, theta0, "adam", dict(lr=rec["lr"]))
    for i, g in enumerate(rec["grads"]):
        gd = unflat(g, names, shapes)
        loop.forget_step(gd, mask=None) if i % 2 == 0 else loop.remain_step(gd, ema=False)
    assert _close(loop.flat("p"), rec["theta"])
    masked = O.FlatReferenceLoop(shapes, theta0, "adam", dict(lr=rec["lr"]))
    for i, g in enumerate(rec["grads"]):
        gd = unflat(g, names, shapes)
        masked.forget_step(gd, mask=masks) if i % 2 == 0 else masked.remain_step(gd, ema=False)
    assert not _close(masked.flat("p"), rec["theta"])                              # the mask really is not applied there
    # gradient_ascent: one masked step per iteration on the joint loss
    rec = fx["gradient_ascent"]
    loop = O.FlatReferenceLoop(shapes, theta0, "adam", dict(lr=rec["lr"]))
    for g in rec["grads"]:
        loop.forget_step(unflat(g, names, shapes), mask=masks)
    assert _close(loop.flat("p"), rec["theta"])


def test_ddpm_sa_forget_ewc_penalty_executed_whole():
    """Diffusion.sa_forget executed whole (fixture ddpm_sa_forget.pt): the per-step gradient the reference's backward
    produced equals (1 + gamma) * c (the linear stand-in loss) + the EWC gradient of `lmbda * sum(F * (p - p_mle)**2)`
    as the oracle forms it (runners/diffusion.py:424-433), and clip -> Adam -> EMA reproduce the saved checkpoint."""
    fx = load_golden("ddpm_sa_forget.pt")
    names, shapes, hp = fx["names"], fx["shapes"], fx["hyper"]
    theta0 = unflat(fx["theta0"], names, shapes)
    fisher = unflat(fx["fisher"], names, shapes)
    loop = O.FlatReferenceLoop(shapes, theta0, "adam", dict(lr=hp["lr"], beta1=hp["beta1"], beta2=hp["beta2"], eps=hp["eps"],
                                                            weight_decay=hp["weight_decay"]), ema_mode="ddpm", ema_a=hp["ema_rate"])
    for step, g_ref in enumerate(fx["grads"]):
        ewc, penalty = O.ewc_penalty_grads({n: p.detach() for n, p in loop.params.items()}, theta0, fisher, hp["lmbda"])
        total = fx["base_grad"] + torch.cat([ewc[n].reshape(-1) for n in names])
        if step == 0:
            assert float(penalty) == 0.0                       # theta == theta_mle: only the base term moves it
        rms = g_ref.double().pow(2).mean().sqrt()
        assert bool(((total.double() - g_ref.double()).abs() <= 1e-6 * (g_ref.double().abs() + rms)).all()), step
        loop.forget_step(unflat(total, names, shapes), mask=None, max_norm=hp["grad_clip"])
        loop.slow_update()
    assert _close(loop.flat("p"), fx["theta"]) and _close(loop.flat("slow"), fx["ema"])


def test_ddpm_save_fim_executed_whole():
    """Diffusion.save_fim executed whole (fixture ddpm_fim.pt): per-sample gradients accumulated over the timestep
    chunks, then `F[name] += tmp_i[name]**2 / |D|` sample by sample, batch by batch (runners/diffusion.py:326-344)."""
    fx = load_golden("ddpm_fim.pt")
    names, shapes = fx["names"], fx["shapes"]
    acc = {n: torch.zeros(shapes[n]) for n in names}
    for batch in fx["chunk_grads"]:                            # [n_chunks, bs, n]
        rows = torch.zeros_like(batch[0])
        for chunk in batch:                                    # fisher_dict_temp_list[i][name] += param.grad.data
            rows = rows + chunk
        O.per_sample_fim(acc, [unflat(r, names, shapes) for r in rows], fx["dataset_len"])
    got = torch.cat([acc[n].reshape(-1) for n in names])
    assert bits_equal(got, fx["fim"])


# ------------------------------------------------------------------ the reference's REAL networks (configs 1 and 2)
def _clip_coef(total_norm, max_norm):
    """clip_grad_norm_'s coefficient from the fp32 total norm the reference computed (torch/nn/utils/clip_grad.py)."""
    return torch.clamp(torch.tensor(max_norm, dtype=torch.float32) / (total_norm + 1e-6), max=1.0)


@pytest.mark.parametrize("which", ["resnet18", "ddpm"])
def test_real_reference_networks_fisher_and_mask_slices(which):
    """real_models.pt: SFRon.get_weight_saliency_mask on the reference's ResNet18 and Diffusion.generate_fisher +
    generate_fisher_mask.py on its Conditional_Model, executed whole.  Key lists / order / prefixes of the files are
    pinned for ALL tensors; for the tensors of <= 8192 elements the oracle reproduces Fisher and mask bit for bit."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    fx = load_golden("real_models.pt")[which]
    names, small = fx["names"], fx["small_names"]
    with torch.device("meta"):
        if which == "resnet18":
            from resnet18_cifar import ResNet18Harness
            harness, prefix = ResNet18Harness(), ""
        else:
            from ddpm_unet import DDPMCondUNet
            harness, prefix = DDPMCondUNet(), "module."
    assert [(prefix + n, list(p.shape)) for n, p in harness.named_parameters()] == [(n, fx["shapes"][n]) for n in names]
    assert fx["forget_fisher_keys"] == names == fx["remain_fisher_keys"] == fx["mask_keys"]
    assert fx["forget_fisher_dtype"] == "torch.float32" and fx["mask_dtype"] == "torch.bool"
    nf, nr = fx["n_forget"], fx["n_remain"]
    for role, recs, count in (("forget_fisher", fx["grads"][:nf], nf), ("remain_fisher", fx["grads"][nf:], nr)):
        acc = O.fisher_init(small)
        for i, g in enumerate(recs):
            if which == "ddpm":                                   # Fisher of the CLIPPED gradient (:1270-1281)
                coef = _clip_coef(fx["torch_total_norms"][i if role == "forget_fisher" else nf + i], fx["grad_clip"])
                g = {n: t * coef for n, t in g.items()}
            O.fisher_accumulate(acc, g, count)
        for n in small:
            assert bits_equal(acc[n], fx[role][n]), (role, n)
    mask, _, _ = O.ratio_mask(fx["forget_fisher"], fx["remain_fisher"], fx["threshold"])
    assert all(torch.equal(mask[n], fx["mask"][n]) for n in small)
