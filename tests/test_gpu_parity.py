"""GPU parity tests: the sm_100a kernels, called through the C ABI (sfron_b200.capi / HotPath),
against the CPU oracle on the same seeded inputs and against the reference-generated golden
fixtures.  Bar (BASELINE.json north_star): masks bit-exact; Fisher values and updated weights
within 1e-6 relative (fp32).  K1 / K2a / K2b are in fact bit-exact; K3 cannot be bit-exact
against CPU torch because torch's AVX-512 `sqrt` is not correctly rounded.
"""

import pytest
import torch

from conftest import bits_equal, load_golden, unflat
from oracle import sfron_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-6
RAGGED_SIZES = [1, 2, 3, 4, 5, 7, 31, 1023, 1024, 4097, 65536 + 3, 1_000_003]


@pytest.fixture(scope="module")
def sfr():
    import sfron_b200
    assert torch.cuda.is_available(), "these tests need a GPU"
    sfron_b200.capi.load()
    return sfron_b200


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def close(a, b, rtol=RTOL):
    """|a-b| <= rtol * (|b| + rms(b)) elementwise: 1e-6 relative to the element, with 1e-6 of the
    vector's scale as the absolute floor.  A floor is unavoidable in fp32: a weight that cancels to
    ~0 along the trajectory still carries the <= 1 ulp (6e-8 x its former magnitude) rounding
    difference that any perturbation of the clip norm in the last bit produces — and torch's own
    fp32 norm differs from the exact one by up to ~6e-7 relative at n = 2e5 (measured).  For
    scale: torch.testing.assert_close's fp32 default is rtol 1.3e-6 + atol 1e-5, >100x looser."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    rms = b.pow(2).mean().sqrt().item() if b.numel() else 0.0
    bad = (a - b).abs() > rtol * (b.abs() + rms)
    return not bool(bad.any())


def gen(seed):
    return torch.Generator().manual_seed(seed)


# =============================================================================== K1 Fisher
@pytest.mark.parametrize("tag", ["default", "beta09"])
def test_k1_golden_classification_fisher(sfr, dev, tag):
    fx = load_golden(f"cls_sfron_{tag}.pt")
    for which in ("forget", "remain"):
        grads = fx[f"fisher_{which}_grads"].to(dev)
        acc = torch.zeros(grads.shape[1], device=dev)
        for g in grads:
            sfr.capi.fisher_accum(acc, g.clone(), float(len(grads)))   # rows of a [B, N] tensor are not 16-B aligned
        assert bits_equal(acc.cpu(), fx[f"{which}_fisher"])


@pytest.mark.parametrize("n", RAGGED_SIZES)
def test_k1_bit_exact_vs_oracle(sfr, dev, n):
    g = gen(n)
    acc0 = torch.rand(n, generator=g) * 1e-4
    grads = [torch.randn(n, generator=g) * 10 ** float(torch.randint(-6, 2, (1,), generator=g)) for _ in range(3)]
    ref = acc0.clone()
    for x in grads:
        O.flat_fisher_accum(ref, x, 7)
    acc = acc0.to(dev)
    for x in grads:
        sfr.capi.fisher_accum(acc, x.to(dev), 7.0)
    assert bits_equal(acc.cpu(), ref)


def test_k1_first_accumulation_matches_python_zero(sfr, dev):
    # the reference accumulators start as Python int 0: 0 + g**2/L
    x = torch.randn(10_001, generator=gen(1))
    acc = torch.zeros(x.numel(), device=dev)
    sfr.capi.fisher_accum(acc, x.to(dev), 13.0)
    assert bits_equal(acc.cpu(), 0 + x ** 2 / 13)


@pytest.mark.parametrize("rows,n", [(2, 1000), (5, 4096 + 8), (8, 100_000)])
def test_k1_per_sample_fim_rows(sfr, dev, rows, n):
    g = gen(rows * n)
    tmp = torch.randn(rows, n, generator=g)
    acc = {"w": torch.zeros(n)}
    O.per_sample_fim(acc, [{"w": tmp[i]} for i in range(rows)], 50)
    out = torch.zeros(n, device=dev)
    sfr.capi.fisher_accum(out, tmp.to(dev), 50.0)
    assert bits_equal(out.cpu(), acc["w"])


def test_k1_clipped_gradient_ddpm(sfr, dev):
    n = 50_003
    hp = sfr.HotPath(n, dev, sfr.OptConfig())
    for scale in (5.0, 1e-3):   # norm above and below max_norm
        x = torch.randn(n, generator=gen(3)) * scale
        acc = {"w": 0}
        O.fisher_accumulate_clipped(acc, {"w": x}, 4, 1.0)
        hp.buffer("forget_fisher").zero_()
        hp.fisher_accumulate("forget", x.to(dev), 4.0, clip_max_norm=1.0)
        assert close(hp.forget_fisher, acc["w"])


@pytest.mark.parametrize("n", [5, 4099, 1_000_003])
def test_saliency_accumulate_matches_clip_then_add(sfr, dev, n):
    """SalUn mask input: `clip_grad_norm_ ; gradients[name] += grad` (runners/diffusion.py:985-994) in one fused
    pass.  Unclipped (salun.py:163-169) it is bit-exact; clipped it follows the oracle's clip to 1e-6 (the clip
    coefficient comes from a double-precision norm here, from torch's fp32 norm there)."""
    hp = sfr.HotPath(n, dev, sfr.OptConfig())
    acc_ref = torch.zeros(n)
    g = gen(21)
    for _ in range(3):
        x = torch.randn(n, generator=g) * 0.3
        acc_ref += x
        hp.saliency_accumulate(x.to(dev))
    assert bits_equal(hp.buffer("grad_sum").cpu(), acc_ref)
    hp.buffer("grad_sum").zero_()
    acc_ref = torch.zeros(n)
    norm_err = 0.0
    for scale in (5.0, 1e-4, 2.0):                      # norms above and below max_norm
        x = torch.randn(n, generator=g) * scale
        y = x.clone()
        tn = float(O.clip_grad_norm([y], 1.0))
        exact = x.double().pow(2).sum().sqrt().item()
        norm_err = max(norm_err, abs(tn - exact) / exact)
        acc_ref += y
        hp.saliency_accumulate(x.to(dev), clip_max_norm=1.0)
    # torch's CPU fp32 norm of ONE million-element tensor is itself ~1e-5 off the exact norm (measured here, and it
    # depends on the host's thread count); the kernel's norm accumulates in double.  The bar is 1e-6 beyond that.
    assert close(hp.buffer("grad_sum"), acc_ref, RTOL + 1.5 * norm_err), norm_err


def test_k1_bf16_gradients(sfr, dev):
    n = 33_333
    x = torch.randn(n, generator=gen(4)).bfloat16()
    ref = torch.zeros(n)
    O.flat_fisher_accum(ref, x.float(), 3)
    acc = torch.zeros(n, device=dev)
    sfr.capi.fisher_accum(acc, x.to(dev), 3.0)
    assert bits_equal(acc.cpu(), ref)


# =============================================================================== K2a ratio mask
@pytest.mark.parametrize("fixture,prefix", [("ddpm_ratio_mask.pt", ""), ("sd_ratio_mask.pt", ""),
                                            ("dit_ratio_mask.pt", "")])
def test_k2a_golden_ratio_masks(sfr, dev, fixture, prefix):
    fx = load_golden(fixture)
    named = [(n, tuple(t.shape)) for n, t in fx["forget"].items() if torch.is_tensor(t)]
    layout = sfr.FlatLayout(named)
    all_names = list(fx["forget"].keys())
    ff = sfr.formats.dict_to_flat(layout, fx["forget"], device=dev)
    rf = sfr.formats.dict_to_flat(layout, fx["remain"], device=dev)
    hp = sfr.HotPath(layout.numel, dev, sfr.OptConfig())
    hp.set_buffer("forget_fisher", ff)
    hp.set_buffer("remain_fisher", rf)
    ths = [float(t) for t in fx["masks"]]
    multi = hp.ratio_masks(ths)
    multi_zeros = hp.zero_count.cpu().tolist()
    for i, (th_s, ref) in enumerate(fx["masks"].items()):
        mask = hp.ratio_mask(float(th_s))
        got = sfr.formats.ratio_mask_to_dict(layout, mask, all_names=all_names)
        assert list(got.keys()) == list(ref.keys())
        zeros = 0
        for name, r in ref.items():
            if torch.is_tensor(r):
                assert got[name].dtype == torch.bool and torch.equal(got[name], r), (th_s, name)
                zeros += int(r.numel() - r.count_nonzero())
            else:
                assert got[name] == 0
        assert int(hp.zero_count[0]) == zeros
        assert torch.equal(multi[i, :layout.numel], mask) and multi_zeros[i] == zeros


@pytest.mark.parametrize("n", RAGGED_SIZES)
def test_k2a_bit_exact_vs_oracle_edge_values(sfr, dev, n):
    g = gen(n + 1)
    ff = torch.randn(n, generator=g).pow(2) * 1e-7
    rf = torch.randn(n, generator=g).pow(2) * 1e-7
    specials = torch.tensor([0.0, 1e-15, 1e-38, 1e-45, float("inf"), float("nan"), 1.0, 3.0e38])
    idx = torch.randint(0, n, (min(n, 64),), generator=g)
    ff[idx] = specials[torch.randint(0, len(specials), idx.shape, generator=g)]
    idx = torch.randint(0, n, (min(n, 64),), generator=g)
    rf[idx] = specials[torch.randint(0, len(specials), idx.shape, generator=g)]
    for th in (1.0, 0.5, 10, 0.1):
        ref = O.flat_ratio_mask(ff, rf, th)
        mask = torch.empty(n, dtype=torch.uint8, device=dev)
        zc = torch.zeros(1, dtype=torch.int64, device=dev)
        sfr.capi.ratio_mask(ff.to(dev), rf.to(dev), th, mask, zc)
        assert torch.equal(mask.cpu().bool(), ref)
        assert int(zc) == int(n - ref.count_nonzero())


def test_k2a_golden_classification_mask(sfr, dev):
    fx = load_golden("cls_sfron_default.pt")
    mask = torch.empty(fx["mask"].numel(), dtype=torch.uint8, device=dev)
    sfr.capi.ratio_mask(fx["forget_fisher"].to(dev), fx["remain_fisher"].to(dev), fx["threshold"], mask)
    assert torch.equal(mask.cpu(), fx["mask"])


# =============================================================================== K2b top-k select
def run_topk(sfr, dev, values, k, other=None):
    """Runs BOTH forms of the select on a garbage-filled mask buffer — the two-pass form (pass 1 leaves the
    provisional mask, apply resolves the staged candidates) and the three-read form (streaming apply) — and
    requires identical bytes; returns the two-pass result."""
    hp = sfr.HotPath(values.numel(), dev, sfr.OptConfig())
    v, o = values.to(dev), None if other is None else other.to(dev)
    out = {}
    for two_pass in (False, True):
        hp.select_two_pass = two_pass
        buf = torch.full((values.numel(),), 0xCC, dtype=torch.uint8, device=dev)
        out[two_pass] = hp.topk_mask(v, k, other=o, out=buf).cpu()
    assert torch.equal(out[True], out[False]), "two-pass and three-read selects disagree"
    return out[True], hp.select_state()


def test_k2b_two_pass_apply_on_another_mask_buffer_falls_back(sfr, dev):
    """sfr_select_apply given a mask buffer other than the one pass 1 wrote must not trust the provisional
    mask: it streams the vector and still produces the exact result; repeating apply is idempotent."""
    n, k = 300_007, 120_000
    x = torch.randn(n, generator=gen(11)) * 1e-2
    x[::7] = x[3]                                   # ties at a value near the threshold region
    xd = x.to(dev)
    hp = sfr.HotPath(n, dev, sfr.OptConfig())
    state, bins, scratch = hp._select_buffers()
    capi = sfr.capi
    first = torch.full((n,), 0xCC, dtype=torch.uint8, device=dev)
    other = torch.full((n,), 0xCC, dtype=torch.uint8, device=dev)
    capi.select_init(state, bins, k)
    capi.select_hist(xd, None, capi.KEY_ABS, 0, state, bins)
    capi.select_scan(0, state, bins)
    capi.select_hist(xd, None, capi.KEY_ABS, 1, state, bins, scratch, mask=first)
    capi.select_scan(1, state, bins)
    capi.select_apply(xd, None, capi.KEY_ABS, state, None, scratch, other)      # not the pass-1 buffer
    capi.select_apply(xd, None, capi.KEY_ABS, state, None, scratch, first)
    capi.select_apply(xd, None, capi.KEY_ABS, state, None, scratch, first)      # again: idempotent
    ref = O.topk_mask_flat(x, k)
    assert torch.equal(other.cpu(), ref) and torch.equal(first.cpu(), ref)


def test_k2b_golden_salun(sfr, dev):
    fx = load_golden("salun_topk.pt")
    for th in ("0.2", "0.5"):
        acc = torch.zeros_like(fx[th]["grads"][0])
        for g in fx[th]["grads"]:
            acc += g
        k = int(acc.numel() * float(th))
        mask, st = run_topk(sfr, dev, acc, k)
        assert torch.equal(mask.to(torch.int64), fx[th]["mask"])
        assert int(mask.sum()) == k


@pytest.mark.parametrize("n", [1, 5, 1000, 8192, 8193, 100_003, 1_000_003])
@pytest.mark.parametrize("ratio", [0.0, 0.2, 0.5, 0.999, 1.0])
def test_k2b_bit_exact_vs_stable_argsort(sfr, dev, n, ratio):
    x = torch.randn(n, generator=gen(n)) * 1e-3
    k = int(n * ratio)
    mask, st = run_topk(sfr, dev, x, k)
    assert torch.equal(mask, O.topk_mask_flat(x, k))
    assert int(mask.sum()) == k


@pytest.mark.parametrize("n", [4096, 70_001, 1_000_003])
def test_k2b_ties_lowest_index_first(sfr, dev, n):
    g = gen(n + 7)
    # heavy ties: 12 distinct magnitudes, many exact zeros, both signs, a few NaN / inf
    x = torch.randint(0, 12, (n,), generator=g).float() * 0.25
    x[torch.rand(n, generator=g) < 0.3] = 0.0
    x *= torch.where(torch.rand(n, generator=g) < 0.5, -1.0, 1.0)
    x[torch.randint(0, n, (5,), generator=g)] = float("nan")
    x[torch.randint(0, n, (3,), generator=g)] = float("inf")
    n_nan = int(x.isnan().sum())
    for k in (1, n // 10, n // 3, n // 2, n - n_nan - 1, n - n_nan):
        mask, st = run_topk(sfr, dev, x, k)
        # oracle: stable argsort of -|x| with NaN mapped to the end (torch sorts NaN last already)
        ref = O.topk_mask_flat(x, k)
        assert torch.equal(mask, ref), k
        assert int(mask.sum()) == k
        assert not bool(mask[x.isnan()].any())


@pytest.mark.parametrize("prefix", [0, 1, 2, 0x007f, 0x0080, 0x3c23, 0x7f7f, 0x7f80])
def test_k2b_values_on_prefix_boundaries(sfr, dev, prefix):
    """Pass 1 tests key[30:16] against the chosen prefix with float compares on the value (select.cu hist1_tiles); the
    bounds are the bit patterns (prefix << 16) - 1 and ((prefix + 1) << 16) - 1.  Values sitting exactly on, one below and
    one above both bounds (subnormal bounds, FLT_MAX / inf at the top), with signs, zeros and NaN mixed in, and k swept so
    that the threshold falls on each of them."""
    g = gen(prefix + 3)
    lo, hi = prefix << 16, (prefix + 1) << 16                      # keys of the bin: [lo, hi); key = bits + 1
    bits = [b for b in (lo - 3, lo - 2, lo - 1, lo, lo + 1, (lo + hi) // 2, hi - 3, hi - 2, hi - 1, hi, hi + 1)
            if 0 <= b <= 0x7f800000]
    pool = torch.tensor(bits, dtype=torch.int32).view(torch.float32)
    n = 40_003
    x = pool[torch.randint(0, len(bits), (n,), generator=g)].clone()
    x *= torch.where(torch.rand(n, generator=g) < 0.5, -1.0, 1.0)
    x[torch.randint(0, n, (7,), generator=g)] = float("nan")
    x[torch.randint(0, n, (9,), generator=g)] = 0.0
    n_num = int((~x.isnan()).sum())
    ks = {1, n_num}
    for b in bits:                                                 # thresholds just before / inside / after each value's run
        c = int((x.abs().view(torch.int32) >= b).logical_and(~x.isnan()).sum())
        ks.update(k for k in (c - 1, c, c + 1) if 1 <= k <= n_num)
    for k in sorted(ks):
        mask, _ = run_topk(sfr, dev, x, k)
        assert torch.equal(mask, O.topk_mask_flat(x, k)), (prefix, k)


@pytest.mark.parametrize("n", [3_000_000, 24_000_000])
def test_k2b_ties_spread_over_many_chunks(sfr, dev, n):
    """One threshold-equal key in EVERY 8192-element chunk, half of them selected (lowest index first).  At 3 M elements the
    ordered apply ranks them straight from the short tie-chunk list (select.cu kTieListFast); at 24 M the list is longer
    than that and the two-level scan of the per-chunk counters runs instead.  The select state says which."""
    g = gen(n)
    x = torch.randn(n, generator=g) * 1e-2
    v = float(x.abs().median())
    x[::8192] = v
    x[1::16384] = -v                                    # two per chunk in every other chunk, both signs
    ties = int((x.abs() == v).sum())
    greater = int((x.abs() > v).sum())
    k = greater + ties // 2
    mask, st = run_topk(sfr, dev, x, k)
    assert st.count_eq == ties and st.tie_budget == ties // 2
    assert torch.equal(mask, O.topk_mask_flat(x, k))
    assert int(mask.sum()) == k


def test_k2b_all_equal_and_all_zero(sfr, dev):
    for val in (0.0, 2.5):
        x = torch.full((50_000,), val)
        for k in (0, 1, 17, 25_000, 50_000):
            mask, _ = run_topk(sfr, dev, x, k)
            ref = torch.zeros(50_000, dtype=torch.uint8)
            ref[:k] = 1
            assert torch.equal(mask, ref)


def test_k2b_ratio_key_mode(sfr, dev):
    n = 200_003
    g = gen(5)
    ff = torch.randn(n, generator=g).pow(2) * 1e-6
    rf = torch.randn(n, generator=g).pow(2) * 1e-6
    ratio = (ff + 1e-15) / (rf + 1e-15)
    k = n // 4
    mask, _ = run_topk(sfr, dev, ff, k, other=rf)
    assert torch.equal(mask, O.topk_mask_flat(ratio, k))


def test_k2b_config2_size_properties(sfr, dev):
    """DDPM-sized vector (N2 = 38,632,323): size-independent properties of an exact top-k."""
    n = 38_632_323
    x = torch.randn(n, device=dev, generator=torch.Generator(device=dev).manual_seed(0)) * 1e-2
    k = int(n * 0.5)
    hp = sfr.HotPath(n, dev, sfr.OptConfig())
    mask = hp.topk_mask(x, k).bool()
    assert int(mask.sum()) == k
    a = x.abs()
    assert a[mask].min() >= a[~mask].max()
    st = hp.select_state()
    assert st.count_gt + st.tie_budget == k
    # idempotence: selecting again gives the same mask
    assert torch.equal(hp.topk_mask(x, k).bool(), mask)


# =============================================================================== clip norm
@pytest.mark.parametrize("n", RAGGED_SIZES)
def test_masked_sumsq(sfr, dev, n):
    g = gen(n + 11)
    x = torch.randn(n, generator=g)
    m = (torch.rand(n, generator=g) < 0.4).to(torch.uint8)
    out = torch.zeros(1, dtype=torch.float64, device=dev)
    sfr.capi.masked_sumsq(x.to(dev), m.to(dev), out)
    ref = (x.double() * m.double()).pow(2).sum().item()
    assert abs(out.item() - ref) <= 1e-7 * max(ref, 1e-30)
    out.zero_()
    sfr.capi.masked_sumsq(x.to(dev), None, out)
    ref = x.double().pow(2).sum().item()
    assert abs(out.item() - ref) <= 1e-7 * ref


# =============================================================================== K3 fused update
def flat_loop(n, theta0, opt, kw, ema_mode, ema_a):
    return O.FlatReferenceLoop({"w": (n,)}, {"w": theta0}, opt, kw, ema_mode=ema_mode, ema_a=ema_a)


def engine_for(sfr, dev, n, theta0, opt, kw, ema_mode, ema_a):
    cfg = sfr.OptConfig(kind=opt, lr=kw["lr"], beta1=kw.get("beta1", 0.9), beta2=kw.get("beta2", 0.999),
                        eps=kw.get("eps", 1e-8), weight_decay=kw.get("weight_decay", 0.0),
                        momentum=kw.get("momentum", 0.0), dampening=kw.get("dampening", 0.0))
    hp = sfr.HotPath(n, dev, cfg, ema_mode=ema_mode, ema_a=ema_a)
    p = theta0.to(dev).clone()
    if ema_mode != "none":
        hp.init_slow(p)
    return hp, p


def test_k3_golden_classification_loop(sfr, dev):
    for tag in ("default", "beta09"):
        fx = load_golden(f"cls_sfron_{tag}.pt")
        hp_ = fx["hyper"]
        n = fx["theta0"].numel()
        hp, p = engine_for(sfr, dev, n, fx["theta0"], "sgd",
                           dict(lr=hp_["retain_lr"], momentum=hp_["momentum"], weight_decay=hp_["weight_decay"]),
                           "slowfast", hp_["ema_beta"])
        hp.set_buffer("mask", fx["mask"].to(dev))
        for kind, lr, g in zip(fx["loop_kinds"], fx["loop_lrs"], fx["loop_grads"]):
            gd = g.to(dev).clone()
            if kind == "forget":
                hp.forget_step(p, gd, max_norm=hp_["max_norm"], lr=lr)
            else:
                hp.remain_step(p, gd, lr=lr, ema=True)
        assert close(p, fx["theta_final"]), tag


def test_k3_golden_ddpm_adam_ema_loop(sfr, dev):
    fx = load_golden("ddpm_adam_ema_loop.pt")
    h = fx["hyper"]
    n = fx["theta0"].numel()
    hp, p = engine_for(sfr, dev, n, fx["theta0"], "adam",
                       dict(lr=h["lr"], beta1=h["beta1"], beta2=h["beta2"], eps=h["eps"],
                            weight_decay=h["weight_decay"]), "ddpm", h["ema_rate"])
    hp.set_buffer("mask", fx["mask"].to(dev))
    for gf, gr in zip(fx["forget_grads"], fx["remain_grads"]):
        hp.forget_step(p, gf.to(dev).clone(), max_norm=h["grad_clip"])
        hp.remain_step(p, gr.to(dev).clone(), max_norm=h["grad_clip"], ema=True)
    assert close(p, fx["theta_final"])
    assert close(hp.m, fx["exp_avg"])
    assert close(hp.v, fx["exp_avg_sq"])
    assert close(hp.slow, fx["ema_final"])


def test_k3_golden_dit_adamw_ema_loop(sfr, dev):
    fx = load_golden("dit_adamw_ema_loop.pt")
    h, names, tnames, shapes = fx["hyper"], fx["names"], fx["train_names"], fx["shapes"]
    theta0 = unflat(fx["theta0"], names, shapes)
    t0 = torch.cat([theta0[n].reshape(-1) for n in tnames])
    frozen = torch.cat([theta0[n].reshape(-1) for n in names if n not in tnames]).to(dev)
    frozen_ema = frozen.clone()
    hp, p = engine_for(sfr, dev, t0.numel(), t0, "adamw", dict(lr=h["lr"], weight_decay=h["weight_decay"]),
                       "dit", h["decay"])
    hp.set_buffer("mask", fx["mask"].to(dev))
    for gf, gr in zip(fx["forget_grads"], fx["remain_grads"]):
        hp.forget_step(p, gf.to(dev).clone(), max_norm=h["grad_clip"])
        hp.remain_step(p, gr.to(dev).clone(), ema=True)
        hp.ema_only(frozen, frozen_ema)
    final = unflat(fx["theta_final"], names, shapes)
    ema_final = unflat(fx["ema_final"], names, shapes)
    assert close(p, torch.cat([final[n].reshape(-1) for n in tnames]))
    assert close(hp.slow, torch.cat([ema_final[n].reshape(-1) for n in tnames]))
    assert close(frozen_ema, torch.cat([ema_final[n].reshape(-1) for n in names if n not in tnames]))


CASES = [
    ("sgd", dict(lr=0.01, momentum=0.9, weight_decay=5e-4), "slowfast", 0.9),
    ("sgd", dict(lr=0.05, momentum=0.0, weight_decay=0.0), "none", 0.0),
    ("sgd", dict(lr=0.01, momentum=0.9, weight_decay=0.0, dampening=0.1), "none", 0.0),
    ("adam", dict(lr=1e-4, weight_decay=0.0), "ddpm", 1e-4),
    ("adam", dict(lr=1e-3, weight_decay=1e-2, beta1=0.5, beta2=0.99, eps=1e-6), "ddpm", 0.999),
    ("adamw", dict(lr=1e-4, weight_decay=0.0), "dit", 0.9999),
    ("adamw", dict(lr=2e-5, weight_decay=0.05), "slowfast", 0.5),
    ("adamw", dict(lr=1e-3, weight_decay=0.01, beta1=0.4), "none", 0.0),
]


@pytest.mark.parametrize("opt,kw,ema_mode,ema_a", CASES)
@pytest.mark.parametrize("n", [1, 6, 4099, 200_001])
@pytest.mark.parametrize("order", ["mask_then_clip", "clip_then_mask"])
@pytest.mark.parametrize("launch", ["coop", "split"])
def test_k3_vs_torch_optim_trajectory(sfr, dev, opt, kw, ema_mode, ema_a, n, order, launch):
    """launch: clipped steps as ONE cooperative launch (sfr_clipped_update, the default for vectors of this size) or
    as the separate norm / scalar-prep / update launches large vectors use — same arithmetic either way."""
    g = gen(n * 31 + len(opt))
    theta0 = torch.randn(n, generator=g) * 0.02
    mask = torch.rand(n, generator=g) < 0.35
    ref = flat_loop(n, theta0, opt, kw, ema_mode, ema_a)
    hp, p = engine_for(sfr, dev, n, theta0, opt, kw, ema_mode, ema_a)
    if launch == "split":
        hp.coop_max_elems = 0
    hp.set_buffer("mask", mask.to(torch.uint8).to(dev))
    for step in range(8):
        gf = torch.randn(n, generator=g) * (4.0 if step % 2 else 0.01)
        gr = torch.randn(n, generator=g) * 0.1
        clip_r = 1.0 if step % 3 == 0 else None
        ref.forget_step({"w": gf}, mask={"w": mask}, max_norm=1.0, order=order)
        ref.remain_step({"w": gr}, max_norm=clip_r, ema=True)
        hp.forget_step(p, gf.to(dev), max_norm=1.0, mask_order=order)
        hp.remain_step(p, gr.to(dev), max_norm=clip_r, ema=True)
    # One configuration is ill-conditioned on purpose (Adam, eps 1e-6, L2 decay 1e-2, lr 1e-3): where
    # the clipped gradient nearly cancels against wd*p, Adam's normalisation amplifies an absolute
    # gradient error by step_size/eps ~ 2e3, and the seed error is torch's OWN fp32 norm, which is off
    # from the exact norm by up to ~6e-7 relative at n = 2e5 (tests above pin ours to 1e-7 of exact).
    # That case is held to 1e-5; every other case to the 1e-6 bar.
    rtol = 1e-5 if (kw.get("eps") == 1e-6 and n > 100_000) else RTOL
    assert close(p, ref.flat("p"), rtol)
    if opt != "sgd":
        assert close(hp.m, ref.flat("m"), rtol) and close(hp.v, ref.flat("v"), rtol)
    elif kw.get("momentum", 0.0):
        assert close(hp.m, ref.flat("buf"), rtol)
    if ema_mode != "none":
        assert close(hp.slow, ref.flat("slow"), rtol)


def test_k3_zero_grad_and_bf16_copy(sfr, dev):
    n = 70_001
    g = gen(9)
    theta0 = torch.randn(n, generator=g) * 0.02
    hp, p = engine_for(sfr, dev, n, theta0, "adamw", dict(lr=1e-3), "none", 0.0)
    grad = (torch.randn(n, generator=g)).to(dev)
    p16 = torch.empty(n, dtype=torch.bfloat16, device=dev)
    hp.remain_step(p, grad, ema=False, zero_grad=True, p_bf16=p16)
    assert int(grad.count_nonzero()) == 0
    assert torch.equal(p16, p.bfloat16())


def test_k3_bf16_gradients_match_fp32_oracle(sfr, dev):
    n = 50_005
    g = gen(10)
    theta0 = torch.randn(n, generator=g) * 0.02
    ref = flat_loop(n, theta0, "adamw", dict(lr=1e-4), "dit", 0.9999)
    hp, p = engine_for(sfr, dev, n, theta0, "adamw", dict(lr=1e-4), "dit", 0.9999)
    for _ in range(4):
        gr = (torch.randn(n, generator=g) * 0.1).bfloat16()
        ref.remain_step({"w": gr.float()}, max_norm=1.0)
        hp.remain_step(p, gr.to(dev), max_norm=1.0)
    assert close(p, ref.flat("p")) and close(hp.slow, ref.flat("slow"))


# =============================================================================== boundary behaviour
def test_flat_params_adopts_model_and_gathers(sfr, dev):
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 8), torch.nn.ReLU(), torch.nn.Linear(8, 3)).to(dev)
    ref_sd = {k: v.clone() for k, v in model.state_dict().items()}
    fp = sfr.FlatParams(model, dev)
    for name, prm in model.named_parameters():
        assert torch.equal(prm, ref_sd[name])
        assert prm.data_ptr() == fp.layout.view(fp.p, name).data_ptr()
    x = torch.randn(5, 16, device=dev)
    model(x).sum().backward()
    g_views = torch.cat([prm.grad.reshape(-1) for prm in model.parameters()])
    assert torch.equal(fp.collect_grads(), g_views) and fp.g.abs().sum() > 0
    # gather path: autograd-allocated grads -> flat vector by one kernel
    model2 = torch.nn.Sequential(torch.nn.Linear(16, 8), torch.nn.ReLU(), torch.nn.Linear(8, 3)).to(dev)
    model2.load_state_dict(ref_sd)
    fp2 = sfr.FlatParams(model2, dev, grads_as_views=False)
    model2(x).sum().backward()
    assert torch.equal(fp2.collect_grads(), g_views)


def test_errors_are_loud(sfr, dev):
    capi = sfr.capi
    a = torch.zeros(64, device=dev)
    with pytest.raises(capi.SfrError) as e:           # CPU tensors never reach a kernel
        capi.fisher_accum(torch.zeros(64), torch.zeros(64), 1.0)
    assert e.value.code == capi.ERR_NO_DEVICE
    with pytest.raises(capi.SfrError) as e:           # misaligned base pointer
        capi.fisher_accum(a[1:33], a[1:33].clone(), 1.0)
    assert e.value.code == capi.ERR_ALIGN
    with pytest.raises(capi.SfrError) as e:           # size mismatch
        capi.fisher_accum(a, torch.zeros(32, device=dev), 1.0)
    assert e.value.code == capi.ERR_ARG
    with pytest.raises(capi.SfrError):
        capi.ratio_mask(a, a, 1.0, torch.zeros(64, dtype=torch.float32, device=dev))
    with pytest.raises(capi.SfrError):
        sfr.HotPath(10, "cpu", sfr.OptConfig())
    # empty input is a no-op
    capi.fisher_accum(torch.zeros(0, device=dev), torch.zeros(0, device=dev), 1.0)
    sm, major, _ = capi.device_info()
    assert major == 10 and sm > 0


# =============================================================================== §8f n2 / n3 consumers
@pytest.mark.parametrize("n", [5, 4099, 300_001])
def test_ewc_penalty_matches_autograd(sfr, dev, n):
    g = gen(n + 3)
    p, ps = torch.randn(n, generator=g) * 0.02, torch.randn(n, generator=g) * 0.02
    fisher = torch.randn(n, generator=g).pow(2) * 1e-3
    g0 = torch.randn(n, generator=g) * 0.01
    lmbda = 50.0
    grads, ewc = O.ewc_penalty_grads({"w": p}, {"w": ps}, {"w": fisher}, lmbda)
    hp = sfr.HotPath(n, dev, sfr.OptConfig())
    gd = g0.to(dev)
    pen = hp.ewc_penalty(p.to(dev), ps.to(dev), fisher.to(dev), gd, lmbda)
    assert bits_equal(gd.cpu(), g0 + grads["w"])                 # model grad + EWC grad, as autograd sums them
    assert abs(pen.item() - ewc.item()) <= 1e-6 * abs(ewc.item())


@pytest.mark.parametrize("n", [7, 10_003, 1_000_003])
@pytest.mark.parametrize("frac", [0.001, 0.3, 0.9])
def test_proximal_shrink_bit_exact(sfr, dev, n, frac):
    g = gen(n + 17)
    p0 = torch.randn(n, generator=g) * 0.02
    p = p0 + torch.randn(n, generator=g) * 1e-3
    same = torch.rand(n, generator=g) < 0.1
    p[same] = p0[same]                                            # exact zeros of theta - theta0
    k = max(1, int(n * frac))
    ref = p.clone()
    thr = O.proximal_step_([ref], [p0.clone()], k)
    hp = sfr.HotPath(n, dev, sfr.OptConfig())
    pd = p.to(dev)
    got_thr = hp.proximal_shrink(pd, p0.to(dev), k)
    assert got_thr.item() == thr.item()
    assert bits_equal(pd.cpu(), ref)


# =============================================================================== CUDA-graph replay
@pytest.mark.parametrize("opt,kw,ema_mode,ema_a", [CASES[0], CASES[3], CASES[5]])
@pytest.mark.parametrize("launch", ["coop", "split"])
def test_k3_cuda_graph_replay_advances_device_step(sfr, dev, opt, kw, ema_mode, ema_a, launch):
    """forget_step + remain_step captured once and replayed: the optimizer step (Adam bias corrections,
    SGD first-step buffer init) must advance on the device at every replay."""
    n = 50_003
    g = gen(77)
    theta0 = torch.randn(n, generator=g) * 0.02
    mask = torch.rand(n, generator=g) < 0.35
    ref = flat_loop(n, theta0, opt, kw, ema_mode, ema_a)
    hp, p = engine_for(sfr, dev, n, theta0, opt, kw, ema_mode, ema_a)
    hp.set_buffer("mask", mask.to(torch.uint8).to(dev))
    if launch == "split":
        hp.coop_max_elems = 0
    hp.enable_graph_replay()
    gf_s, gr_s = torch.zeros(n, device=dev), torch.zeros(n, device=dev)

    def body():
        hp.forget_step(p, gf_s, max_norm=1.0)
        hp.remain_step(p, gr_s, ema=True)

    steps = [(torch.randn(n, generator=g) * (3.0 if i % 2 else 0.01), torch.randn(n, generator=g) * 0.1)
             for i in range(6)]
    # eager step 0 on a side stream (warm-up, also exercises the device counter outside a graph)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        gf_s.copy_(steps[0][0]); gr_s.copy_(steps[0][1])
        body()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    gf_s.copy_(steps[1][0]); gr_s.copy_(steps[1][1])
    with torch.cuda.graph(graph):
        body()
    graph.replay()                                     # capture does not execute: first real run of step 1
    for gf, gr in steps[2:]:
        gf_s.copy_(gf); gr_s.copy_(gr)
        graph.replay()
    for gf, gr in steps:
        ref.forget_step({"w": gf}, mask={"w": mask}, max_norm=1.0)
        ref.remain_step({"w": gr}, ema=True)
    assert int(hp.step_dev) == 2 * len(steps)
    assert close(p, ref.flat("p"))
    if ema_mode != "none":
        assert close(hp.slow, ref.flat("slow"))


# =============================================================================== own bounds check
def _guarded(dev, n, dtype, fill):
    """A 16-byte-aligned n-element view in the middle of a larger allocation whose remainder holds a
    sentinel: any out-of-bounds WRITE of a kernel lands in a guard zone (compute-sanitizer is closed on
    this pool)."""
    item = torch.empty(0, dtype=dtype).element_size()
    guard = 4096 // item
    n_alloc = (n + 15) // 16 * 16
    buf = torch.full((guard + n_alloc + guard,), fill, dtype=dtype, device=dev)
    view = buf[guard:guard + n]
    return buf, view, guard


def _guards_intact(buf, view_n, guard, fill):
    head, tail = buf[:guard], buf[guard + view_n:]
    return bool((head == fill).all()) and bool((tail == fill).all())


@pytest.mark.parametrize("n", [1, 3, 5, 1023, 4097, 8191, 8193, 100_003])
def test_no_out_of_bounds_writes(sfr, dev, n):
    capi = sfr.capi
    g = gen(n)
    S = 12345.0
    bufs = {}
    for name in ("acc", "p", "m", "v", "ema", "gz"):
        bufs[name] = _guarded(dev, n, torch.float32, S)
        bufs[name][1].copy_(torch.randn(n, generator=g) * 0.01)
    bufs["v"][1].abs_()
    mb, mask, mg = _guarded(dev, n, torch.uint8, 77)
    tb, topk, tg = _guarded(dev, n, torch.uint8, 77)
    pb, p16, pg = _guarded(dev, n, torch.bfloat16, 3.0)
    grad = (torch.randn(n, generator=g) * 0.1).to(dev)
    ff = torch.randn(n, generator=g).pow(2).to(dev)
    rf = torch.randn(n, generator=g).pow(2).to(dev)
    capi.fisher_accum(bufs["acc"][1], grad, 3.0)
    capi.ratio_mask(ff, rf, 1.0, mask)
    hp = sfr.HotPath(n, dev, sfr.OptConfig(kind="adamw", lr=1e-3, weight_decay=0.01), ema_mode="dit", ema_a=0.99)
    hp.set_buffer("m", bufs["m"][1]); hp.set_buffer("v", bufs["v"][1]); hp.set_buffer("slow", bufs["ema"][1])
    hp.set_buffer("mask", mask)
    bufs["gz"][1].copy_(grad)
    hp.forget_step(bufs["p"][1], bufs["gz"][1], max_norm=1.0, zero_grad=True, p_bf16=p16)
    bufs["gz"][1].copy_(grad)
    hp.remain_step(bufs["p"][1], bufs["gz"][1], ema=True, zero_grad=True, p_bf16=p16)
    quant = (grad * 8).round() / 8                      # heavy ties -> ordered apply path
    hp.topk_mask(quant, max(1, n // 3), out=topk)
    hp.topk_mask(grad, n // 2, out=topk)
    capi.ewc_penalty(bufs["p"][1], bufs["ema"][1], ff, bufs["gz"][1], 2.0)
    hp.proximal_shrink(bufs["p"][1], bufs["ema"][1], max(1, n // 2))
    torch.cuda.synchronize()
    for name, (buf, view, guard) in bufs.items():
        assert _guards_intact(buf, n, guard, S), f"{name}: write outside [0, n)"
    assert _guards_intact(mb, n, mg, 77) and _guards_intact(tb, n, tg, 77) and _guards_intact(pb, n, pg, 3.0)


# =============================================================================== property tests (hypothesis)
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(data=st.data())
def test_k2b_property_random_distributions(sfr, dev, data):
    """Exact top-k over random sizes / value distributions / k: bit-exact against the stable-argsort oracle."""
    n = data.draw(st.integers(1, 60_000))
    seed = data.draw(st.integers(0, 2 ** 31 - 1))
    kind = data.draw(st.sampled_from(["gauss", "lognormal", "quantized", "sparse", "denormal", "mixed_special"]))
    g = gen(seed)
    if kind == "gauss":
        x = torch.randn(n, generator=g) * 10 ** data.draw(st.integers(-20, 10))
    elif kind == "lognormal":
        x = torch.exp(torch.randn(n, generator=g) * 8) * torch.where(torch.rand(n, generator=g) < 0.5, -1.0, 1.0)
    elif kind == "quantized":
        x = torch.randint(-5, 6, (n,), generator=g).float() * 0.5
    elif kind == "sparse":
        x = torch.randn(n, generator=g) * (torch.rand(n, generator=g) < 0.05)
    elif kind == "denormal":
        x = torch.randn(n, generator=g) * 1e-41
    else:
        x = torch.randn(n, generator=g)
        idx = torch.randint(0, n, (max(1, n // 50),), generator=g)
        x[idx] = torch.tensor([float("inf"), float("-inf"), float("nan"), 0.0, -0.0])[torch.randint(0, 5, idx.shape, generator=g)]
    n_num = int((~x.isnan()).sum())
    k = data.draw(st.integers(0, n_num))              # NaN never ranks ahead of a number
    mask, _ = run_topk(sfr, dev, x, k)
    assert torch.equal(mask, O.topk_mask_flat(x, k))


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(n=st.integers(1, 40_000), seed=st.integers(0, 2 ** 31 - 1), rows=st.integers(1, 4),
       divisor=st.sampled_from([1.0, 3.0, 50.0, 2000.0]))
def test_k1_k2a_property_bit_exact(sfr, dev, n, seed, rows, divisor):
    g = gen(seed)
    n_pad = (n + 7) // 8 * 8
    gr = torch.randn(rows, n_pad, generator=g)[:, :n] * 10 ** float(torch.randint(-8, 3, (1,), generator=g))
    acc0 = torch.rand(n, generator=g) * 1e-6
    ref = acc0.clone()
    for r in gr:
        O.flat_fisher_accum(ref, r, divisor)
    acc = acc0.to(dev)
    full = torch.zeros(rows, n_pad, device=dev)
    full[:, :n] = gr.to(dev)
    sfr.capi.fisher_accum(acc, full[:, :n] if rows > 1 else full[0, :n].clone(), divisor)
    assert bits_equal(acc.cpu(), ref)
    rf = torch.rand(n, generator=g) * 1e-6
    th = float(torch.rand(1, generator=g) * 3)
    mask = torch.empty(n, dtype=torch.uint8, device=dev)
    sfr.capi.ratio_mask(acc, rf.to(dev), th, mask)
    assert torch.equal(mask.cpu().bool(), O.flat_ratio_mask(ref, rf, th))


# =============================================================================== full-size properties
N3_FULL = 675_129_632          # DiT-XL/2 (BASELINE.json config 3): no CPU oracle at this size, exact identities instead


def test_full_size_dit_properties(sfr, dev):
    capi = sfr.capi
    n = N3_FULL
    gd = torch.Generator(device=dev).manual_seed(0)
    g = torch.empty(n, device=dev).normal_(0, 1e-2, generator=gd)
    # K1: two accumulations with divisor 2 give g*g exactly (x/2 is exact; x/2 + x/2 = x)
    acc = torch.zeros(n, device=dev)
    capi.fisher_accum(acc, g, 2.0)
    capi.fisher_accum(acc, g, 2.0)
    assert torch.equal(acc, g * g)
    # K2a: equal Fishers -> ratio exactly 1 -> every element kept at th = 1, none at th just above 1;
    # zero count + ones == n (checksum of checksums)
    hp = sfr.HotPath(n, dev, sfr.OptConfig(kind="sgd", lr=0.0, momentum=0.9, weight_decay=0.0))
    hp.set_buffer("forget_fisher", acc)
    hp.set_buffer("remain_fisher", acc)
    mask = hp.ratio_mask(1.0)
    assert int(hp.zero_count[0]) == 0 and int(mask.sum(dtype=torch.int64)) == n
    mask = hp.ratio_mask(1.0000001)
    assert int(hp.zero_count[0]) == n and int(mask.sum(dtype=torch.int64)) == 0
    hp.set_buffer("remain_fisher", torch.full((n,), 1e-4, device=dev))
    mask = hp.ratio_mask(1.0)
    ones = int(mask.sum(dtype=torch.int64))
    assert int(hp.zero_count[0]) + ones == n and 0 < ones < n
    assert torch.equal(mask.bool(), (acc + 1e-15) >= (torch.full_like(acc, 1e-4) + 1e-15))   # (a/b >= 1) == (a >= b), b > 0
    # clip norm: masked + complement-masked sums of squares add up to the unmasked one
    s_all, s_m, s_c = (torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(3))
    capi.masked_sumsq(g, None, s_all)
    capi.masked_sumsq(g, mask, s_m)
    capi.masked_sumsq(g, 1 - mask, s_c)
    assert abs((s_m + s_c - s_all).item()) <= 1e-9 * s_all.item()
    # K3: a step with lr = 0 leaves the weights bit-identical and builds the momentum buffer = g * mask
    p = torch.empty(n, device=dev).normal_(0, 0.02, generator=gd)
    p0 = p.clone()
    hp.set_buffer("mask", mask)
    hp.forget_step(p, g, lr=0.0)
    assert torch.equal(p, p0) and torch.equal(hp.m, g * mask)
    del p0, acc
    # K2b: exact count, ordering property, idempotence
    k = n // 5
    sel = hp.topk_mask(g, k, out=torch.empty(n, dtype=torch.uint8, device=dev)).bool()
    assert int(sel.sum(dtype=torch.int64)) == k
    a = g.abs()
    assert a[sel].min() >= a[~sel].max()
    st = hp.select_state()
    assert st.count_gt + st.tie_budget == k and st.tie_budget <= st.count_eq


def test_indexing_past_2_pow_31_and_degenerate_selects():
    """n = 2^31 + 4099 elements through every kernel, and all-zero / 99.96 %-zero / two-valued selects at N3
    (tools/stress.py; ~60 GB of device memory, ~20 s)."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "stress.py")], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"ok": true' in r.stdout
