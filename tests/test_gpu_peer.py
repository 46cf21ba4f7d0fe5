"""Parity of the fused exchange kernels (csrc/peer.cu) on ONE GPU: the `world` ranks of a data-parallel
step are emulated by `world` sets of buffers on the same device — the P2P transport only needs `world`
mapped addresses per buffer, and those may all be local.  Each emulated rank's kernels are launched one
after another through the C ABI (no rank waits for another, so no barrier is involved; the barrier itself
is covered with real processes in tests/test_gpu_dist.py and with world = 1 here).

Checked against the CPU oracle: the data-parallel gradient is the reference's DataParallel reduce_add
(sum in device order) divided by the number of ranks (oracle.dp_reduce); K1 on it is bit-exact, the
sharded K3 is within 1e-6 of `FlatReferenceLoop`, every rank ends with identical full weight vectors,
and an unclipped sharded step is bit-identical to the single-vector kernel fed the same gradient.
"""
import pytest
import torch

from conftest import bits_equal
from oracle import sfron_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-6


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
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    rms = b.pow(2).mean().sqrt().item() if b.numel() else 0.0
    return not bool(((a - b).abs() > rtol * (b.abs() + rms)).any())


class Emulated:
    """`world` ranks on one device: full-vector gradient / weight buffers per rank, equal 16-aligned shards."""

    def __init__(self, sfr, dev, n, world, g_dtype=torch.float32, seed=0):
        from sfron_b200 import capi
        self.capi, self.dev, self.n, self.world = capi, dev, n, world
        self.per = (n + world * 16 - 1) // (world * 16) * 16
        self.n_pad = self.per * world
        self.bounds = [(min(r * self.per, n), min((r + 1) * self.per, n)) for r in range(world)]
        gen = torch.Generator().manual_seed(seed)
        self.theta0 = torch.randn(n, generator=gen) * 0.02
        self.g_host = [(torch.randn(n, generator=gen) * 0.05).to(g_dtype) for _ in range(world)]
        self.mask_host = torch.rand(n, generator=gen) < 0.5
        self.g = [self._padded(h, g_dtype) for h in self.g_host]
        self.w = [self._padded(self.theta0, torch.float32) for _ in range(world)]
        self.w16 = [self._padded(self.theta0, torch.bfloat16) for _ in range(world)]
        self.g_buf = capi.peer_buf([t.data_ptr() for t in self.g])
        self.w_buf = capi.peer_buf([t.data_ptr() for t in self.w])
        self.w16_buf = capi.peer_buf([t.data_ptr() for t in self.w16])

    def _padded(self, host, dtype):
        t = torch.zeros(self.n_pad, dtype=dtype, device=self.dev)
        t[:self.n].copy_(host.to(dtype))
        return t

    def geom(self, r):
        lo, hi = self.bounds[r]
        return self.capi.PeerGeom(self.world, r, lo, hi - lo)

    def reduced_oracle(self, average=True):
        return O.dp_reduce([h.float() for h in self.g_host], average)


TRANSPORTS = ["p2p", "tma"]      # load/store threads | TMA bulk copies through shared memory: same sums, same order


def xp(sfr, name):
    return {"p2p": sfr.capi.XP_P2P, "tma": sfr.capi.XP_TMA}[name]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("n", [16 * 8 * 5, 4099, 1_000_003])
@pytest.mark.parametrize("g_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("transport", TRANSPORTS)
def test_peer_reduce_k1_norm_bit_exact(sfr, dev, world, n, g_dtype, transport):
    """reduce-scatter + K1 + clip norm in one kernel == oracle on the DataParallel-reduced gradient."""
    capi = sfr.capi
    em = Emulated(sfr, dev, n, world, g_dtype)
    gbar = em.reduced_oracle()
    acc_ref = torch.full((n,), 1e-7)
    O.flat_fisher_accum(acc_ref, gbar, 7.0)
    sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
    got_red, got_acc = torch.empty(n), torch.empty(n)
    for r in range(world):
        lo, hi = em.bounds[r]
        if hi == lo:
            continue
        red = torch.full((hi - lo,), float("nan"), device=dev)
        acc = torch.full((hi - lo,), 1e-7, device=dev)
        mask = em.mask_host[lo:hi].to(dev)
        capi.peer_reduce(em.g_buf, g_dtype, em.geom(r), xp(sfr, transport), True, g_red=red, mask=mask, sumsq=sumsq,
                         fisher=acc, fisher_divisor=7.0)
        got_red[lo:hi] = red.cpu()
        got_acc[lo:hi] = acc.cpu()
    assert bits_equal(got_red, gbar), "reduced gradient differs from the sequential rank-order sum / world"
    assert bits_equal(got_acc, acc_ref), "fused K1 not bit-exact"
    want = (gbar.double() * em.mask_host.double()).pow(2).sum().item()
    assert abs(sumsq.item() - want) <= 2e-7 * max(want, 1e-30)     # fp32 partials of 4, folded in double


@pytest.mark.parametrize("world", [2, 8])
@pytest.mark.parametrize("kind,kw,ema_mode,ema_a", [
    ("adamw", dict(lr=1e-4, weight_decay=0.0), "dit", 0.9999),
    ("adam", dict(lr=2e-4), "ddpm", 1e-4),
    ("sgd", dict(lr=0.01, momentum=0.9, weight_decay=5e-4), "slowfast", 0.9),
])
@pytest.mark.parametrize("g_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("transport", TRANSPORTS)
def test_peer_sharded_steps_match_oracle(sfr, dev, world, kind, kw, ema_mode, ema_a, g_dtype, transport):
    """forget step (mask, clip; gradient from the local reduced shard) then remain step + EMA (gradient
    reduced on the fly), weights pushed to every rank: == FlatReferenceLoop on the reduced gradients."""
    capi = sfr.capi
    n = 250_007
    em = Emulated(sfr, dev, n, world, g_dtype, seed=3)
    em2 = Emulated(sfr, dev, n, world, g_dtype, seed=4)          # the remain pass's gradients
    ref = O.FlatReferenceLoop({"w": (n,)}, {"w": em.theta0}, kind, kw, ema_mode=ema_mode, ema_a=ema_a)
    ref.forget_step({"w": em.reduced_oracle()}, mask={"w": em.mask_host}, max_norm=1.0)
    ref.remain_step({"w": em2.reduced_oracle()}, ema=True)

    opt = sfr.OptConfig(kind=kind, **kw)
    hps = []
    for r in range(world):
        lo, hi = em.bounds[r]
        hp = sfr.HotPath(hi - lo, dev, opt, ema_mode=ema_mode, ema_a=ema_a)
        hp.set_buffer("mask", em.mask_host[lo:hi].to(dev).to(torch.uint8))
        hp.init_slow(em.w[r][lo:hi])
        hps.append(hp)
    # ---- forget: reduce (+ masked sum of squares) on every rank, "all-reduce" of the norm, then K3 + push
    sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
    reds = []
    for r in range(world):
        lo, hi = em.bounds[r]
        red = torch.empty(hi - lo, device=dev)
        capi.peer_reduce(em.g_buf, g_dtype, em.geom(r), xp(sfr, transport), True, g_red=red, mask=hps[r].mask, sumsq=sumsq)
        reds.append(red)
    sgd = kind == "sgd"
    for r in range(world):
        lo, hi = em.bounds[r]
        hp = hps[r]
        flags = capi.F_MASK | (capi.F_SGD_FIRST_STEP if sgd else 0)
        hp.step_count = 1
        a = hp._args(flags, False, 1.0, None)
        capi.peer_fused_update(em.w[r][lo:hi], em.geom(r), a, g_red=reds[r], m=hp.m, v=None if sgd else hp.v,
                               mask=hp.mask, bc_f32=em.w_buf, bc_bf16=em.w16_buf, bc_transport=xp(sfr, transport),
                               clip_sumsq=sumsq, consts_scratch=hp._consts_dev)
    # ---- remain: gradient reduced inside the update kernel, EMA, push
    for r in range(world):
        lo, hi = em.bounds[r]
        hp = hps[r]
        hp.step_count = 2
        a = hp._args(0, True, None, None)
        capi.peer_fused_update(em.w[r][lo:hi], em.geom(r), a, g=em2.g_buf, g_dtype=g_dtype, m=hp.m,
                               v=None if sgd else hp.v, ema=hp.slow, bc_f32=em.w_buf, bc_bf16=em.w16_buf,
                               g_transport=xp(sfr, transport), bc_transport=xp(sfr, transport))
    torch.cuda.synchronize()
    want = ref.flat("p")
    for r in range(world):
        assert bits_equal(em.w[r][:n].cpu(), em.w[0][:n].cpu()), f"rank {r} holds different weights"
        assert torch.equal(em.w16[r][:n].cpu(), em.w[0][:n].bfloat16().cpu()), f"rank {r}: bf16 working copy"
    assert close(em.w[0][:n], want), "sharded data-parallel update differs from the oracle"
    slow = torch.cat([hps[r].slow for r in range(world)])
    assert close(slow, ref.flat("slow"))


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("transport", TRANSPORTS)
def test_peer_unclipped_step_is_bit_identical_to_single_vector_kernel(sfr, dev, world, transport):
    capi = sfr.capi
    n = 300_013
    em = Emulated(sfr, dev, n, world, seed=9)
    opt = sfr.OptConfig(kind="adamw", lr=1e-4)
    one = sfr.HotPath(n, dev, opt, ema_mode="dit", ema_a=0.9999)
    p_one = em.theta0.to(dev).clone()
    one.init_slow(p_one)
    one.remain_step(p_one, em.reduced_oracle().to(dev), ema=True)
    for r in range(world):
        lo, hi = em.bounds[r]
        hp = sfr.HotPath(hi - lo, dev, opt, ema_mode="dit", ema_a=0.9999)
        hp.init_slow(em.w[r][lo:hi])
        hp.step_count = 1
        capi.peer_fused_update(em.w[r][lo:hi], em.geom(r), hp._args(0, True, None, None), g=em.g_buf, m=hp.m, v=hp.v,
                               ema=hp.slow, bc_f32=em.w_buf, g_transport=xp(sfr, transport),
                               bc_transport=xp(sfr, transport))
        assert bits_equal(hp.slow.cpu(), one.slow[lo:hi].cpu())
    assert bits_equal(em.w[world - 1][:n].cpu(), p_one.cpu())


def test_peer_broadcast_and_argument_errors(sfr, dev):
    capi = sfr.capi
    em = Emulated(sfr, dev, 5003, 4)
    for r in range(4):
        lo, hi = em.bounds[r]
        capi.peer_broadcast(torch.full((hi - lo,), float(r + 1), device=dev), em.w_buf, em.geom(r), capi.XP_P2P)
        capi.peer_broadcast(torch.full((hi - lo,), float(r + 1), device=dev).bfloat16(), em.w16_buf, em.geom(r),
                            capi.XP_P2P)
    want = torch.cat([torch.full((hi - lo,), float(r + 1)) for r, (lo, hi) in enumerate(em.bounds)])
    for r in range(4):
        assert torch.equal(em.w[r][:5003].cpu(), want) and torch.equal(em.w16[r][:5003].float().cpu(), want)
    with pytest.raises(capi.SfrError):                    # multimem without a multicast address
        capi.peer_reduce(em.g_buf, torch.float32, em.geom(0), capi.XP_MULTIMEM, True,
                         g_red=torch.empty(em.bounds[0][1], device=dev))
    with pytest.raises(capi.SfrError):                    # shard start not a multiple of 16
        capi.peer_reduce(em.g_buf, torch.float32, capi.PeerGeom(4, 0, 8, 16), capi.XP_P2P, True,
                         g_red=torch.empty(16, device=dev))
    with pytest.raises(capi.SfrError):                    # nothing to produce
        capi.peer_reduce(em.g_buf, torch.float32, em.geom(0), capi.XP_P2P, True)
    with pytest.raises(capi.SfrError):                    # one fused kernel = one transport for both directions
        hp = sfr.HotPath(em.bounds[0][1], dev, sfr.OptConfig(kind="adamw", lr=1e-4))
        hp.step_count = 1
        capi.peer_fused_update(em.w[0][:em.bounds[0][1]], em.geom(0), hp._args(0, False, None, None), g=em.g_buf,
                               m=hp.m, v=hp.v, bc_f32=em.w_buf, g_transport=capi.XP_TMA, bc_transport=capi.XP_P2P)


def test_peer_barrier_world_one_sums_its_payload(sfr, dev):
    """The barrier with one rank: epoch advances, the payload comes back as its own sum, no timeout."""
    capi = sfr.capi
    pad = torch.zeros(capi.peer_pad_bytes() // 8, dtype=torch.int64, device=dev)
    buf = capi.peer_buf([pad.data_ptr()])
    vals = torch.tensor([1.5, -2.0, 1e300], dtype=torch.float64, device=dev)
    sums = torch.zeros(3, dtype=torch.float64, device=dev)
    for _ in range(3):
        capi.peer_barrier(buf, 1, 0, vals, sums)
    torch.cuda.synchronize()
    assert torch.equal(sums.cpu(), vals.cpu()) and int(pad[0]) == 3 and int(pad[1]) == 0


def test_kernels_follow_the_tensors_device_not_the_current_one(sfr, dev):
    """a HotPath built for cuda:1 must run on cuda:1 while cuda:0 is current."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    other = torch.device("cuda:1")
    assert torch.cuda.current_device() == 0
    n = 100_003
    g = torch.randn(n, generator=torch.Generator().manual_seed(0))
    hp = sfr.HotPath(n, other, sfr.OptConfig(kind="adamw", lr=1e-4))
    hp.fisher_accumulate("forget", g.to(other), 3.0)
    hp.topk_mask(g.to(other), n // 2)                     # pass 0 needs the 128 KB shared-memory opt-in on cuda:1
    assert torch.cuda.current_device() == 0
    ref = torch.zeros(n)
    O.flat_fisher_accum(ref, g, 3)
    assert bits_equal(hp.forget_fisher.cpu(), ref)
    assert torch.equal(hp.mask.cpu(), O.topk_mask_flat(g, n // 2))
    with pytest.raises(sfr.capi.SfrError):
        sfr.capi.fisher_accum(hp.forget_fisher, g.to(dev), 3.0)     # tensors of one call on two devices
