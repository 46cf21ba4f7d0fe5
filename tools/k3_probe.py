"""Probe: K3 forget step alone, with and without the clip coefficient, with / without the L2 flush."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfron_b200 as sfr
from sfron_b200 import capi
n = 675_129_632
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
hp = sfr.HotPath(n, dev, sfr.OptConfig(kind="adamw", lr=1e-4), ema_mode="dit", ema_a=0.9999)
p = torch.empty(n, device=dev).normal_(0, 0.02, generator=gen)
g = torch.empty(n, device=dev).normal_(0, 1e-2, generator=gen)
hp.init_slow(p); hp.mask.copy_((torch.rand(n, device=dev, generator=gen) < 0.5).to(torch.uint8)); hp.mark_mask_ready()
hp.remain_step(p, g, ema=False)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
hp.sumsq.fill_(float(n) * 1e-4 * 0.5)

def k3(clip, mask=True, ema=False, scratch=False):
    hp.step_count += 1
    flags = (capi.F_MASK if mask else 0)
    a = hp._args(flags, ema, 1.0 if clip else None, None)
    capi.fused_update(p, g, hp.m, hp.v, hp.mask if mask else None, hp.slow if ema else None, a,
                      clip_sumsq=hp.sumsq if clip else None, consts_scratch=hp._consts_dev if scratch else None)

def timeit(fn, do_flush):
    ts = []
    for i in range(9):
        if do_flush: flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[4]

for do_flush in (True, False):
    for name, fn in (("masked,noclip", lambda: k3(False)), ("masked,clip", lambda: k3(True)),
                     ("nomask,noclip", lambda: k3(False, mask=False)), ("nomask,clip", lambda: k3(True, mask=False)),
                     ("nomask,noclip,ema", lambda: k3(False, mask=False, ema=True)),
                     ("masked,clip,scratch", lambda: k3(True, scratch=True)),
                     ("masked,noclip,scratch", lambda: k3(False, scratch=True))):
        ms = timeit(fn, do_flush)
        bpe = 28 + (1 if "masked" in name else 0) + (8 if "ema" in name else 0)
        print(f"flush={do_flush!s:5s} {name:20s} {ms:.4f} ms  {bpe*n/ms/1e6:.1f} GB/s")
# back-to-back pair as in the step: sumsq then K3
def pair():
    hp.sumsq.zero_(); capi.masked_sumsq(g, hp.mask, hp.sumsq); k3(True)
print("sumsq+K3(clip) pair", timeit(pair, False))
def pair2():
    capi.masked_sumsq(g, hp.mask, hp.zero_count.view(torch.float64)[:1]); k3(False)
print("sumsq+K3(noclip) pair", timeit(pair2, False))
