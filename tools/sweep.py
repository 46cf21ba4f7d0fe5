#!/usr/bin/env python
"""Kernel sweep over flat parameter vectors (BASELINE.json configs[4]): every kernel of the hot
path, timed alone with CUDA events, at n in {1e7 .. 2e9}, fp32 and bf16 gradients.

    python tools/sweep.py [--sizes 10000000,38632323,...] [--iters 10] [--out gpurun_out/sweep.jsonl]

One JSON line per (kernel, n, dtype): algorithmic bytes, ms, GB/s, fraction of the measured HBM
peak.  Inputs per SURVEY.md §8d config 5: g ~ N(0,1)*1e-2, F = mean of 8 squared draws (chi^2-like,
realistic exponent clustering), p ~ N(0, 0.02).  Buffers < 126 MB would sit in L2, so an L2 flush
(a 256 MB write) precedes every timed launch at every size.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import sfron_b200 as sfr  # noqa: E402
from sfron_b200 import capi  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


class Timer:
    def __init__(self, dev, iters):
        self.iters = iters
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def __call__(self, fn, setup=None):
        """Median device time (ms) of fn() over `iters` launches, L2 flushed before each."""
        times = []
        for i in range(self.iters + 2):
            if setup is not None:
                setup()
            self.flush.fill_(i & 0xFF)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            if i >= 2:
                times.append(a.elapsed_time(b))
        times.sort()
        return times[len(times) // 2]


def chi2_like(n, dev, gen, scale):
    acc = torch.zeros(n, device=dev)
    chunk = torch.empty(n, device=dev)
    for _ in range(8):
        chunk.normal_(0, scale, generator=gen)
        acc.addcmul_(chunk, chunk, value=1 / 8)
    del chunk
    return acc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="10000000,38632323,100000000,675129632,859520964,2000000000")
    ap.add_argument("--iters", type=int, default=9)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--skip-select", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    capi.load()
    pk = peak()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    out = open(args.out, "w")
    gen = torch.Generator(device=dev).manual_seed(1234)
    timer = Timer(dev, args.iters)

    def emit(kernel, n, dtype, bpe, ms, **extra):
        gbs = bpe * n / (ms * 1e-3) / 1e9
        rec = dict(kernel=kernel, n=n, grad_dtype=dtype, bytes_per_elem=bpe, ms=round(ms, 4),
                   GBps=round(gbs, 1), frac_of_measured_peak=round(gbs / pk, 4), **extra)
        out.write(json.dumps(rec) + "\n")
        out.flush()
        print(json.dumps(rec), flush=True)

    for n in [int(x) for x in args.sizes.split(",")]:
        hp = sfr.HotPath(n, dev, sfr.OptConfig(kind="adamw", lr=1e-4), ema_mode="dit", ema_a=0.9999)
        p = torch.empty(n, device=dev).normal_(0, 0.02, generator=gen)
        g32 = torch.empty(n, device=dev).normal_(0, 1e-2, generator=gen)
        hp.init_slow(p)
        hp.set_buffer("forget_fisher", chi2_like(n, dev, gen, 1e-2))
        hp.set_buffer("remain_fisher", chi2_like(n, dev, gen, 1e-2))
        hp.ratio_mask(1.0)
        for _ in range(3):   # m, v from warm-up steps
            hp.remain_step(p, g32, ema=False)
        acc = torch.zeros(n, device=dev)
        for dtype, g in (("f32", g32), ("bf16", g32.bfloat16())):
            sg = 4 if dtype == "f32" else 2
            emit("fisher_accum", n, dtype, 8 + sg, timer(lambda: capi.fisher_accum(acc, g, 2000.0)))
            emit("masked_sumsq", n, dtype, sg + 1, timer(lambda: capi.masked_sumsq(g, hp.mask, hp.sumsq)))
            emit("fused_update_adamw_masked_clip", n, dtype, 25 + sg,
                 timer(lambda: hp.forget_step(p, g, max_norm=None)))
            emit("fused_update_adamw_ema", n, dtype, 32 + sg, timer(lambda: hp.remain_step(p, g, ema=True)))
            if n <= 120_000_000:
                # clipped forget step (norm + update): one cooperative launch vs memset + norm + scalar prep + update
                hp.coop_max_elems, keep0 = 1 << 40, hp.coop_max_elems
                emit("clipped_forget_step_one_cooperative_launch", n, dtype, 30 + 2 * sg,
                     timer(lambda: hp.forget_step(p, g, max_norm=1.0)),
                     note="sfr_clipped_update: zero + masked sum of squares + grid barrier + AdamW; g is read twice")
                keep = hp.coop_max_elems
                hp.coop_max_elems = 0
                emit("clipped_forget_step_four_launches", n, dtype, 30 + 2 * sg,
                     timer(lambda: hp.forget_step(p, g, max_norm=1.0)))
                hp.coop_max_elems = keep0
        if n <= 1_000_000_000:   # per-sample FIM, 4 rows (16 B/elem of gradient rows)
            n_pad = (n + 7) // 8 * 8          # row stride must keep every row 16-byte aligned
            rows = torch.empty(4, n_pad, device=dev).normal_(0, 1e-2, generator=gen)[:, :n]
            emit("fisher_accum_rows4", n, "f32", 8 + 16, timer(lambda: capi.fisher_accum(acc, rows, 50.0)))
            del rows
        emit("ratio_mask", n, "-", 9, timer(lambda: hp.ratio_mask(1.0)))
        ths = [0.5, 1.0, 3.0, 5.0, 10.0]
        emit("ratio_mask_x5_thresholds", n, "-", 8 + 5, timer(lambda: hp.ratio_masks(ths)),
             note="one pass for DiT/generate_mask.py's 5 thresholds (reference: 5 passes, 45 B/elem)")
        # SGD variant (Classification): p,g,buf,prev
        hs = sfr.HotPath(n, dev, sfr.OptConfig(kind="sgd", lr=0.01, momentum=0.9, weight_decay=5e-4),
                         ema_mode="slowfast", ema_a=0.9)
        hs.init_slow(p)
        hs.set_buffer("mask", hp.mask)
        hs.remain_step(p, g32, ema=False)
        emit("fused_update_sgd_masked", n, "f32", 21, timer(lambda: hs.forget_step(p, g32, max_norm=None)))
        emit("fused_update_sgd_slowfast", n, "f32", 28, timer(lambda: hs.remain_step(p, g32, ema=True)))
        del hs
        if not args.skip_select:
            # K2b: whole select (init, hist0, scan, hist1, scan, apply) and its passes
            state, bins, scratch = hp._select_buffers()
            topk = torch.empty(n, dtype=torch.uint8, device=dev)
            k = n // 2
            emit("topk_select_total", n, "-", 13, timer(lambda: hp.topk_mask(g32, k, out=topk)),
                 note="product form: two-pass select, apply stage as ONE cooperative launch, the whole select replayed "
                      "from a CUDA graph; 13 B/elem is the three-read algorithmic figure, actual traffic ~9 B/elem")
            assert int(topk.sum()) == k
            hp.select_graphs = False
            emit("topk_select_total_eager_launches", n, "-", 13, timer(lambda: hp.topk_mask(g32, k, out=topk)),
                 note="the same kernels launched one by one (no graph)")
            hp.select_graphs = True
            hp.select_two_pass = False
            emit("topk_select_total_three_reads", n, "-", 13, timer(lambda: hp.topk_mask(g32, k, out=topk)),
                 note="hist pass 0 + pass 1 + streaming apply = 4+4+4+1 B/elem")
            hp.select_two_pass = True
            capi.select_init(state, bins, k)
            emit("topk_hist_pass0", n, "-", 4,
                 timer(lambda: capi.select_hist(g32, None, capi.KEY_ABS, 0, state, bins)))
            capi.select_init(state, bins, k)
            capi.select_hist(g32, None, capi.KEY_ABS, 0, state, bins)
            capi.select_scan(0, state, bins)
            emit("topk_hist_pass1", n, "-", 4,
                 timer(lambda: capi.select_hist(g32, None, capi.KEY_ABS, 1, state, bins, scratch)))
            bins.zero_()
            capi.select_hist(g32, None, capi.KEY_ABS, 1, state, bins, scratch)
            capi.select_scan(1, state, bins)
            emit("topk_apply", n, "-", 5,
                 timer(lambda: capi.select_apply(g32, None, capi.KEY_ABS, state, None, scratch, topk)))
            assert int(topk.sum()) == k
            # the two-pass form: pass 1 with the provisional mask, then the candidate-only apply
            capi.select_init(state, bins, k)
            capi.select_hist(g32, None, capi.KEY_ABS, 0, state, bins)
            capi.select_scan(0, state, bins)
            emit("topk_hist_pass1_with_mask", n, "-", 5,
                 timer(lambda: capi.select_hist(g32, None, capi.KEY_ABS, 1, state, bins, scratch, mask=topk)))
            bins.zero_()
            capi.select_hist(g32, None, capi.KEY_ABS, 1, state, bins, scratch, mask=topk)
            capi.select_scan(1, state, bins)
            emit("topk_apply_candidates_only", n, "-", 0,
                 timer(lambda: capi.select_apply(g32, None, capi.KEY_ABS, state, None, scratch, topk)))
            assert int(topk.sum()) == k
            del topk
        del hp, p, g32, g, acc
        torch.cuda.empty_cache()
    out.close()


if __name__ == "__main__":
    main()
