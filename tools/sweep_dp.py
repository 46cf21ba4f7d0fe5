#!/usr/bin/env python
"""BASELINE config 5 on N GPUs: the data-parallel hot-path step of bench.py (two gradient reduces fused with K1, ratio
mask, clip norm, two optimizer steps each followed by a weight push) swept over flat vectors of 1e7 .. 2e9 elements, fp32
and bf16 exchange, at the N ranks of this torchrun job.  One JSON line per (n, dtype):

    ms_per_step (CUDA events, max over ranks), algorithmic GB/s of the step (103 B/elem fp32, 97 bf16), the NVLink bytes
    each GPU sends per step and the resulting link rate, the transport.

    python -m torch.distributed.run --nproc-per-node N ... tools/sweep_dp.py --out profiles/r2_sweep_dp_nN.jsonl
With one rank it runs the single-GPU step (no exchange) at the same sizes, for the N = 1 column.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FISHER_L = 2000.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="10000000,38632323,100000000,675129632,859520964,2000000000")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--transport", default="auto")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    import sfron_b200 as sfr
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    opt = sfr.OptConfig(kind="adamw", lr=1e-4, weight_decay=0.0)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, steps):
        for _ in range(3):
            step()
        sync()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            step()
        b.record()
        sync()
        t = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    lines = []
    for n in [int(x) for x in args.sizes.split(",")]:
        for g_dtype, name, bpe in ((torch.float32, "f32", 103), (torch.bfloat16, "bf16", 97)):
            gen = torch.Generator(device=dev).manual_seed(5 + rank)
            if world == 1:
                hp = sfr.HotPath(n, dev, opt, ema_mode="dit", ema_a=0.9999)
                p = torch.randn(n, device=dev, generator=gen) * 0.02
                gf = (torch.randn(n, device=dev, generator=gen) * 1e-2).to(g_dtype)
                gr = (torch.randn(n, device=dev, generator=gen) * 1e-2).to(g_dtype)
                w16 = torch.empty(n, dtype=torch.bfloat16, device=dev) if g_dtype == torch.bfloat16 else None
                hp.init_slow(p)

                def step():
                    hp.fisher_accumulate("forget", gf, FISHER_L)
                    hp.fisher_accumulate("remain", gr, FISHER_L)
                    hp.ratio_mask(1.0)
                    hp.forget_step(p, gf, max_norm=1.0, p_bf16=w16)
                    hp.remain_step(p, gr, ema=True, p_bf16=w16)

                transport, out_b = "single GPU", 0.0
            else:
                from sfron_b200.dist import PeerExchange, ShardGroup, ShardedHotPath
                n_pad = -(-n // (16 * world)) * (16 * world)
                sg = ShardGroup(n, padded_len=n_pad)
                xchg = PeerExchange(sg, dev, transport=args.transport)
                hp = ShardedHotPath(sg, dev, opt, ema_mode="dit", ema_a=0.9999)
                hp.attach_exchange(xchg)
                gf, gr = xchg.alloc(n_pad, g_dtype), xchg.alloc(n_pad, g_dtype)
                for b in (gf, gr):
                    b.tensor.copy_((torch.randn(n_pad, device=dev, generator=gen) * 1e-2).to(g_dtype))
                w0 = torch.randn(n_pad, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 0.02
                if g_dtype == torch.float32:
                    w, w16 = xchg.alloc(n_pad, torch.float32), None
                    w.tensor.copy_(w0)
                    p = w.tensor[sg.lo:sg.hi]
                else:
                    w, w16 = None, xchg.alloc(n_pad, torch.bfloat16)
                    w16.tensor.copy_(w0)
                    p = w0[sg.lo:sg.hi].clone()
                del w0
                hp.init_slow(p)

                def step():
                    hp.dp_fisher_accumulate("forget", gf, FISHER_L, keep="forget")
                    hp.dp_fisher_accumulate("remain", gr, FISHER_L, keep="remain")
                    hp.ratio_mask(1.0)
                    hp.dp_forget_step(p, hp.reduced("forget"), weights=w, weights_bf16=w16, max_norm=1.0)
                    hp.dp_remain_step(p, hp.reduced("remain"), weights=w, weights_bf16=w16, ema=True)

                transport = xchg.transport_name
                es, ws = gf.tensor.element_size(), 4 if w is not None else 2
                red_t, push_t = (transport.split("+") * 2)[:2]
                frac = (world - 1) / world
                out_b = 2 * (n_pad * es if red_t == "multimem" else frac * n_pad * es) + \
                    2 * (n_pad * ws / world if push_t == "multimem" else frac * n_pad * ws)
                in_b = 2 * (n_pad * es / world if red_t == "multimem" else frac * n_pad * es) + \
                    2 * (n_pad * ws if push_t == "multimem" else frac * n_pad * ws)
                out_b = max(out_b, in_b)
            ms = timed(step, args.steps)
            if world > 1:
                torch.cuda.synchronize()
                xchg.check()
            rec = dict(n=n, n_gpus=world, exchange_dtype=name, transport=transport, ms_per_step=round(ms, 4),
                       bytes_per_elem=bpe, GBps=round(bpe * n / ms / 1e6, 1),
                       nvlink_GB_per_gpu_per_direction=round(out_b / 1e9, 3),
                       link_GBps=round(out_b / ms / 1e6, 1) if out_b else None)
            lines.append(rec)
            if rank == 0:
                print(json.dumps(rec), flush=True)
            del hp, p, gf, gr, step
            if world > 1:
                del xchg, w, w16
            torch.cuda.empty_cache()
    if rank == 0 and args.out:
        with open(args.out, "a") as f:
            for rec in lines:
                f.write(json.dumps(rec) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
