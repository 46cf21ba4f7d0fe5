#!/usr/bin/env python
"""Device-timed comparison of the hot path's cross-GPU exchange on a DiT-XL/2-sized vector (run under torchrun):

  NCCL baseline      reduce_scatter_tensor / all_gather_into_tensor / all_reduce           (torch.distributed)
  fused (peer.cu)    sfr_peer_reduce (+K1), sfr_peer_fused_update (local | peers -> K3 -> push), sfr_peer_barrier
                     over P2P pointers and over the NVLS multicast address, fp32 and bf16 streams

Every line: ms (CUDA events on the launching stream, max over ranks), the bytes each GPU must send / receive over
NVLink for that op, and the resulting per-direction link rate against the 770 GB/s peer-copy figure
(B200_PROFILING.md).  Output: JSON lines on stdout (rank 0) and --out.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N3 = 675_129_632
LINK_GBPS = 770.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--elems", type=int, default=N3)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import sfron_b200 as sfr
    from sfron_b200.dist import PeerExchange, ShardGroup, ShardedHotPath

    n = args.elems
    n_pad = (n + 16 * world - 1) // (16 * world) * (16 * world)
    sg = ShardGroup(n, padded_len=n_pad)
    xchg = PeerExchange(sg, dev)
    hp = ShardedHotPath(sg, dev, sfr.OptConfig(kind="adamw", lr=1e-4), ema_mode="dit", ema_a=0.9999)
    hp.attach_exchange(xchg)
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    lines = []

    def timed(fn, iters=args.iters, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / iters], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def emit(op, transport, dtype, ms, out_bytes, in_bytes, **kw):
        rec = dict(op=op, transport=transport, dtype=dtype, world=world, n=n, ms=round(ms, 4),
                   nvlink_out_GB=round(out_bytes / 1e9, 4), nvlink_in_GB=round(in_bytes / 1e9, 4),
                   link_GBps=round(max(out_bytes, in_bytes) / ms / 1e6, 1),
                   frac_of_770=round(max(out_bytes, in_bytes) / ms / 1e6 / LINK_GBPS, 4), **kw)
        lines.append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)

    frac = (world - 1) / world
    w32 = xchg.alloc(n_pad, torch.float32)
    w16 = xchg.alloc(n_pad, torch.bfloat16)
    w32.tensor.normal_(0, 0.02, generator=gen)
    p = w32.tensor[sg.lo:sg.hi]
    hp.init_slow(p)
    hp.mask.copy_((torch.rand(sg.n_local, device=dev, generator=gen) < 0.5).to(torch.uint8))
    hp.mark_mask_ready()
    hp.enable_graph_replay()
    transports = ["p2p", "tma"] + (["multimem", "p2p+multimem"] if w32.has_multicast else [])

    ms = timed(lambda: xchg.barrier(), iters=50)
    emit("barrier", "flags", "-", ms, 0, 0)
    ms = timed(lambda: xchg.barrier(hp.sumsq, hp.sumsq), iters=50)
    emit("barrier+sum(1 double)", "flags", "-", ms, 0, 0)
    one = torch.zeros(1, dtype=torch.float64, device=dev)
    ms = timed(lambda: dist.all_reduce(one), iters=50)
    emit("all_reduce(1 double)", "nccl", "-", ms, 0, 0)

    for dt, name, es in ((torch.float32, "f32", 4), (torch.bfloat16, "bf16", 2)):
        g = xchg.alloc(n_pad, dt)
        g.tensor.copy_(torch.empty(n_pad, device=dev).normal_(0, 1e-2, generator=gen))
        xchg.barrier()
        nb = frac * n_pad * es
        # ---- NCCL baselines (in place on the symmetric buffers: NCCL sees ordinary device memory)
        shard = g.tensor[rank * sg.per:(rank + 1) * sg.per]
        ms = timed(lambda: dist.reduce_scatter_tensor(shard, g.tensor, op=dist.ReduceOp.AVG))
        emit("reduce_scatter", "nccl", name, ms, nb, nb)
        ms = timed(lambda: dist.all_gather_into_tensor(g.tensor, shard))
        emit("all_gather", "nccl", name, ms, nb, nb)
        ms = timed(lambda: dist.all_reduce(g.tensor, op=dist.ReduceOp.AVG))
        emit("all_reduce", "nccl", name, ms, 2 * nb, 2 * nb)
        g.tensor.copy_(torch.empty(n_pad, device=dev).normal_(0, 1e-2, generator=gen))
        xchg.barrier()
        torch.cuda.synchronize()
        for tname in transports:
            xchg._want = tname
            red_mc, push_mc = [t == "multimem" for t in (tname.split("+") * 2)[:2]]
            # bytes over NVLink per GPU: multimem -> the whole vector leaves (reduce) / enters (push) each GPU once,
            # own shard included; P2P -> the other ranks' shards, both directions
            r_out, r_in = (n_pad * es, n_pad / world * es) if red_mc else (nb, nb)

            def push_bytes(esz):
                return (n_pad / world * esz, n_pad * esz) if push_mc else (frac * n_pad * esz,) * 2

            red = hp.reduced("bench")
            if tname != "p2p+multimem":
                ms = timed(lambda: hp.dp_fisher_accumulate("forget", g, 2000.0, keep="bench"))
                emit("reduce+K1+keep (2 barriers)", tname, name, ms, r_out, r_in,
                     hbm_local_GB=round(sg.n_local * 12 / 1e9, 3))
            if tname != "multimem" or name == "f32":            # the push does not depend on the gradient dtype
                for esz, wname, kw in ((4, "f32", dict(weights=w32)), (2, "bf16", dict(weights_bf16=w16))):
                    po, pi = push_bytes(esz)
                    ms = timed(lambda: hp.dp_step(p, red, mask=hp.mask, max_norm=1.0, **kw))
                    emit(f"norm+K3(local g)+push {wname} (2 barriers)", tname, name, ms, po, pi)
            for esz, wname, kw in ((4, "f32", dict(weights=w32)), (2, "bf16", dict(weights_bf16=w16))):
                po, pi = push_bytes(esz)
                ms = timed(lambda: hp.dp_step(p, g, ema=True, **kw))
                emit(f"reduce+K3+EMA+push {wname}, ONE kernel (2 barriers)", tname, name, ms, r_out + po, r_in + pi)
            torch.cuda.synchronize()
            xchg.check()
        del g
    if rank == 0 and args.out:
        with open(args.out, "a") as f:
            for rec in lines:
                f.write(json.dumps(rec) + "\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
