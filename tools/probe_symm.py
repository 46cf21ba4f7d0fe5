#!/usr/bin/env python
"""Probe of the multi-GPU plumbing on a B200 box (run under torchrun, >= 2 ranks):

  * torch symmetric memory: allocation, rendezvous, peer pointers, signal pads, multicast (NVLS) pointer;
  * the NCCL baselines the fused exchange kernels are measured against: all-reduce, reduce-scatter and
    all-gather of a DiT-XL/2-sized vector (fp32 and bf16), device-timed, max over ranks.

Writes one JSON object per rank to gpurun_out/probe_symm_rank{r}.json; never raises (every stage is guarded).
"""
from __future__ import annotations

import json
import os
import traceback

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N3 = 675_129_632


def timed(fn, iters=5, warm=2):
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
    t = torch.tensor([a.elapsed_time(b) / iters], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"rank": rank, "world": world, "torch": torch.__version__}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)

    try:
        out["p2p"] = [bool(torch.cuda.can_device_access_peer(local, j)) for j in range(torch.cuda.device_count())
                      if j != local]
    except Exception as e:
        out["p2p_error"] = repr(e)

    # ---- NCCL baselines --------------------------------------------------------------------------
    try:
        per = (N3 + world * 16 - 1) // (world * 16) * 16
        n_pad = per * world
        res = {}
        for dt, name in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
            g = torch.randn(n_pad, device=dev, dtype=torch.float32).to(dt)
            shard = g[rank * per:(rank + 1) * per]
            es = g.element_size()
            ms = timed(lambda: dist.all_reduce(g, op=dist.ReduceOp.AVG))
            res[f"all_reduce_{name}"] = {"ms": ms, "busbw_GBps": 2 * (world - 1) / world * n_pad * es / ms / 1e6}
            ms = timed(lambda: dist.reduce_scatter_tensor(shard, g, op=dist.ReduceOp.AVG))
            res[f"reduce_scatter_{name}"] = {"ms": ms, "busbw_GBps": (world - 1) / world * n_pad * es / ms / 1e6}
            ms = timed(lambda: dist.all_gather_into_tensor(g, shard))
            res[f"all_gather_{name}"] = {"ms": ms, "busbw_GBps": (world - 1) / world * n_pad * es / ms / 1e6}
            del g, shard
        out["nccl"] = res
    except Exception:
        out["nccl_error"] = traceback.format_exc()

    # ---- symmetric memory ------------------------------------------------------------------------
    try:
        import torch.distributed._symmetric_memory as symm
        out["symm_backend"] = str(symm.get_backend(dev)) if hasattr(symm, "get_backend") else None
        t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
        h = symm.rendezvous(t, dist.group.WORLD)
        out["symm"] = {
            "buffer_ptrs": [hex(p) for p in h.buffer_ptrs], "signal_pad_ptrs": [hex(p) for p in h.signal_pad_ptrs],
            "multicast_ptr": hex(h.multicast_ptr), "has_multicast_support":
                bool(type(h).has_multicast_support(torch._C._autograd.DeviceType.CUDA, local))
                if hasattr(type(h), "has_multicast_support") else None,
            "signal_pad_size": h.signal_pad_size, "buffer_size": h.buffer_size, "world_size": h.world_size,
        }
        # functional check of a peer read through torch (plumbing sanity): rank r fills with r+1
        t.fill_(float(rank + 1))
        h.barrier(channel=0)
        peer = h.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
        out["symm"]["peer_read_ok"] = bool((peer == float((rank + 1) % world + 1)).all())
        h.barrier(channel=0)
        # big allocation: can a DiT-XL/2 gradient live in symmetric memory?
        big = symm.empty(N3 + 1024, dtype=torch.float32, device=dev)
        hb = symm.rendezvous(big, dist.group.WORLD)
        out["symm"]["big_ok"] = True
        out["symm"]["big_multicast_ptr"] = hex(hb.multicast_ptr)
        # torch's own multimem all-reduce as a functional check of NVLS + a timing reference
        try:
            big.fill_(1.0)
            hb.barrier(channel=0)
            nbytes_ok = (N3 // (world * 16)) * world * 16
            view = big[:nbytes_ok]
            ms = timed(lambda: torch.ops.symm_mem.multimem_all_reduce_(view, "sum", dist.group.WORLD.group_name))
            out["symm"]["multimem_all_reduce_f32"] = {"ms": ms, "busbw_GBps": 2 * (world - 1) / world * nbytes_ok * 4 / ms / 1e6}
        except Exception:
            out["symm"]["multimem_all_reduce_error"] = traceback.format_exc()[-1500:]
    except Exception:
        out["symm_error"] = traceback.format_exc()[-3000:]

    with open(os.path.join(ROOT, "gpurun_out", f"probe_symm_rank{rank}.json"), "w") as f:
        json.dump(out, f, indent=1)
    if rank == 0:
        print(json.dumps(out, indent=1))
    try:
        dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
