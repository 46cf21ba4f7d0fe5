#!/usr/bin/env python
"""End-to-end DiT-XL/2 unlearning steps/s on B200 (BASELINE.json: "DiT-XL/2 unlearn steps/s").

Model forward/backward in PyTorch (tools/dit_xl2.py harness, TF32 matmuls as DiT/forget.py:6-8),
synthetic 4x32x32 latents, random-init weights.  Two arms on the SAME GPU:

  stock : the reference's loop as written — DiT/forget.py:285-322 for the forget loop
          (per-parameter `grad *= mask[name].to(device)`, clip_grad_norm_, AdamW.step, update_ema)
          and DiT/generate_fisher.py:232-239 for the Fisher loop (`F[name] += grad.cpu()**2 / n`).
          `--mask-on-device` is a best-effort variant with the mask uploaded once.
  ours  : the same iterations with everything after backward done by the flat-vector kernels.

Single GPU: `python tools/dit_e2e.py`.  N GPUs (ours only; the reference's DataParallel is
single-process): `python -m torch.distributed.run --nproc-per-node N ... tools/dit_e2e.py --arm ours`
— data parallel, gradient all-reduce over NCCL, sharded update, weight all-gather.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from copy import deepcopy

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

from dit_xl2 import DiTXL2Harness, synthetic_loss  # noqa: E402


def make_batch(bs, dev, gen, forget_class=207, dtype=torch.float32):
    x = torch.randn(bs, 4, 32, 32, device=dev, generator=gen).to(dtype)
    t = torch.randint(0, 1000, (bs,), device=dev, generator=gen)
    noise = torch.randn(bs, 4, 32, 32, device=dev, generator=gen).to(dtype)
    y_f = torch.full((bs,), forget_class, device=dev)
    y_r = torch.randint(0, 1000, (bs,), device=dev, generator=gen)
    return x, t, noise, y_f, y_r


def timed(fn, steps, warmup, sync):
    for i in range(warmup):
        fn(i)
    sync()
    t0 = time.perf_counter()
    for i in range(steps):
        fn(warmup + i)
    sync()
    return (time.perf_counter() - t0) / steps


def run_stock(args, dev, ac):
    torch.manual_seed(0)
    model = DiTXL2Harness().to(dev)
    ema = deepcopy(model)
    for p in ema.parameters():
        p.requires_grad = False
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0)
    gen = torch.Generator(device=dev).manual_seed(1)
    cg = torch.Generator().manual_seed(2)
    mask = {n: (torch.rand(p.shape, generator=cg) < 0.5) if p.requires_grad else 0
            for n, p in model.named_parameters()}              # what torch.load(mask_path) returns: CPU bools
    if args.mask_on_device:
        mask = {n: (m.to(dev) if torch.is_tensor(m) else m) for n, m in mask.items()}
    model.train()
    res = {}

    def forget_iter(i):
        x, t, noise, y_f, y_r = make_batch(args.batch_size, dev, gen)
        loss = -synthetic_loss(model, x, t, y_f, noise, ac)
        opt.zero_grad()
        (args.forget_alpha * loss).backward()
        for name, param in model.named_parameters():
            if param.grad is not None:
                param.grad *= mask[name].to(param.grad.device)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        loss = synthetic_loss(model, x, t, y_r, noise, ac)
        opt.zero_grad()
        loss.backward()
        opt.step()
        with torch.no_grad():
            for (_, pe), (_, pm) in zip(ema.named_parameters(), model.named_parameters()):
                pe.mul_(0.9999).add_(pm.data, alpha=1 - 0.9999)

    res["forget_s_per_it"] = timed(forget_iter, args.steps, args.warmup, torch.cuda.synchronize)

    fisher = {n: 0 for n, _ in model.named_parameters()}

    def fisher_iter(i):
        x, t, noise, y_f, _ = make_batch(args.batch_size, dev, gen)
        loss = synthetic_loss(model, x, t, y_f, noise, ac)
        opt.zero_grad()
        loss.backward()
        with torch.no_grad():
            for name, param in model.named_parameters():
                if param.grad is not None:
                    fisher[name] += (param.grad.data.cpu() ** 2) / 2000

    res["fisher_s_per_it"] = timed(fisher_iter, max(2, args.steps // 4), 1, torch.cuda.synchronize)
    return res


def run_ours(args, dev, ac, rank, world):
    import sfron_b200 as sfr
    torch.manual_seed(0)
    model = DiTXL2Harness().to(dev)
    if args.dtype == "bf16":
        model = model.to(torch.bfloat16)      # bf16 working weights + grads; fp32 master / state in the kernels
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    opt = sfr.OptConfig(kind="adamw", lr=1e-4, weight_decay=0.0)
    peer = world > 1 and args.dp_exchange.startswith("peer")
    overlap = peer and args.dp_exchange.startswith("peer-overlap")
    xchg, sym, ov = None, {}, None
    if peer:
        # gradients and the weights the model reads live in symmetric (peer-mapped) memory: the data-parallel
        # kernels pull gradient shards out of every rank's buffer and push updated weights into every rank's
        import torch.distributed as dist
        from sfron_b200.dist import PeerExchange, ShardGroup, ShardedHotPath
        n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
        parts = args.overlap_parts if overlap else 1
        n_pad = -(-n_train // (16 * world * parts)) * (16 * world * parts)
        sg = ShardGroup(n_train, padded_len=n_pad, parts=parts)
        xchg = PeerExchange(sg, dev, transport=args.dp_exchange.partition(":")[2] or "auto")

        def alloc(role, numel, dtype):
            if role == "p" and args.dtype == "bf16":          # fp32 master: only this rank's shard is ever used
                return torch.zeros(numel, dtype=dtype, device=dev)
            sym[role] = xchg.alloc(numel, dtype)
            return sym[role].tensor

        flat = sfr.FlatParams(model, dev, pad_multiple=16 * world * parts, alloc=alloc)
        assert flat.n == n_train and flat.n_padded == n_pad
        hp = ShardedHotPath(sg, dev, opt, ema_mode="dit", ema_a=0.9999)
        hp.attach_exchange(xchg)
        lo, hi = (sg.lo, sg.hi) if not overlap else (0, sg.n_local)
        if overlap:
            # the late half of the vector is exchanged on a side stream, by a few CTAs, while backward still runs
            from sfron_b200.dist import OverlappedBackward
            ov = OverlappedBackward(flat, sg, xchg.sibling(), max_ctas=args.overlap_ctas)
    else:
        flat = sfr.FlatParams(model, dev, pad_multiple=16 * world)
    if peer:
        pass
    elif world > 1:
        import torch.distributed as dist
        from sfron_b200.dist import ShardGroup, ShardedHotPath
        sg = ShardGroup(flat.n, padded_len=flat.n_padded)
        hp = ShardedHotPath(sg, dev, opt, ema_mode="dit", ema_a=0.9999)
        lo, hi = sg.lo, sg.hi
    else:
        sg = None
        hp = sfr.HotPath(flat.n, dev, opt, ema_mode="dit", ema_a=0.9999)
        lo, hi = 0, flat.n
    if overlap:
        # two spans per rank: the fp32 master shard is their concatenation (bf16 model) or two views of the weights
        p_loc = sg.local(flat.p) if flat.p_work is not None else [flat.p[g:g + c] for g, _, c in sg.spans]
        w_loc = None
    else:
        p_loc = flat.p[lo:hi]
        w_loc = None if flat.p_work is None else flat.p_work[lo:hi]
    weights_full = flat.p_padded if flat.p_work is None else flat.p_work_padded   # what the model reads
    hp.init_slow(sg.local(flat.p) if overlap else p_loc)
    frozen_slow = flat.frozen.clone()
    hp.mask.copy_((torch.rand(hi - lo, device=dev, generator=gen) < 0.5).to(torch.uint8))
    hp.mark_mask_ready()
    model.train()
    res = {}

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    reducer = None
    if sg is not None and args.dp_exchange == "bucketed":
        from sfron_b200.dist import BucketedGradReducer
        assert not args.cuda_graph, "bucketed exchange is an eager-mode option"
        reducer = BucketedGradReducer(flat, sg, bucket_bytes=args.bucket_mb << 20)

    def grads():
        if sg is not None:
            if reducer is not None:
                return reducer.finish()                      # buckets were all-reduced while backward was running
            if args.dp_exchange == "allreduce":
                return sg.reduce_gradients_(flat.g_padded, average=True)[:sg.n_local]
            return sg.reduce_scatter_gradients_(flat.g_padded, average=True)
        return flat.g

    dt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    ac = ac.to(dt)

    def forget_body(x, t, noise, y_f, y_r):
        if overlap:
            side = ov.side
            main = torch.cuda.current_stream()
            kw = dict(weights=sym.get("p"), weights_bf16=sym.get("p_work"))
            loss = args.forget_alpha * -synthetic_loss(model, x, t, y_f, noise, ac)
            hp.dp_begin_step(ov, p_loc, sym["g"], mask=hp.require_mask(), max_norm=1.0, **kw)
            loss.backward()                              # late half: reduce + norm on the side stream meanwhile
            hp.dp_finish_step()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                flat.g.zero_()                           # the gradient memset runs beside the remain forward
            loss = synthetic_loss(model, x, t, y_r, noise, ac)
            main.wait_stream(side)
            hp.dp_begin_step(ov, p_loc, sym["g"], ema=True, **kw)
            loss.backward()                              # late half: reduce + K3 + EMA + push on the side stream
            hp.dp_finish_step()
            hp.ema_only(flat.frozen, frozen_slow)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                flat.g.zero_()
            main.wait_stream(side)
            return
        (args.forget_alpha * -synthetic_loss(model, x, t, y_f, noise, ac)).backward()
        if peer:
            # barrier -> reduce + masked norm (one kernel over NVLink) -> barrier carrying the norm -> K3 + weight
            # push (one kernel) -> barrier
            hp.dp_forget_step(p_loc, sym["g"], weights=sym.get("p"), weights_bf16=sym.get("p_work"), max_norm=1.0)
            flat.g.zero_()
            synthetic_loss(model, x, t, y_r, noise, ac).backward()
            # barrier -> reduce + K3 + EMA + weight push in ONE kernel -> barrier
            hp.dp_remain_step(p_loc, sym["g"], weights=sym.get("p"), weights_bf16=sym.get("p_work"), ema=True)
            hp.ema_only(flat.frozen, frozen_slow)
            flat.g.zero_()
            return
        # single GPU: the update kernel zeroes g on its way out (fused optimizer.zero_grad());
        # sharded: each rank only rewrites its slice, so the full local gradient is memset
        hp.forget_step(p_loc, grads(), max_norm=1.0, zero_grad=sg is None, p_bf16=w_loc)
        if sg is not None:
            sg.all_gather_params_(weights_full)
            flat.g.zero_()
        synthetic_loss(model, x, t, y_r, noise, ac).backward()
        hp.remain_step(p_loc, grads(), ema=True, zero_grad=sg is None, p_bf16=w_loc)
        hp.ema_only(flat.frozen, frozen_slow)
        if sg is not None:
            sg.all_gather_params_(weights_full)
            flat.g.zero_()

    def forget_iter(i):
        forget_body(*make_batch(args.batch_size, dev, gen, dtype=dt))

    flat.g.zero_()
    if args.cuda_graph:
        # Whole iteration (2 forward/backward passes in PyTorch + the hot-path kernels) captured ONCE and
        # replayed: at batch 1 the forward/backward is launch-bound.  Inputs live in static buffers that
        # are refilled before each replay; the optimizer step count lives on the device (replay-safe).
        hp.enable_graph_replay()
        static = list(make_batch(args.batch_size, dev, gen, dtype=dt))

        def body():
            forget_body(*static)

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                body()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            body()

        def graph_iter(i):
            for dst, src in zip(static, make_batch(args.batch_size, dev, gen, dtype=dt)):
                dst.copy_(src)
            graph.replay()

        res["forget_s_per_it"] = timed(graph_iter, args.steps, args.warmup, sync)
        res["graph_steps_on_device"] = int(hp.step_dev)
    else:
        res["forget_s_per_it"] = timed(forget_iter, args.steps, args.warmup, sync)

    def fisher_step():
        if peer:
            hp.dp_fisher_accumulate("forget", sym["g"], 2000.0)     # barrier -> reduce + K1 in one kernel -> barrier
        else:
            hp.fisher_accumulate("forget", grads(), 2000.0)
        flat.g.zero_()

    def fisher_iter(i):
        x, t, noise, y_f, _ = make_batch(args.batch_size, dev, gen, dtype=dt)
        synthetic_loss(model, x, t, y_f, noise, ac).backward()
        fisher_step()

    if args.cuda_graph:
        static_f = list(make_batch(args.batch_size, dev, gen, dtype=dt))

        def fisher_body():
            x, t, noise, y_f, _ = static_f
            synthetic_loss(model, x, t, y_f, noise, ac).backward()
            fisher_step()

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fisher_body()
        torch.cuda.current_stream().wait_stream(side)
        fgraph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(fgraph):
            fisher_body()

        def fisher_graph_iter(i):
            for dst, src in zip(static_f, make_batch(args.batch_size, dev, gen, dtype=dt)):
                dst.copy_(src)
            fgraph.replay()

        res["fisher_s_per_it"] = timed(fisher_graph_iter, args.steps, args.warmup, sync)
    else:
        res["fisher_s_per_it"] = timed(fisher_iter, args.steps, args.warmup, sync)
    if xchg is not None:
        torch.cuda.synchronize()
        xchg.check()
        res["transport"] = xchg.transport_name
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", default="both", choices=["both", "stock", "ours"])
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch-size", type=int, default=1, help="per GPU (DiT/forget.py default 1)")
    ap.add_argument("--forget-alpha", type=float, default=1e-3)
    ap.add_argument("--mask-on-device", action="store_true")
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"],
                    help="ours arm only: bf16 working weights / gradients with fp32 master + state (BASELINE config 3)")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="ours arm: capture the whole forget iteration (collectives included) in a CUDA graph and replay it")
    ap.add_argument("--dp-exchange", default="peer",
                    help="peer[:transport] (default): the library's fused exchange kernels over NVLink peer memory "
                         "(transport auto | tma | p2p | multimem | <reduce>+<push>); peer-overlap[:transport]: the same, with "
                         "the late half of the vector exchanged beside the backward pass; reduce_scatter | allreduce: NCCL collectives "
                         "around the shard-local kernels; bucketed: NCCL all-reduce in buckets started from autograd "
                         "hooks while backward still runs (sfron_b200.dist.BucketedGradReducer)")
    ap.add_argument("--bucket-mb", type=int, default=64)
    ap.add_argument("--overlap-parts", type=int, default=8,
                    help="peer-overlap: pieces the flat vector is cut into; all but the first are exchanged during backward")
    ap.add_argument("--overlap-ctas", type=int, default=32,
                    help="peer-overlap[:transport]: CTAs of the exchange kernels that run beside the backward pass")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cuda.matmul.allow_tf32 = True       # DiT/forget.py:6-8
    torch.backends.cudnn.allow_tf32 = True
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    betas = torch.linspace(1e-4, 0.02, 1000, dtype=torch.float64)
    ac = torch.cumprod(1 - betas, 0).float().to(dev)
    out = {"model": "DiT-XL/2 harness, 675,129,632 params, random init", "batch_size_per_gpu": args.batch_size,
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "tf32": True}
    if args.arm in ("both", "stock") and world == 1:
        r = run_stock(args, dev, ac)
        out["stock" + ("_mask_on_device" if args.mask_on_device else "")] = {
            "forget_steps_per_s": 1 / r["forget_s_per_it"], "fisher_steps_per_s": 1 / r["fisher_s_per_it"]}
        torch.cuda.empty_cache()
    if args.arm in ("both", "ours"):
        r = run_ours(args, dev, ac, rank, world)
        out["ours" + ("_bf16" if args.dtype == "bf16" else "") + ("_cudagraph" if args.cuda_graph else "")] = {"forget_steps_per_s": 1 / r["forget_s_per_it"], "fisher_steps_per_s": 1 / r["fisher_s_per_it"],
                       "global_batch": args.batch_size * world,
                       "dp_exchange": (args.dp_exchange if world > 1 else None),
                       "transport": r.get("transport")}
    if rank == 0:
        line = json.dumps(out)
        print(line, flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(line + "\n")
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
