#!/usr/bin/env python
"""bench.py — SFR-on hot-path throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA kernels via the C ABI)
  python bench.py --impl reference [...]                          the reference's CPU op sequences

Workload (config.workload): a DiT-XL/2-sized flat fp32 parameter vector per GPU
(N3 = 675,129,632 elements).  ONE step = one pass of the whole hot path over one
synthetic forget batch gradient and one remain batch gradient:

    K1   F_f += g_f**2 / L          12 B/elem         DiT/generate_fisher.py:236-239
    K1   F_r += g_r**2 / L          12 B/elem         DiT/generate_fisher.py:276-279
    K2a  mask = (F_f+e)/(F_r+e)>=th  9 B/elem         DiT/generate_mask.py:34-39
    norm sum((g_f*mask)**2)          5 B/elem         DiT/forget.py:293-298
    K3   AdamW(g_f*mask, clipped)   29 B/elem         DiT/forget.py:289-299
    K3   AdamW(g_r) + EMA           36 B/elem         DiT/forget.py:310-322
                                   ----
                                   103 B/elem algorithmic HBM traffic per step

`value` = algorithmic bytes of all ranks / device time (CUDA events, max over ranks), inputs
resident in HBM.  `e2e` = the same step through the public API (sfron_b200.HotPath) with the two
gradient vectors arriving from PINNED HOST memory and the step's scalars (clip norm, mask zero
count) read back, host<->device copies inside the timed region.
Multi-GPU: weak scaling — every rank owns one N3-element shard of an N x N3 vector; the kernels are
shard-local and the path's real exchange steps (clip-norm scalar and mask zero-count all-reduce over
NCCL) run inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N3 = 675_129_632                       # DiT-XL/2 parameter count
BYTES = {"fisher_forget": 12, "fisher_remain": 12, "ratio_mask": 9, "masked_sumsq": 5,
         "fused_update_forget": 29, "fused_update_remain_ema": 36}
BYTES_PER_ELEM = sum(BYTES.values())   # 103
FISHER_L = 2000.0                      # DiT/generate_fisher.py default n_iters
METRIC = "sfron_hot_path_GBps"         # SFR-on masked-update + Fisher algorithmic GB/s
UNIT = "GB/s"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------------------------- clocks
CLOCK_QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
               "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
               "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={CLOCK_QUERY}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock and the throttle reasons seen between the two wall-clock marks
        (the timed region); all samples if no mark is given."""
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                if t_begin is not None:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if not (t_begin - 0.05 <= ts <= t_end + 0.05):
                        continue
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------- reference arm
def reference_state(n, seed=1234):
    g = torch.Generator().manual_seed(seed)
    p = torch.nn.Parameter(torch.randn(n, generator=g) * 0.02)
    st = dict(p=p, g_f=torch.randn(n, generator=g) * 1e-2, g_r=torch.randn(n, generator=g) * 1e-2,
              acc_f=torch.zeros(n), acc_r=torch.zeros(n), ema=p.detach().clone(),
              opt=torch.optim.AdamW([p], lr=1e-4, weight_decay=0))          # DiT/forget.py:199
    return st


def reference_step(st):
    """The reference's op sequences for one step, in its own torch form, on the host cores
    (restated in oracle/sfron_oracle.py; flat-vector form of SURVEY.md §2.1)."""
    from oracle import sfron_oracle as O
    p, opt = st["p"], st["opt"]
    O.flat_fisher_accum(st["acc_f"], st["g_f"], FISHER_L)                    # F_f += grad.cpu()**2 / n_iters
    O.flat_fisher_accum(st["acc_r"], st["g_r"], FISHER_L)
    mask = O.flat_ratio_mask(st["acc_f"], st["acc_r"], 1.0)                  # generate_mask.py:34-35
    p.grad = st["g_f"]
    O.flat_masked_clip_(p.grad, mask, 1.0)                                   # grad *= mask ; clip_grad_norm_
    opt.step()
    p.grad = st["g_r"]
    opt.step()                                                               # remain step, no clip
    with torch.no_grad():
        O.ema_dit_({"w": st["ema"]}, {"w": p}, 0.9999)                       # update_ema
    return mask


def time_reference(n, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    st = reference_state(n)
    for _ in range(warmup):
        reference_step(st)
    t0 = time.perf_counter()
    for _ in range(steps):
        reference_step(st)
    dt = time.perf_counter() - t0
    return BYTES_PER_ELEM * n * steps / dt / 1e9, dt / steps * 1e3


def run_reference(args, rank, world):
    if rank != 0:
        return
    n = args.ref_elems
    gbs, ms = time_reference(n, args.steps, args.warmup)
    cores = torch.get_num_threads()
    sample = f"{n} of {N3} elements per step (same 103 B/elem op sequence), torch {torch.__version__} CPU"
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.elems, args.gpus),
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n, gpus):
    return {"workload": "DiT-XL/2-sized flat fp32 parameter vector, one SFR-on hot-path pass per step "
                        "(Fisher forget+remain, ratio mask, masked clip norm, masked AdamW forget step, "
                        "AdamW remain step + EMA)",
            "elements_per_gpu": n, "bytes_per_element": BYTES_PER_ELEM, "optimizer": "AdamW lr 1e-4 wd 0",
            "ema_decay": 0.9999, "grad_clip": 1.0, "threshold": 1.0, "parallelism": f"shard x{gpus}",
            "l2": "every stream (>= 0.67 GB per launch) exceeds the 126 MB L2: inputs larger than L2, no flush"}


# ------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    import sfron_b200 as sfr

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the SFR-on hot path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    sfr.capi.load()
    n = args.elems
    opt = sfr.OptConfig(kind="adamw", lr=1e-4, weight_decay=0.0)
    if world > 1:
        from sfron_b200.dist import ShardGroup, ShardedHotPath
        # weak scaling: the global vector has world x n elements, rank r owns [r*n, (r+1)*n)
        hp = ShardedHotPath(ShardGroup(n * world), dev, opt, ema_mode="dit", ema_a=0.9999)
        n = hp.n
    else:
        hp = sfr.HotPath(n, dev, opt, ema_mode="dit", ema_a=0.9999)

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    p = torch.randn(n, device=dev, generator=gen) * 0.02
    g_f = torch.randn(n, device=dev, generator=gen) * 1e-2
    g_r = torch.randn(n, device=dev, generator=gen) * 1e-2
    hp.init_slow(p)

    def step():
        hp.fisher_accumulate("forget", g_f, FISHER_L)
        hp.fisher_accumulate("remain", g_r, FISHER_L)
        hp.ratio_mask(1.0)
        hp.forget_step(p, g_f, max_norm=1.0)
        hp.remain_step(p, g_r, ema=True)

    labels = ["fisher_forget", "fisher_remain", "ratio_mask", "masked_sumsq", "fused_update_forget",
              "fused_update_remain_ema"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: K steps, CUDA events on the launching (current) stream -------------------------
    events = []

    def probe(_label):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        events.append(e)

    hp.trace = probe
    barrier()
    wall_begin = time.time()
    start = torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        probe("step_start")
        step()
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    barrier()
    wall_end = time.time()
    hp.trace = None
    clocks = sampler.stop(wall_begin, wall_end) if sampler else None
    total_ms = start.elapsed_time(end)
    per = len(labels) + 1
    assert len(events) == per * args.steps, (len(events), per, args.steps)
    kernel_ms = {lab: 0.0 for lab in labels}
    for s in range(args.steps):
        ev = events[s * per:(s + 1) * per]
        for i, lab in enumerate(labels):
            kernel_ms[lab] += ev[i].elapsed_time(ev[i + 1])
    kernel_ms = {k: v / args.steps for k, v in kernel_ms.items()}

    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = BYTES_PER_ELEM * n * world / (ms_per_step * 1e-3) / 1e9

    # ---- outside the step: the top-k select (K2b), which the DiT flow does not use (SalUn / DDPM generate_mask)
    extra = {}
    if not args.no_extra:
        topk = torch.empty(n, dtype=torch.uint8, device=dev)
        k_global = (n * world) // 2
        hp.topk_mask(g_f, k_global, out=topk)
        times = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            hp.topk_mask(g_f, k_global, out=topk)
            b.record()
            b.synchronize()
            times.append(a.elapsed_time(b))
        ms = sorted(times)[1]
        extra["topk_select_k_half"] = {"ms": round(ms, 4), "GBps": round(13 * n / (ms * 1e-3) / 1e9, 1),
                                       "bytes_per_elem": 13, "traffic_bytes_per_elem": 9,
                                       "selected": int(topk.sum(dtype=torch.int64))}
        del topk

    # ---- e2e: gradients from pinned host memory, scalars read back, through the public API ------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    host_g = torch.empty(n, dtype=torch.float32).pin_memory()
    blk_n = min(n, 1 << 24)   # host RNG over 675M elements is slow: draw 16 Mi values and tile them
    blk = torch.randn(blk_n, generator=torch.Generator().manual_seed(7 + rank)) * 1e-2
    for off in range(0, n, blk_n):
        m = min(blk_n, n - off)
        host_g[off:off + m].copy_(blk[:m])
    result_host = torch.empty(2, dtype=torch.float64).pin_memory()

    feeder = sfr.HostGradientFeeder(n, dev, slots=("forget", "remain"))

    def step_on(gf, gr):
        hp.fisher_accumulate("forget", gf, FISHER_L)
        hp.fisher_accumulate("remain", gr, FISHER_L)
        hp.ratio_mask(1.0)
        hp.forget_step(p, gf, max_norm=1.0)
        hp.remain_step(p, gr, ema=True)

    def e2e_step():
        g = feeder.acquire()                               # this step's gradients, copied from pinned host memory
        feeder.submit(forget=host_g, remain=host_g)        # next step's H2D overlaps this step's kernels
        step_on(g["forget"], g["remain"])
        feeder.release()
        res = torch.stack([hp.sumsq[0], hp.zero_count[0].double()])
        result_host.copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the step's result is on the host
        return result_host

    # Pipelined: every step submits the NEXT step's host->device copy before running its kernels, so the
    # timed region holds exactly e2e_steps submissions (2 x 4n bytes each) and e2e_steps kernel passes; the
    # copy that is still in flight at the end is waited for inside the timed region.
    feeder.submit(forget=host_g, remain=host_g)
    e2e_step()                                             # warm-up step
    barrier()                                              # the prefetched copy has landed: start from a full pipe
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_dt = time.perf_counter() - t0
    t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = BYTES_PER_ELEM * n * world * e2e_steps / float(t.item()) / 1e9

    if rank != 0:
        return
    peak, peak_kind = measured_peak()
    dom = "fused_update_remain_ema"
    achieved = BYTES[dom] * n / (kernel_ms[dom] * 1e-3) / 1e9
    kernels = {k: {"ms": round(v, 4), "GBps": round(BYTES[k] * n / (v * 1e-3) / 1e9, 1),
                   "frac": round(BYTES[k] * n / (v * 1e-3) / 1e9 / peak, 4), "bytes_per_elem": BYTES[k]}
               for k, v in kernel_ms.items()}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        gbs, ms = time_reference(args.ref_elems, args.cpu_steps, 2)
        cpu = {"value": gbs, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{args.ref_elems} of {N3} elements per step x {args.cpu_steps} steps, reference-form "
                         f"torch CPU ops (oracle/sfron_oracle.py), {ms:.0f} ms/step"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n, world),
        "roofline": {"bound": "hbm", "kernel": "fused_update_kernel<AdamW,EMA_DIT,f32> (remain step + EMA)",
                     "achieved": achieved, "peak": peak, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})"
                     if peak_kind == "measured" else "fallback 6.65 TB/s", "unit": "GB/s",
                     "frac": achieved / peak, "traffic": args.traffic},
        "hot_path_frac_of_peak": value / world / peak,
        "kernels": kernels,
        "extra_kernels": {k: dict(v, frac=round(v["GBps"] / peak, 4)) for k, v in extra.items()},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * 4 * n, "d2h_bytes_per_step": 16,
                "steps": e2e_steps},
        "gpu_launches": len(labels) * args.steps,
        "steps_per_s": 1e3 / ms_per_step,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elems", type=int, default=N3, help="elements per GPU (default: DiT-XL/2, 675,129,632)")
    ap.add_argument("--ref-elems", type=int, default=1 << 26, help="bounded CPU sample per step")
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the K2b select timing outside the step")
    ap.add_argument("--traffic", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel (default: the committed ncu capture "
                         "profiles/r1_ncu_full_step_kernels_n675M_raw.csv when --elems is the default)")
    args = ap.parse_args()
    if args.traffic is None and args.elems == N3:
        # ncu --set full (profiles/r1_ncu_full_step_kernels_n675M_raw.csv), fused_update_kernel<AdamW,EMA_DIT,f32>
        # at n = N3: dram__bytes_read.sum 13.502611 GB + dram__bytes_write.sum 10.745242 GB per launch
        # (algorithmic: 36 B x N3 = 24.305 GB)
        args.traffic = 13.502611e9 + 10.745242e9
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                                  # timing rule: W >= 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on STDOUT at NCCL_DEBUG=VERSION; stdout carries exactly one JSON line
        os.environ["NCCL_DEBUG"] = os.environ.get("SFR_NCCL_DEBUG", "WARN")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
