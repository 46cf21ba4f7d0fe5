#!/usr/bin/env python
"""bench.py — SFR-on hot-path throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA kernels via the C ABI)
  python bench.py --impl reference [...]                          the reference's CPU op sequences

Workload (config.workload): ONE DiT-XL/2-sized flat fp32 parameter vector (N3 = 675,129,632 elements;
BASELINE.md §3).  ONE step = one pass of the whole hot path over one synthetic forget batch gradient and one
remain batch gradient:

    K1   F_f += g_f**2 / L          12 B/elem         DiT/generate_fisher.py:236-239
    K1   F_r += g_r**2 / L          12 B/elem         DiT/generate_fisher.py:276-279
    K2a  mask = (F_f+e)/(F_r+e)>=th  9 B/elem         DiT/generate_mask.py:34-39
    norm sum((g_f*mask)**2)          5 B/elem         DiT/forget.py:293-298
    K3   AdamW(g_f*mask, clipped)   29 B/elem         DiT/forget.py:289-299
    K3   AdamW(g_r) + EMA           36 B/elem         DiT/forget.py:310-322
                                   ----
                                   103 B/elem algorithmic HBM traffic per step

`value` = algorithmic bytes of the step (103 x N3) / device time (CUDA events, max over ranks), inputs
resident in HBM.  `e2e` = the same step through the public API with every rank's two gradient vectors
arriving from PINNED HOST memory and the step's scalars (clip norm, mask zero count) read back, host<->device
copies inside the timed region.

N = 1: the whole vector on one GPU.
N > 1: STRONG scaling on the SAME vector, data parallel as the reference's DataParallel loops
(DiT/forget.py:193,285-322; DiT/generate_fisher.py:173): every rank holds a FULL gradient of its own batch
(as after its backward pass) and a full copy of the weights; the vector's state is sharded.  Inside the timed
region, per step: the two gradients are reduced (mean over ranks) shard-wise — fused with K1 — the mask, clip
norm (summed across ranks) and both optimizer steps run on the shard, and after EACH optimizer step the updated
shard is pushed into every rank's weight vector (the next forward needs it).  Transport: the library's own
kernels over NVLink peer memory (csrc/peer.cu: P2P loads/stores or NVLS multimem) — `--transport nccl` runs the
same step with NCCL reduce-scatter / all-gather around shard-local kernels, and that line is reported beside
the main one (`extra.nccl_transport`), as are the r1 weak-scaling shard-local line (`extra.weak_shard_local`)
and the bf16-exchange variant of BASELINE config 3 (`extra.bf16_exchange`).
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
    cfg = workload_config(args.elems, args.gpus)
    # the CPU arm times a bounded sample of the vector (same op sequence, bandwidth-bound, so GB/s carries over);
    # it does NOT include the per-tensor device->host copies the real reference pays before these ops
    cfg["elements_timed"] = n
    cfg["note"] = "CPU sample of the workload; the reference's own D2H gradient copies are not charged to it"
    gbs, ms = time_reference(n, args.steps, args.warmup)
    cores = torch.get_num_threads()
    sample = f"{n} of {N3} elements per step (same 103 B/elem op sequence), torch {torch.__version__} CPU"
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n, gpus):
    return {"workload": "DiT-XL/2-sized flat fp32 parameter vector, one SFR-on hot-path pass per step "
                        "(Fisher forget+remain, ratio mask, masked clip norm, masked AdamW forget step, "
                        "AdamW remain step + EMA)",
            "elements": n, "elements_per_gpu": n if gpus == 1 else -(-n // gpus), "bytes_per_element": BYTES_PER_ELEM,
            "optimizer": "AdamW lr 1e-4 wd 0", "ema_decay": 0.9999, "grad_clip": 1.0, "threshold": 1.0,
            "parallelism": "single GPU" if gpus == 1 else
            f"data parallel x{gpus}: per-rank full gradients reduced shard-wise, sharded state, weights pushed to all ranks",
            "l2": "every stream (>= 0.67 GB per launch) exceeds the 126 MB L2: inputs larger than L2, no flush"}


# ------------------------------------------------------------------------------------- our arm
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r2_traffic.json")


def kernel_source_digest():
    """sha256 over the sources that define the dominant kernel (fused_update_kernel): the ncu DRAM-traffic figure
    quoted in `roofline.traffic` is only valid for the code it was captured on."""
    import hashlib
    h = hashlib.sha256()
    base = os.path.join(ROOT, "unified-unlearning-w-remain-geometry_b200", "csrc")
    for f in ("update.cu", "update_core.cuh", "common.cuh"):
        with open(os.path.join(base, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def recorded_traffic(n):
    """(bytes per launch | None, source note).  The value comes from a committed `ncu --set full` capture
    (dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel at this n); it is refused when the
    kernel's sources have changed since."""
    try:
        with open(TRAFFIC_FILE) as f:
            rec = json.load(f)
    except Exception:
        return None, "no committed ncu capture for this build"
    if rec.get("elements") != n:
        return None, f"capture was taken at n={rec.get('elements')}"
    if rec.get("kernel_source_sha256") != kernel_source_digest():
        return None, f"kernel sources changed since {rec.get('source')}"
    return float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"]), rec.get("source")


class Marks:
    """CUDA events recorded on the launching stream at named points; per-name mean interval since the previous mark."""

    def __init__(self):
        self.items = []

    def __call__(self, label):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.items.append((label, e))

    def intervals(self, names, steps):
        per = len(names) + 1
        assert len(self.items) == per * steps, (len(self.items), per, steps)
        out = {k: 0.0 for k in names}
        for s in range(steps):
            ev = self.items[s * per:(s + 1) * per]
            for i, k in enumerate(names):
                out[k] += ev[i][1].elapsed_time(ev[i + 1][1])
        return {k: v / steps for k, v in out.items()}


def timed_steps(step, marks_names, hp, steps, warmup, barrier, world, dev, sampler=None):
    """W warm-up + K timed steps; returns (ms_per_step max over ranks, per-mark ms of this rank, clocks)."""
    import torch.distributed as dist
    for _ in range(warmup):
        step()
    barrier()
    marks = Marks()
    hp.trace = marks
    barrier()
    wall_begin = time.time()
    start = torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        marks("step_start")
        step()
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    barrier()
    wall_end = time.time()
    hp.trace = None
    clocks = sampler.stop(wall_begin, wall_end) if sampler else None
    t = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps, marks.intervals(marks_names, steps), clocks


def fill_pinned(n, rank, dtype=torch.float32):
    host = torch.empty(n, dtype=dtype).pin_memory()
    blk_n = min(n, 1 << 24)   # host RNG over 675M elements is slow: draw 16 Mi values and tile them
    blk = (torch.randn(blk_n, generator=torch.Generator().manual_seed(7 + rank)) * 1e-2).to(dtype)
    for off in range(0, n, blk_n):
        m = min(blk_n, n - off)
        host[off:off + m].copy_(blk[:m])
    return host


STEP_MARKS = ["fisher_forget", "fisher_remain", "ratio_mask", "masked_sumsq", "fused_update_forget",
              "fused_update_remain_ema"]


def run_single(args, dev):
    """N = 1: the whole vector on one GPU (the configuration the metric is quoted on)."""
    import sfron_b200 as sfr
    n = args.elems
    opt = sfr.OptConfig(kind="adamw", lr=1e-4, weight_decay=0.0)
    hp = sfr.HotPath(n, dev, opt, ema_mode="dit", ema_a=0.9999)
    gen = torch.Generator(device=dev).manual_seed(1234)
    p = torch.randn(n, device=dev, generator=gen) * 0.02
    g_f = torch.randn(n, device=dev, generator=gen) * 1e-2
    g_r = torch.randn(n, device=dev, generator=gen) * 1e-2
    hp.init_slow(p)

    def step_on(gf, gr):
        hp.fisher_accumulate("forget", gf, FISHER_L)
        hp.fisher_accumulate("remain", gr, FISHER_L)
        hp.ratio_mask(1.0)
        hp.forget_step(p, gf, max_norm=1.0)
        hp.remain_step(p, gr, ema=True)

    def barrier():
        torch.cuda.synchronize()

    sampler = ClockSampler(dev.index)
    ms_per_step, kernel_ms, clocks = timed_steps(lambda: step_on(g_f, g_r), STEP_MARKS, hp, args.steps, args.warmup,
                                                 barrier, 1, dev, sampler)
    value = BYTES_PER_ELEM * n / (ms_per_step * 1e-3) / 1e9

    # ---- outside the step: the top-k select (K2b), which the DiT flow does not use (SalUn / DDPM generate_mask)
    extra = {}
    if not args.no_extra:
        topk = torch.empty(n, dtype=torch.uint8, device=dev)
        hp.topk_mask(g_f, n // 2, out=topk)
        times = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            hp.topk_mask(g_f, n // 2, out=topk)
            b.record()
            b.synchronize()
            times.append(a.elapsed_time(b))
        ms = sorted(times)[1]
        extra["topk_select_k_half"] = {
            "ms": round(ms, 4), "GBps_on_13B_per_elem": round(13 * n / (ms * 1e-3) / 1e9, 1),
            "GBps_on_bytes_moved": round(9 * n / (ms * 1e-3) / 1e9, 1), "bytes_per_elem_algorithmic": 13,
            "bytes_per_elem_moved": 9, "selected": int(topk.sum(dtype=torch.int64))}
        del topk

    # ---- e2e: gradients from pinned host memory, scalars read back, through the public API ------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    result_host = torch.empty(2, dtype=torch.float64).pin_memory()

    def e2e_run(host_dtype):
        host_g = fill_pinned(n, 0, host_dtype)
        feeder = sfr.HostGradientFeeder(n, dev, slots=("forget", "remain"), dtype=host_dtype)

        def e2e_step():
            g = feeder.acquire()                               # this step's gradients, copied from pinned host memory
            feeder.submit(forget=host_g, remain=host_g)        # next step's H2D overlaps this step's kernels
            step_on(g["forget"], g["remain"])
            feeder.release()
            res = torch.stack([hp.sumsq[0], hp.zero_count[0].double()])
            result_host.copy_(res, non_blocking=True)
            torch.cuda.current_stream().synchronize()          # the step's result is on the host

        # Pipelined: every step submits the NEXT step's host->device copy before running its kernels, so the
        # timed region holds exactly e2e_steps submissions and e2e_steps kernel passes; the copy that is still
        # in flight at the end is waited for inside the timed region.
        feeder.submit(forget=host_g, remain=host_g)
        e2e_step()                                             # warm-up step
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        return BYTES_PER_ELEM * n * e2e_steps / dt / 1e9, feeder.bytes_per_step

    e2e_value, h2d = e2e_run(torch.float32)
    if not args.no_extra:
        v16, h16 = e2e_run(torch.bfloat16)
        extra["e2e_bf16_host_gradients"] = {
            "value": round(v16, 1), "unit": UNIT, "h2d_bytes_per_step": h16,
            "note": "same step fed bf16 gradients from the host (half the PCIe bytes; K1/K3 widen them exactly; "
                    "BASELINE config 3's dtype) — the fp32 line above is the reference's own arithmetic"}

    peak, peak_kind = measured_peak()
    cpu = None
    if not args.no_cpu_baseline:
        gbs, ms = time_reference(args.ref_elems, args.cpu_steps, 2)
        cpu = {"value": gbs, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{args.ref_elems} of {N3} elements per step x {args.cpu_steps} steps, reference-form "
                         f"torch CPU ops (oracle/sfron_oracle.py), {ms:.0f} ms/step"}
    return dict(n_local=n, ms_per_step=ms_per_step, value=value, kernel_ms=kernel_ms, clocks=clocks, extra=extra,
                e2e={"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16,
                     "steps": e2e_steps},
                cpu=cpu, launches=(len(STEP_MARKS) + 1) * args.steps,      # + the one-thread update_consts kernel
                scaling="weak", exchange=None)


def run_dp(args, rank, world, dev):
    """N > 1: strong scaling on the same vector, data parallel (see the module docstring)."""
    import torch.distributed as dist
    import sfron_b200 as sfr
    from sfron_b200.dist import PeerExchange, ShardGroup, ShardedHotPath

    n = args.elems
    n_pad = -(-n // (16 * world)) * (16 * world)
    sg = ShardGroup(n, padded_len=n_pad)
    opt = sfr.OptConfig(kind="adamw", lr=1e-4, weight_decay=0.0)
    frac = (world - 1) / world

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    xchg, peer_error = None, None
    if args.transport != "nccl":
        try:
            xchg = PeerExchange(sg, dev, transport=args.transport)
        except Exception as e:                                  # no symmetric memory on this box: NCCL transport
            peer_error = repr(e)[:300]
        ok = torch.tensor([0 if xchg is None else 1], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0:
            xchg = None

    def gradients(dtype, seed):
        gen = torch.Generator(device=dev).manual_seed(seed + rank)
        t = torch.randn(n_pad, device=dev, generator=gen) * 1e-2
        t[n:].zero_()
        return t.to(dtype)

    weights0 = torch.randn(n_pad, device=dev, generator=torch.Generator(device=dev).manual_seed(1234)) * 0.02

    def build(transport, g_dtype, push):
        """One configured data-parallel step: returns (hot path, step(gf, gr), slots) where slots are the two
        pairs of per-rank gradient buffers (double-buffered for the host feeder)."""
        hp = ShardedHotPath(sg, dev, opt, ema_mode="dit", ema_a=0.9999)
        if transport == "nccl":
            w = weights0.clone()
            w16 = torch.empty(n_pad, dtype=torch.bfloat16, device=dev) if push == "bf16" else None
            slots = [{k: gradients(g_dtype, s) for k, s in (("forget", 11 + 100 * i), ("remain", 12 + 100 * i))}
                     for i in range(2)]
            p = w[sg.lo:sg.hi] if push == "f32" else w[sg.lo:sg.hi].clone()
            hp.init_slow(p)

            def step(gf, gr):
                gf_s = sg.reduce_scatter_gradients_(gf, average=True)
                hp._t("reduce_scatter_forget")
                hp.fisher_accumulate("forget", gf_s, FISHER_L)
                gr_s = sg.reduce_scatter_gradients_(gr, average=True)
                hp._t("reduce_scatter_remain")
                hp.fisher_accumulate("remain", gr_s, FISHER_L)
                hp.ratio_mask(1.0)
                hp.forget_step(p, gf_s, max_norm=1.0, p_bf16=None if w16 is None else w16[sg.lo:sg.hi])
                sg.all_gather_params_(w if w16 is None else w16)
                hp._t("all_gather_forget")
                hp.remain_step(p, gr_s, ema=True, p_bf16=None if w16 is None else w16[sg.lo:sg.hi])
                sg.all_gather_params_(w if w16 is None else w16)
                hp._t("all_gather_remain")

            names = ["reduce_scatter_forget", "fisher_forget", "reduce_scatter_remain", "fisher_remain", "ratio_mask",
                     "masked_sumsq", "fused_update_forget", "all_gather_forget", "fused_update_remain_ema",
                     "all_gather_remain"]
            return hp, step, slots, names
        hp.attach_exchange(xchg)
        xchg._want = transport
        w = xchg.alloc(n_pad, torch.float32) if push == "f32" else None
        w16 = xchg.alloc(n_pad, torch.bfloat16) if push == "bf16" else None
        if w is not None:
            w.tensor.copy_(weights0)
            p = w.tensor[sg.lo:sg.hi]
        else:
            w16.tensor.copy_(weights0)
            p = weights0[sg.lo:sg.hi].clone()                   # fp32 master shard; the ranks share bf16 working weights
        hp.init_slow(p)
        sym = []
        for i in range(2):
            pair = {}
            for k, s in (("forget", 11 + 100 * i), ("remain", 12 + 100 * i)):
                b = xchg.alloc(n_pad, g_dtype)
                b.tensor.copy_(gradients(g_dtype, s))
                pair[k] = b
            sym.append(pair)
        by_ptr = {b.tensor.data_ptr(): b for pair in sym for b in pair.values()}
        slots = [{k: b.tensor for k, b in pair.items()} for pair in sym]

        def step(gf, gr):
            gf, gr = by_ptr[gf.data_ptr()], by_ptr[gr.data_ptr()]
            # reduce-scatter fused with K1; the reduced shard stays local for the optimizer step on the same gradient
            hp.dp_fisher_accumulate("forget", gf, FISHER_L, keep="forget")
            hp.dp_fisher_accumulate("remain", gr, FISHER_L, keep="remain")
            hp.ratio_mask(1.0)
            hp.dp_forget_step(p, hp.reduced("forget"), weights=w, weights_bf16=w16, max_norm=1.0)
            hp.dp_remain_step(p, hp.reduced("remain"), weights=w, weights_bf16=w16, ema=True)

        return hp, step, slots, STEP_MARKS

    def measure(transport, g_dtype, push, steps, warmup, sampler=None, e2e_steps=0):
        hp, step, slots, names = build(transport, g_dtype, push)
        ms, marks, clocks = timed_steps(lambda: step(slots[0]["forget"], slots[0]["remain"]), names, hp, steps,
                                        warmup, barrier, world, dev, sampler)
        out = dict(ms_per_step=ms, marks=marks, clocks=clocks, hp=hp)
        if e2e_steps:
            host = fill_pinned(n, rank, g_dtype)
            feeder = sfr.HostGradientFeeder(n, dev, slots=("forget", "remain"), buffers=slots)
            result_host = torch.empty(2, dtype=torch.float64).pin_memory()

            def e2e_step():
                g = feeder.acquire()
                feeder.submit(forget=host, remain=host)
                step(g["forget"], g["remain"])
                feeder.release()
                res = torch.stack([hp.sumsq[0], hp.zero_count[0].double()])
                result_host.copy_(res, non_blocking=True)
                torch.cuda.current_stream().synchronize()

            feeder.submit(forget=host, remain=host)
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            barrier()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out["e2e"] = {"value": BYTES_PER_ELEM * n * e2e_steps / float(t.item()) / 1e9, "unit": UNIT,
                          "h2d_bytes_per_step": feeder.bytes_per_step * world,
                          "h2d_bytes_per_step_per_gpu": feeder.bytes_per_step, "d2h_bytes_per_step": 16 * world,
                          "steps": e2e_steps}
        if xchg is not None and transport != "nccl":
            xchg.check()
        return out

    if xchg is not None:
        xchg.alloc(16, torch.float32)                            # a data buffer, so "auto" can see whether multicast exists
    main_transport = "nccl" if xchg is None else xchg.transport_name
    sampler = ClockSampler(dev.index) if rank == 0 else None
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    main = measure(main_transport, torch.float32, "f32", args.steps, args.warmup, sampler, e2e_steps)
    value = BYTES_PER_ELEM * n / (main["ms_per_step"] * 1e-3) / 1e9

    es = 4
    red_t, push_t = (main_transport.split("+") * 2)[:2]
    # multimem: the switch fans in / out, so every byte of the vector leaves (reduce) or enters (push) each GPU once,
    # this GPU's own shard included; P2P: only the other ranks' shards cross, in both directions
    per_reduce_out, per_reduce_in = (n_pad * es, n_pad * es / world) if red_t == "multimem" else (frac * n_pad * es,) * 2
    per_push_out, per_push_in = (n_pad * 4 / world, n_pad * 4) if push_t == "multimem" else (frac * n_pad * 4,) * 2
    out_b, in_b = 2 * per_reduce_out + 2 * per_push_out, 2 * per_reduce_in + 2 * per_push_in
    exchange = {
        "transport": main_transport, "peer_error": peer_error,
        "what": "per step and per GPU: 2 gradient reduces (fused with K1) + 2 weight pushes (fused with K3), "
                "7 cross-GPU barriers, the clip norm summed on one of them",
        "nvlink_out_bytes_per_gpu": out_b, "nvlink_in_bytes_per_gpu": in_b, "push_in_bytes_per_gpu": per_push_in,
        "link_GBps_per_direction": max(out_b, in_b) / (main["ms_per_step"] * 1e-3) / 1e9,
        "link_reference_GBps": 770.0,
        "note": "link rate over the WHOLE step (shard-local K2a / norm time included): a lower bound on the "
                "exchange kernels' own rate; per-op rates: profiles/r2_xchg_*.jsonl (tools/xchg_bench.py)"}

    extra = {}
    if not args.no_extra:
        k = max(3, min(args.steps, 10))
        if xchg is not None:
            r = measure("nccl", torch.float32, "f32", k, 3)
            extra["nccl_transport"] = {
                "ms_per_step": round(r["ms_per_step"], 4), "value": round(BYTES_PER_ELEM * n / r["ms_per_step"] / 1e6, 1),
                "marks_ms": {a: round(b, 4) for a, b in r["marks"].items()},
                "note": "same data-parallel step with torch.distributed reduce_scatter_tensor / all_gather_into_tensor "
                        "(NCCL) around the shard-local kernels"}
            del r
            others = [t for t in ("p2p", "tma") + (("multimem",) if xchg.pad.has_multicast else ()) if t != main_transport]
            for other in others:
                r = measure(other, torch.float32, "f32", k, 3)
                extra[f"{other}_transport"] = {"ms_per_step": round(r["ms_per_step"], 4),
                                               "value": round(BYTES_PER_ELEM * n / r["ms_per_step"] / 1e6, 1),
                                               "marks_ms": {a: round(b, 4) for a, b in r["marks"].items()}}
                del r
            r = measure(main_transport, torch.bfloat16, "bf16", k, 3)
            extra["bf16_exchange"] = {
                "ms_per_step": round(r["ms_per_step"], 4), "bytes_per_element": 97,
                "value": round(97 * n / r["ms_per_step"] / 1e6, 1),
                "marks_ms": {a: round(b, 4) for a, b in r["marks"].items()},
                "note": "BASELINE config 3 (bf16): bf16 gradients reduced over NVLink, fp32 master / moments / EMA "
                        "sharded, bf16 working weights pushed to all ranks — half the NVLink bytes"}
            del r
        torch.cuda.empty_cache()
        # r1's line: weak scaling, every rank owns an N3-element shard of an N x N3 vector, shard-local kernels,
        # only the clip-norm and zero-count scalars cross GPUs
        hpw = ShardedHotPath(ShardGroup(n * world), dev, opt, ema_mode="dit", ema_a=0.9999)
        nl = hpw.n
        gen = torch.Generator(device=dev).manual_seed(99 + rank)
        pw = torch.randn(nl, device=dev, generator=gen) * 0.02
        gfw = torch.randn(nl, device=dev, generator=gen) * 1e-2
        grw = torch.randn(nl, device=dev, generator=gen) * 1e-2
        hpw.init_slow(pw)

        def weak_step():
            hpw.fisher_accumulate("forget", gfw, FISHER_L)
            hpw.fisher_accumulate("remain", grw, FISHER_L)
            hpw.ratio_mask(1.0)
            hpw.forget_step(pw, gfw, max_norm=1.0)
            hpw.remain_step(pw, grw, ema=True)

        ms, marks, _ = timed_steps(weak_step, STEP_MARKS, hpw, k, 3, barrier, world, dev)
        extra["weak_shard_local"] = {
            "ms_per_step": round(ms, 4), "elements_per_gpu": nl,
            "value": round(BYTES_PER_ELEM * nl * world / ms / 1e6, 1),
            "note": "round-1 headline: N x N3-element vector, one N3 shard per GPU, no gradient exchange"}
        del hpw, pw, gfw, grw

    # kernels of the library per step: K1 x2, K2a, norm, consts, K3 x2 (= 7), + 7 barrier kernels on the peer transports
    launches_per_step = 7 if main_transport == "nccl" else 14
    return dict(n_local=sg.n_local, ms_per_step=main["ms_per_step"], value=value, kernel_ms=main["marks"],
                clocks=main["clocks"], extra=extra, e2e=main["e2e"], cpu=None,
                launches=launches_per_step * args.steps, scaling="strong", exchange=exchange)


def run_ours(args, rank, world, local_rank):
    import sfron_b200 as sfr

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the SFR-on hot path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    sfr.capi.load()
    r = run_single(args, dev) if world == 1 else run_dp(args, rank, world, dev)
    if rank != 0:
        return
    n = args.elems
    peak, peak_kind = measured_peak()
    dom = "fused_update_remain_ema"
    nl = r["n_local"]
    kernel_ms = r["kernel_ms"]
    per_elem = dict(BYTES)
    achieved = BYTES[dom] * nl / (kernel_ms[dom] * 1e-3) / 1e9
    kernels = {}
    for k, v in kernel_ms.items():
        rec = {"ms": round(v, 4)}
        if k in per_elem:
            rec.update(GBps=round(per_elem[k] * nl / (v * 1e-3) / 1e9, 1),
                       frac=round(per_elem[k] * nl / (v * 1e-3) / 1e9 / peak, 4), bytes_per_elem=per_elem[k])
        kernels[k] = rec
    traffic, traffic_src = recorded_traffic(nl) if world == 1 else (None, "single-GPU capture only")
    xk = {"tma": "fused_update_tma_kernel", "nccl": "fused_update_kernel"}.get(
        (r["exchange"] or {}).get("transport", "").split("+")[-1], "fused_update_xchg_kernel")
    roofline = {"bound": "hbm", "kernel": "fused_update_kernel<AdamW,EMA_DIT,f32> (remain step + EMA)"
                if world == 1 else f"{xk}<AdamW,EMA_DIT> (remain step + EMA + weight push; `achieved` = its HBM "
                "bytes on the shard; the kernel is NVLink-bound: see `nvlink` below and `exchange`)",
                "achieved": achieved, "peak": peak,
                "peak_source": f"MEASURED_PEAKS.json ({peak_kind})" if peak_kind == "measured" else "fallback 6.65 TB/s",
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src}
    if world > 1 and r["exchange"] and r["exchange"]["transport"] != "nccl":
        # the same kernel against the link: bytes of the weight push this GPU must receive / its device time
        push_in = r["exchange"]["push_in_bytes_per_gpu"]
        link = push_in / (kernel_ms[dom] * 1e-3) / 1e9
        roofline["nvlink"] = {"bound": "nvlink", "achieved": link, "peak": 770.0, "unit": "GB/s", "frac": link / 770.0,
                              "peak_source": "peer-copy figure of B200_PROFILING.md (NCCL collectives reach 510-670 "
                                             "GB/s bus bandwidth on these boxes)"}
    line = {
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": r["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(n, world),
        "roofline": roofline,
        "hot_path_frac_of_peak": r["value"] / world / peak,
        "kernels": kernels,
        "exchange": r["exchange"],
        "extra": r["extra"],
        "cpu_baseline": r["cpu"],
        "e2e": r["e2e"],
        "gpu_launches": r["launches"],
        "steps_per_s": 1e3 / r["ms_per_step"],
        "clocks": r["clocks"],
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--elems", type=int, default=N3, help="elements of the vector (default: DiT-XL/2, 675,129,632)")
    ap.add_argument("--ref-elems", type=int, default=1 << 26, help="bounded CPU sample per step")
    ap.add_argument("--cpu-steps", type=int, default=6)
    ap.add_argument("--e2e-steps", type=int, default=8)
    ap.add_argument("--transport", default="auto",
                    help="N > 1: how gradients and weights cross GPUs (auto: the library's peer-memory kernels, "
                         "multimem when the fabric has multicast)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the comparison lines outside the main measurement")
    args = ap.parse_args()
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
        # stdout carries exactly one JSON line: NCCL's init lines (rank / nranks, transports) go to stderr
        os.environ["NCCL_DEBUG"] = os.environ.get("SFR_NCCL_DEBUG", "INFO")
        os.environ["NCCL_DEBUG_SUBSYS"] = os.environ.get("SFR_NCCL_DEBUG_SUBSYS", "INIT")
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
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
