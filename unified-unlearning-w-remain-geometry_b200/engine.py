"""Host side of the SFR-on hot path over flat device vectors.

`HotPath` owns the per-role flat buffers of ONE shard (Fisher accumulators, mask, optimizer
state, slow/EMA weights) and sequences the C-ABI kernels in the order the reference's loops
do their torch ops.  It holds no model: callers hand it the flat weight / gradient vectors
(`FlatParams.p`, `.g`, or any 16-byte-aligned slices of them when sharded).

Reference order of operations:
  Fisher    F += grad**2 / L                     per batch          -> fisher_accumulate()
  mask      (F_f+1e-15)/(F_r+1e-15) >= th        once per threshold -> ratio_mask()
            top-k of |sum grad|                  SalUn              -> topk_mask()
  forget    grad *= mask ; clip ; optimizer.step                    -> forget_step()
  remain    [clip ;] optimizer.step ; EMA / slow-fast               -> remain_step()
  SalUn     clip ; grad *= mask ; optimizer.step ; EMA (joint loss) -> joint_step()
One optimizer state serves both steps (`step` advances twice per iteration), exactly as the
single torch optimizer object of every reference loop.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import torch

from . import capi


@dataclass
class OptConfig:
    """Hyper-parameters of the optimizer the reference constructs.

    kind: "sgd" (sfron.py:167-170), "adam" (DDPM/functions/__init__.py:9-18, SD nsfw_removal.py:81),
    "adamw" (DiT/forget.py:199)."""
    kind: str = "adamw"
    lr: float = 1e-4
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    weight_decay: float = 0.0
    momentum: float = 0.0
    dampening: float = 0.0

    def code(self) -> int:
        return {"sgd": capi.OPT_SGD, "adam": capi.OPT_ADAM, "adamw": capi.OPT_ADAMW}[self.kind]


_EMA_CODES = {"none": capi.EMA_NONE, "ddpm": capi.EMA_DDPM, "dit": capi.EMA_DIT, "slowfast": capi.EMA_SLOWFAST}


class HotPath:
    def __init__(self, n: int, device, opt: OptConfig, *, ema_mode: str = "none", ema_a: float = 0.0,
                 grad_dtype: torch.dtype = torch.float32):
        self.n = int(n)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise capi.SfrError(capi.ERR_NO_DEVICE, "HotPath", "the SFR-on hot path runs on CUDA only")
        capi.load()
        self.opt = opt
        self.ema_mode = ema_mode
        self.ema_a = float(ema_a)
        self.grad_dtype = grad_dtype
        self.step_count = 0            # torch optimizer state['step'] (shared by forget and remain)
        self._buf: Dict[str, torch.Tensor] = {}
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.zero_count = torch.zeros(capi.MAX_THRESHOLDS, dtype=torch.int64, device=self.device)
        self._select = None
        # the whole select (every launch is stream-ordered, nothing is read back) replayed from one CUDA graph per
        # (buffers, n, k): the first call runs eagerly, the second captures, later ones replay.  Off when the select
        # spans ranks (its collectives are torch.distributed calls) — dist.ShardedHotPath clears the flag.
        # device-side learning-rate schedule (set_lr_schedule): table of per-iteration rates + the iteration index
        self.lr_table: Optional[torch.Tensor] = None
        self.lr_index: Optional[torch.Tensor] = None
        # clipped steps of vectors up to this size run as ONE cooperative launch (sfr_clipped_update).  Measured
        # (profiles/r2_sweep_small.jsonl): on the device the one launch costs what the four stream-ordered launches cost
        # at 1e7 elements (70.6 vs 68.6 us) and 5-7 % MORE at 3.9e7 / 1e8 (a persistent grid streams slower than one
        # tile per CTA), so it is used where the saving is on the HOST side — three fewer launches per step in loops
        # that are CPU-bound at this size (ResNet-18) — and nowhere else.
        self.coop_max_elems = 20_000_000
        self.select_graphs = True
        self._select_graph_cache: Dict[tuple, object] = {}
        # a saliency mask exists only once it was BUILT (ratio_mask / topk_mask) or LOADED (set_buffer / load_mask):
        # forget_step(use_mask=True) without one raises instead of multiplying every gradient by a zero buffer
        # (the reference applies no mask at all when mask_path is unset: DDPM/runners/diffusion.py:1050)
        self._mask_ready = False
        self.select_two_pass = True    # False: pass 1 without the provisional mask, apply streams the vector again
        # replay-safe mode (CUDA graphs): the optimizer step counter lives on the device
        self.step_dev: Optional[torch.Tensor] = None
        # device scratch for the prep kernel of the fused update (step-dependent scalars, clip coefficient)
        self._consts_dev: Optional[torch.Tensor] = torch.zeros(128, dtype=torch.uint8, device=self.device)
        # optional probe called with a label right after each kernel launch (bench.py records a
        # CUDA event there to attribute device time per kernel); None on the normal path
        self.trace = None

    def _t(self, label: str) -> None:
        if self.trace is not None:
            self.trace(label)

    def enable_graph_replay(self) -> None:
        """Make forget_step / remain_step safe to capture in a CUDA graph and replay: the step count
        (Adam bias corrections, SGD's first-step momentum-buffer init) moves to a device counter that a
        one-thread kernel advances on every launch/replay.  Allocates everything the steps need up front
        (nothing may be allocated during capture)."""
        if self.step_dev is None:
            self.step_dev = torch.full((1,), self.step_count, dtype=torch.int64, device=self.device)
        sgd = self.opt.kind == "sgd"
        if not sgd or self.opt.momentum != 0.0:
            self.buffer("m")
        if not sgd:
            self.buffer("v")
        if self.ema_mode != "none":
            self.buffer("slow")

    def set_lr_schedule(self, rates: Sequence[float]) -> None:
        """Learning rate per ITERATION, kept on the device: every step reads `rates[lr_index]` in its scalar-prep
        kernel, so a CUDA graph of one iteration follows the schedule at every replay (the classification loop's
        per-iteration CosineAnnealingLR, sfron.py:172-174,259).  `advance_lr()` moves to the next iteration."""
        self.lr_table = torch.tensor(list(rates), dtype=torch.float64, device=self.device)
        self.lr_index = torch.zeros(1, dtype=torch.int64, device=self.device)

    def advance_lr(self) -> None:
        """scheduler.step(): stream-ordered, capturable."""
        self.lr_index.add_(1)

    # ---- lazily allocated role buffers ------------------------------------------------------------
    def buffer(self, role: str, dtype=torch.float32) -> torch.Tensor:
        t = self._buf.get(role)
        if t is None:
            t = torch.zeros(self.n, dtype=dtype, device=self.device)
            self._buf[role] = t
        return t

    def has(self, role: str) -> bool:
        return role in self._buf

    def set_buffer(self, role: str, t: torch.Tensor) -> None:
        if t.numel() != self.n or t.device != self.device:
            raise ValueError(f"buffer {role}: wrong size or device")
        self._buf[role] = t.contiguous()
        if role == "mask":
            self._mask_ready = True

    @property
    def forget_fisher(self) -> torch.Tensor:
        return self.buffer("forget_fisher")

    @property
    def remain_fisher(self) -> torch.Tensor:
        return self.buffer("remain_fisher")

    @property
    def mask(self) -> torch.Tensor:
        return self.buffer("mask", torch.uint8)

    @property
    def m(self) -> torch.Tensor:
        return self.buffer("m")

    @property
    def v(self) -> torch.Tensor:
        return self.buffer("v")

    @property
    def slow(self) -> torch.Tensor:
        return self.buffer("slow")

    def mark_mask_ready(self) -> None:
        """For callers that filled `self.mask` themselves."""
        self._mask_ready = True

    def require_mask(self) -> torch.Tensor:
        if not self._mask_ready:
            raise capi.SfrError(
                capi.ERR_ARG, "HotPath", "use_mask=True but no saliency mask has been built (ratio_mask / topk_mask) "
                "or loaded (set_buffer('mask', ...) / load_mask); pass use_mask=False for the reference's unmasked "
                "run (mask_path unset)")
        return self.mask

    def init_slow(self, p: torch.Tensor) -> None:
        """EMAHelper.register / `ema = deepcopy(model)` / `ori_model = deepcopy(model)`:
        the slow weights start as a copy of the weights."""
        self.buffer("slow").copy_(p)

    # ---- K1 ---------------------------------------------------------------------------------------
    def fisher_accumulate(self, which: str, g: torch.Tensor, divisor: float, *,
                          clip_max_norm: Optional[float] = None) -> None:
        """`F[which] += g**2 / divisor`; with `clip_max_norm` the gradient is clipped first, as in
        DDPM/runners/diffusion.py:1270-1281."""
        acc = self.buffer({"forget": "forget_fisher", "remain": "remain_fisher"}.get(which, which))
        if clip_max_norm is None:
            capi.fisher_accum(acc, g, divisor)
            self._t("fisher_accum")
        else:
            self.sumsq.zero_()
            capi.masked_sumsq(g if g.dim() == 1 else g.reshape(-1), None, self.sumsq)
            self.reduce_scalar_(self.sumsq)
            capi.fisher_accum(acc, g, divisor, clip_sumsq=self.sumsq, clip_max_norm=clip_max_norm)
            self._t("fisher_accum_clipped")

    def saliency_accumulate(self, g: torch.Tensor, *, clip_max_norm: Optional[float] = None,
                            role: str = "grad_sum") -> torch.Tensor:
        """`clip_grad_norm_ ; gradients[name] += param.grad` of the SalUn mask generation
        (DDPM/runners/diffusion.py:985-994; salun.py:163-169 without the clip): one norm pass + ONE
        accumulate pass (12 B/elem), no temporary."""
        acc = self.buffer(role)
        if clip_max_norm is None:
            capi.grad_accum(acc, g)
        else:
            self.sumsq.zero_()
            capi.masked_sumsq(g, None, self.sumsq)
            self.reduce_scalar_(self.sumsq)
            capi.grad_accum(acc, g, clip_sumsq=self.sumsq, clip_max_norm=clip_max_norm)
        self._t("saliency_accum")
        return acc

    # ---- K2a --------------------------------------------------------------------------------------
    def ratio_mask(self, threshold: float, *, eps: float = 1e-15, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Writes the bool mask (uint8 storage) and leaves the zero count in `self.zero_count[0]`."""
        mask = self.mask if out is None else out
        self.zero_count.zero_()
        capi.ratio_mask(self.forget_fisher, self.remain_fisher, threshold, mask, self.zero_count, eps)
        self._t("ratio_mask")
        if out is None:
            self._mask_ready = True
        return mask

    def ratio_masks(self, thresholds: Sequence[float], *, eps: float = 1e-15) -> torch.Tensor:
        """All thresholds in one pass over the two Fisher vectors -> [T, stride] uint8."""
        stride = (self.n + 15) // 16 * 16
        masks = torch.empty(len(thresholds), stride, dtype=torch.uint8, device=self.device)
        self.zero_count.zero_()
        capi.ratio_mask_multi(self.forget_fisher, self.remain_fisher, thresholds, masks, self.zero_count, eps)
        return masks

    # ---- K2b --------------------------------------------------------------------------------------
    def _select_buffers(self):
        if self._select is None:
            state = torch.zeros(capi.SELECT_STATE_BYTES // 8, dtype=torch.int64, device=self.device)
            bins = torch.zeros(capi.SELECT_BINS_ALLOC, dtype=torch.int64, device=self.device)
            scratch = torch.zeros(capi.select_scratch_elems(self.n), dtype=torch.int64, device=self.device)
            self._select = (state, bins, scratch)
        return self._select

    def topk_mask(self, values: torch.Tensor, k: int, *, other: Optional[torch.Tensor] = None,
                  eps: float = 1e-15, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Exact global top-k of |values| (or of the ratio (values+eps)/(other+eps)); ties at the
        threshold go to the lowest flat index.  `k` is the GLOBAL count across shards."""
        mode = capi.KEY_ABS if other is None else capi.KEY_RATIO
        state, bins, scratch = self._select_buffers()
        mask = self.mask if out is None else out

        def launches():
            capi.select_init(state, bins, k)
            capi.select_hist(values, other, mode, 0, state, bins, None, eps)
            self.reduce_bins_(bins, capi.SELECT_BINS0)
            capi.select_scan(0, state, bins)
            # pass 1 also leaves the provisional mask, so apply only has to resolve the staged candidates
            capi.select_hist(values, other, mode, 1, state, bins, scratch, eps,
                             mask=mask if self.select_two_pass else None)
            local_bins = self.keep_local_bins_(bins)
            self.reduce_bins_(bins, capi.SELECT_BINS1)
            capi.select_scan(1, state, bins)
            tie_base = self.tie_base_(state, local_bins)
            capi.select_apply(values, other, mode, state, tie_base, scratch, mask, eps)

        if self.select_graphs and not torch.cuda.is_current_stream_capturing():
            key = (values.data_ptr(), None if other is None else other.data_ptr(), mask.data_ptr(), values.numel(),
                   int(k), mode, float(eps), self.select_two_pass)
            entry = self._select_graph_cache.get(key)
            if entry is None:
                launches()
                if len(self._select_graph_cache) >= 4:                      # a handful of live (buffers, k) pairs
                    self._select_graph_cache.pop(next(iter(self._select_graph_cache)))
                self._select_graph_cache[key] = "warm"
            else:
                if entry == "warm":
                    with torch.cuda.device(self.device):
                        entry = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(entry):
                            launches()
                    self._select_graph_cache[key] = entry
                entry.replay()
        else:
            launches()
        if out is None:
            self._mask_ready = True
        return mask

    def kth_smallest_absdiff(self, a: torch.Tensor, b: torch.Tensor, k: int) -> torch.Tensor:
        """The k-th smallest |a - b| as a device fp32 scalar (k is global, 1-based): the threshold of
        SD/train-scripts/proximal_gradient.py:161 (`-topk(-|theta - theta0|, k)[0][-1]`)."""
        state, bins, scratch = self._select_buffers()
        n_total = self.total_elements()
        capi.select_init(state, bins, n_total - k + 1)          # k-th smallest == (n-k+1)-th largest
        capi.select_hist(a, b, capi.KEY_ABSDIFF, 0, state, bins, None)
        self.reduce_bins_(bins, capi.SELECT_BINS0)
        capi.select_scan(0, state, bins)
        capi.select_hist(a, b, capi.KEY_ABSDIFF, 1, state, bins, scratch)
        self.reduce_bins_(bins, capi.SELECT_BINS1)
        capi.select_scan(1, state, bins)
        out = torch.empty(1, dtype=torch.float32, device=self.device)
        capi.select_threshold_value(state, out)
        return out

    def proximal_shrink(self, p: torch.Tensor, p0: torch.Tensor, k: int) -> torch.Tensor:
        """proximal_gradient.py:151-183: soft-threshold theta - theta0 at its k-th smallest magnitude."""
        thr = self.kth_smallest_absdiff(p, p0, k)
        capi.soft_threshold(p, p0, thr)
        return thr

    def ewc_penalty(self, p: torch.Tensor, p_star: torch.Tensor, fisher: torch.Tensor, g: torch.Tensor,
                    lmbda: float) -> torch.Tensor:
        """runners/diffusion.py:424-433: adds the EWC gradient into g; returns the penalty (device double)."""
        pen = torch.zeros(1, dtype=torch.float64, device=self.device)
        capi.ewc_penalty(p, p_star, fisher, g, lmbda, pen)
        self.reduce_scalar_(pen)
        return pen

    def total_elements(self) -> int:
        return self.n

    def select_state(self) -> capi.SelectState:
        return capi.read_select_state(self._select_buffers()[0])

    # ---- collectives: identity on one GPU, overridden by dist.ShardedHotPath --------------------------
    def reduce_scalar_(self, t: torch.Tensor) -> None:
        pass

    def reduce_bins_(self, bins: torch.Tensor, count: int) -> None:
        pass

    def keep_local_bins_(self, bins: torch.Tensor) -> Optional[torch.Tensor]:
        return None

    def tie_base_(self, state: torch.Tensor, local_bins: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        return None

    # ---- clip norm + K3 ---------------------------------------------------------------------------
    def _args(self, flags: int, ema: bool, max_norm: Optional[float], lr: Optional[float]) -> capi.UpdateArgs:
        o = self.opt
        a = capi.UpdateArgs()
        a.opt = o.code()
        a.ema_mode = _EMA_CODES[self.ema_mode] if ema else capi.EMA_NONE
        a.flags = flags
        a.step = self.step_count
        a.lr = o.lr if lr is None else lr
        a.beta1, a.beta2, a.eps = o.beta1, o.beta2, o.eps
        a.weight_decay, a.momentum, a.dampening = o.weight_decay, o.momentum, o.dampening
        a.ema_a = self.ema_a
        a.clip_max_norm = 0.0 if max_norm is None else max_norm
        if self.lr_table is not None and lr is None:
            a.lr_table_dev, a.lr_index_dev = self.lr_table.data_ptr(), self.lr_index.data_ptr()
        return a

    def _step(self, p: torch.Tensor, g: torch.Tensor, *, mask: Optional[torch.Tensor], mask_order: str,
              max_norm: Optional[float], ema: bool, lr: Optional[float], zero_grad: bool,
              p_bf16: Optional[torch.Tensor], norm_sq: Optional[torch.Tensor] = None) -> None:
        """norm_sq (device float64[1]): the squared gradient norm to clip by, when the caller already has it (a norm
        reduced elsewhere; the parity tests feed the reference's own fp32 norm to isolate the update arithmetic
        from torch's CPU norm error).  Default: computed here by the masked sum-of-squares kernel."""
        flags = 0
        if mask is not None:
            flags |= capi.F_MASK if mask_order == "mask_then_clip" else capi.F_MASK_AFTER_CLIP
        if zero_grad:
            flags |= capi.F_ZERO_GRAD
        if p_bf16 is not None:
            flags |= capi.F_WRITE_BF16
        sgd = self.opt.kind == "sgd"
        if self.step_dev is None and sgd and self.opt.momentum != 0.0 and not self.has("m"):
            flags |= capi.F_SGD_FIRST_STEP      # torch creates momentum_buffer = clone(grad) on first use
        sharded = type(self).reduce_scalar_ is not HotPath.reduce_scalar_
        if max_norm is not None and norm_sq is None and not sharded and self.n <= self.coop_max_elems:
            # small vector on one GPU: zero + norm + scalars + update as ONE cooperative launch (launch-bound otherwise)
            self.step_count += 1
            a = self._args(flags, ema, max_norm, lr)
            use_ema = ema and self.ema_mode != "none"
            capi.clipped_update(p, g, None if (sgd and self.opt.momentum == 0.0) else self.m, None if sgd else self.v,
                                mask, self.slow if use_ema else None, a, self.sumsq, p_bf16=p_bf16,
                                step_counter=self.step_dev)
            self._t("masked_sumsq")
            self._t("fused_update_ema" if use_ema else "fused_update")
            return
        clip = None
        if max_norm is not None and norm_sq is not None:
            self.sumsq.copy_(norm_sq.reshape(1))
            self._t("masked_sumsq")
            clip = self.sumsq
        elif max_norm is not None:
            # norm of the gradient as clip_grad_norm_ sees it: masked already (SFR-on order) or raw
            self.sumsq.zero_()
            capi.masked_sumsq(g, mask if (mask is not None and mask_order == "mask_then_clip") else None, self.sumsq)
            self.reduce_scalar_(self.sumsq)
            self._t("masked_sumsq")
            clip = self.sumsq
        self.step_count += 1
        a = self._args(flags, ema, max_norm, lr)
        use_ema = ema and self.ema_mode != "none"
        capi.fused_update(p, g, None if (sgd and self.opt.momentum == 0.0) else self.m,
                          None if sgd else self.v, mask, self.slow if use_ema else None, a,
                          clip_sumsq=clip, p_bf16=p_bf16, step_counter=self.step_dev,
                          # the prep kernel pays off when there is a clip coefficient to precompute or a
                          # device step counter to advance; otherwise everything goes by value
                          consts_scratch=self._consts_dev if (clip is not None or self.step_dev is not None
                                                              or self.lr_table is not None) else None)
        self._t("fused_update_ema" if use_ema else "fused_update")

    def forget_step(self, p: torch.Tensor, g: torch.Tensor, *, mask: Optional[torch.Tensor] = None,
                    use_mask: bool = True, max_norm: Optional[float] = None, lr: Optional[float] = None,
                    mask_order: str = "mask_then_clip", zero_grad: bool = False,
                    p_bf16: Optional[torch.Tensor] = None, norm_sq: Optional[torch.Tensor] = None) -> None:
        """grad *= mask ; clip_grad_norm_(max_norm) ; optimizer.step()
        (sfron.py:201-206; runners/diffusion.py:1126-1138; DiT/forget.py:289-299)."""
        if mask is None and use_mask:
            mask = self.require_mask()
        self._step(p, g, mask=mask if use_mask else None, mask_order=mask_order, max_norm=max_norm,
                   ema=False, lr=lr, zero_grad=zero_grad, p_bf16=p_bf16, norm_sq=norm_sq)

    def remain_step(self, p: torch.Tensor, g: torch.Tensor, *, max_norm: Optional[float] = None,
                    lr: Optional[float] = None, ema: bool = True, zero_grad: bool = False,
                    p_bf16: Optional[torch.Tensor] = None, norm_sq: Optional[torch.Tensor] = None) -> None:
        """[clip ;] optimizer.step() ; EMA / slow-fast update
        (sfron.py:213-222,255-257; runners/diffusion.py:1156-1180; DiT/forget.py:310-322)."""
        self._step(p, g, mask=None, mask_order="mask_then_clip", max_norm=max_norm, ema=ema, lr=lr,
                   zero_grad=zero_grad, p_bf16=p_bf16, norm_sq=norm_sq)

    def joint_step(self, p: torch.Tensor, g: torch.Tensor, *, mask: Optional[torch.Tensor] = None,
                   use_mask: bool = True, max_norm: Optional[float] = None, lr: Optional[float] = None,
                   mask_order: str = "clip_then_mask", ema: bool = True, zero_grad: bool = False,
                   p_bf16: Optional[torch.Tensor] = None) -> None:
        """SalUn's single step on the joint forget+remain loss: clip_grad_norm_ ; grad *= mask ;
        optimizer.step() ; EMA (runners/diffusion.py:575-594 — the clip precedes the mask there)."""
        if mask is None and use_mask:
            mask = self.require_mask()
        self._step(p, g, mask=mask if use_mask else None, mask_order=mask_order, max_norm=max_norm,
                   ema=ema, lr=lr, zero_grad=zero_grad, p_bf16=p_bf16)

    def grad_norm(self) -> torch.Tensor:
        """Total norm of the last clipped step (what clip_grad_norm_ returns), fp32 device scalar."""
        return self.sumsq.sqrt().float()

    def ema_only(self, p: torch.Tensor, slow: torch.Tensor) -> None:
        """EMA of parameters the optimizer never touches (frozen DiT pos_embed, DiT/forget.py:58-62)."""
        capi.ema_update(p, slow, _EMA_CODES[self.ema_mode], self.ema_a)


class HostGradientFeeder:
    """Host-buffer entry of the path: gradients that arrive in PINNED HOST memory (a host-side or
    off-device producer — the direction in which the reference moves them every step, D2H then CPU math)
    are prefetched to the device on a copy stream, double-buffered, so the H2D of step i+1 overlaps the
    kernels of step i.  Steady state is bound by the PCIe copy alone.

        feeder = HostGradientFeeder(n, device, slots=("forget", "remain"))
        feeder.submit(forget=host_gf, remain=host_gr)            # step 0
        for step in ...:
            g = feeder.acquire()                                 # device views of this step's gradients
            feeder.submit(forget=..., remain=...)                # next step's copy starts now
            hot_path.fisher_accumulate("forget", g["forget"], L) ; ...
            feeder.release()                                     # buffers may be overwritten
    """

    def __init__(self, n: int, device, slots: Sequence[str] = ("forget", "remain"),
                 dtype: torch.dtype = torch.float32, depth: int = 2,
                 buffers: Optional[Sequence[Dict[str, torch.Tensor]]] = None):
        """dtype: fp32, or bf16 host gradients (half the PCIe bytes; the kernels widen them exactly).
        buffers: `depth` dicts slot -> device tensor to copy into instead of allocating — e.g. views of
        symmetric (peer-mapped) gradient buffers, so the data-parallel kernels read them in place."""
        self.device = torch.device(device)
        self.slots = tuple(slots)
        self.depth = depth if buffers is None else len(buffers)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        if buffers is None:
            buffers = [{s: torch.empty(n, dtype=dtype, device=self.device) for s in self.slots}
                       for _ in range(depth)]
        self.buffers = [dict(b) for b in buffers]
        self.ready = [torch.cuda.Event() for _ in range(self.depth)]
        self.free = [torch.cuda.Event() for _ in range(self.depth)]
        self._submitted = 0
        self._acquired = 0
        self.current_index = -1
        first = self.buffers[0][self.slots[0]]
        self.bytes_per_step = n * first.element_size() * len(self.slots)

    def submit(self, **host_tensors: torch.Tensor) -> None:
        if self._submitted - self._acquired >= self.depth:
            raise RuntimeError("feeder full: acquire()/release() a step before submitting another")
        i = self._submitted % self.depth
        with torch.cuda.stream(self.copy_stream):
            if self._submitted >= self.depth:
                self.copy_stream.wait_event(self.free[i])       # the kernels that read this buffer are done
            for s in self.slots:
                h = host_tensors[s]
                if not h.is_pinned():
                    raise ValueError(f"{s}: host gradient buffers must be pinned for an asynchronous copy")
                dst = self.buffers[i][s]
                dst[:h.numel()].copy_(h, non_blocking=True)      # (a padded device buffer keeps its tail)
            self.ready[i].record(self.copy_stream)
        self._submitted += 1

    def acquire(self) -> Dict[str, torch.Tensor]:
        if self._acquired >= self._submitted:
            raise RuntimeError("nothing submitted")
        i = self._acquired % self.depth
        torch.cuda.current_stream(self.device).wait_event(self.ready[i])
        self._current = self.current_index = i
        self._acquired += 1
        return self.buffers[i]

    def release(self) -> None:
        self.free[self._current].record(torch.cuda.current_stream(self.device))
