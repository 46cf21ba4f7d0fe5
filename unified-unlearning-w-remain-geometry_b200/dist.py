"""Multi-GPU plumbing of the hot path: one process per GPU, the flat vector sharded across ranks.

Replaces the reference's single-process `torch.nn.DataParallel` (DDPM/runners/diffusion.py:110,...,
DiT/forget.py:193, DiT/generate_fisher.py:173), whose replicate / scatter / gather / reduce_add runs
inside torch on GPU 0.  Here:

  * every kernel runs shard-local on the rank's contiguous, 16-element-aligned slice [lo, hi);
  * the only data-path collectives are
      - gradients: all-reduce (replicated update) or reduce-scatter (sharded update) of the flat
        gradient produced by each rank's backward pass               -> `reduce_gradients_`
        (`BucketedGradReducer`: the same all-reduce cut into buckets that start while backward still runs)
      - clip norm: all-reduce of ONE double (the masked sum of squares) -> `reduce_scalar_`
      - top-k select: all-reduce of the histogram bins (256 KB, then 512 KB) and an all-gather of
        one tie count per rank                                        -> `reduce_bins_`, `tie_base_`
      - mask statistics: all-reduce of the zero counts (<= 8 x u64)
      - weights: all-gather of the updated shard                      -> `all_gather_params_`
  * collectives are torch.distributed calls (NCCL over NVLink on GPUs; gloo in the CPU tests of the
    host-side logic) issued on the current stream, so they stay ordered with the kernels.

The pure functions at the top (histogram scan, tie bases) restate on the host what the device
scan kernel computes; they exist so the cross-rank logic can be tested on CPU with gloo.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import capi
from .engine import HotPath, OptConfig
from .flat import shard_bounds


# ---- host restatement of the select scan (CPU-testable) -----------------------------------------
def scan_from_top(bins: torch.Tensor, want: int) -> Tuple[int, int]:
    """Largest bin B with  above(B) < want <= above(B) + bins[B]; returns (B, above(B)) or (-1, 0)."""
    if want <= 0:
        return -1, 0
    rev = torch.flip(bins.to(torch.int64), dims=[0])
    incl = torch.cumsum(rev, 0)
    hit = torch.nonzero(incl >= want)
    if hit.numel() == 0:
        return -1, 0
    j = int(hit[0])
    b = bins.numel() - 1 - j
    above = int(incl[j] - rev[j])
    return b, above


def tie_bases(local_eq_counts: Sequence[int]) -> List[int]:
    """Exclusive prefix over ranks of the per-rank number of threshold-equal keys."""
    out, run = [], 0
    for c in local_eq_counts:
        out.append(run)
        run += int(c)
    return out


class ShardGroup:
    """A process group plus the shard arithmetic of one flat vector.

    padded_len: length of the (padded) flat buffers, a multiple of world x align.  When given, shard r
    is the r-th EQUAL slice of the padded buffer clipped to [0, n_total): reduce-scatter and all-gather
    then work in place on equal chunks, while kernels only ever see the valid elements.

    parts > 1 (needs padded_len, a multiple of world x align x parts): the vector is cut into `parts` equal
    consecutive pieces and EVERY piece is split over the ranks, so a rank owns `parts` spans — its local buffers
    are their concatenation.  The late pieces of a model's flat vector receive their gradients first during
    backward, so their exchange can run while backward is still producing the early ones (OverlappedBackward).
    `spans` = [(first global element, offset in the local buffers, elements)], one per part.
    """

    def __init__(self, n_total: int, group=None, align: int = 16, padded_len: Optional[int] = None, parts: int = 1):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_total = int(n_total)
        self.align = align
        self.per = None
        if padded_len is None:
            self.bounds = [shard_bounds(self.n_total, self.world, r, align) for r in range(self.world)]
        else:
            if padded_len < self.n_total or padded_len % (self.world * align):
                raise ValueError("padded_len must be >= n_total and a multiple of world * align")
            self.per = padded_len // self.world
            self.bounds = [(min(r * self.per, self.n_total), min((r + 1) * self.per, self.n_total))
                           for r in range(self.world)]
        self.lo, self.hi = self.bounds[self.rank]
        self.parts = int(parts)
        if self.parts == 1:
            self.spans = [(self.lo, 0, self.hi - self.lo)]
        else:
            if padded_len is None or padded_len % (self.world * align * self.parts):
                raise ValueError("parts > 1 needs padded_len, a multiple of world * align * parts")
            piece = padded_len // self.parts
            per = piece // self.world
            self.spans, off = [], 0
            for k in range(self.parts):
                glo = min(k * piece + self.rank * per, self.n_total)
                ghi = min(k * piece + (self.rank + 1) * per, self.n_total)
                self.spans.append((glo, off, ghi - glo))
                off += ghi - glo
            self.lo = self.hi = None               # no single contiguous shard: the NCCL helpers do not apply
            self.per = None
        self._nccl = dist.get_backend(group) == "nccl"

    @staticmethod
    def pad_multiple(world: int, align: int = 16) -> int:
        return world * align

    @property
    def n_local(self) -> int:
        return sum(c for _, _, c in self.spans)

    def local(self, flat: torch.Tensor) -> torch.Tensor:
        """This rank's elements of a full vector: a view for one span, the concatenation of the spans otherwise."""
        if self.parts == 1:
            return flat[self.lo:self.hi]
        return torch.cat([flat[g:g + c] for g, _, c in self.spans])

    def all_reduce_(self, t: torch.Tensor) -> None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def reduce_gradients_(self, g_full: torch.Tensor, average: bool = True) -> torch.Tensor:
        """Combine the per-rank gradients of a data-parallel backward pass; returns this rank's shard.
        (all-reduce: every rank keeps the full summed gradient, as DataParallel's reduce_add leaves on
        GPU 0; the sharded update only reads its slice.)"""
        if average and self._nccl:
            dist.all_reduce(g_full, op=dist.ReduceOp.AVG, group=self.group)
        else:
            self.all_reduce_(g_full)
            if average:
                g_full.div_(self.world)
        return self.local(g_full)

    def reduce_scatter_gradients_(self, g_padded: torch.Tensor, average: bool = True) -> torch.Tensor:
        """Sharded update only needs this rank's slice of the reduced gradient: reduce-scatter moves
        half the bytes of an all-reduce.  In place on the padded gradient buffer (equal chunks)."""
        if self.per is None:
            raise RuntimeError("reduce_scatter needs a ShardGroup built with padded_len")
        out = g_padded[self.rank * self.per:(self.rank + 1) * self.per]
        op = dist.ReduceOp.AVG if (average and self._nccl) else dist.ReduceOp.SUM
        dist.reduce_scatter_tensor(out, g_padded, op=op, group=self.group)
        if average and not self._nccl:
            out.div_(self.world)
        return out[:self.n_local]

    def all_gather_params_(self, p_full: torch.Tensor) -> None:
        """Every rank updated its shard of p_full; make the whole vector consistent again."""
        if self.per is not None and p_full.numel() == self.per * self.world:
            dist.all_gather_into_tensor(p_full, p_full[self.rank * self.per:(self.rank + 1) * self.per],
                                        group=self.group)
            return
        shards = [p_full[lo:hi] for lo, hi in self.bounds]
        if len({s.numel() for s in shards}) == 1 and p_full.is_cuda:
            dist.all_gather_into_tensor(p_full[:shards[0].numel() * self.world], shards[self.rank], group=self.group)
        else:
            # uneven shards (ragged tail): one broadcast per shard, in place
            for r, s in enumerate(shards):
                if s.numel():
                    dist.broadcast(s, src=dist.get_global_rank(self.group, r) if self.group else r, group=self.group)


class BucketedGradReducer:
    """Overlaps the data-parallel gradient exchange with the backward pass.

    `FlatParams` makes every `param.grad` a view of the flat gradient `g`, and autograd produces gradients roughly in
    reverse parameter order, so the tail of `g` is final long before the head.  The flat vector is cut into contiguous
    buckets (whole parameters, ~`bucket_bytes` each); a post-accumulate-grad hook counts the parameters of each bucket
    and, when a bucket is complete, starts its all-reduce asynchronously (NCCL runs it on its own stream while the rest
    of the backward pass keeps the SMs busy).  `finish()` waits for every bucket and returns this rank's shard, exactly
    what `ShardGroup.reduce_gradients_` returns after a monolithic all-reduce (same sums: every element is reduced once).

    One backward pass per `finish()`; parameters that receive no gradient in a pass (their hook never fires) are
    reduced by `finish()` with whatever their bucket holds (zeros after `zero_grad`).
    """

    def __init__(self, flat, shards: ShardGroup, bucket_bytes: int = 64 << 20, average: bool = True):
        self.flat, self.shards, self.average = flat, shards, average
        elem = flat.g.element_size()
        self.buckets: List[Tuple[int, int]] = []            # [lo, hi) element ranges, in flat order
        self._bucket_of: List[int] = []                      # per trainable parameter (layout order)
        lo = 0
        for seg in flat.layout:
            end = seg.offset + seg.numel
            self._bucket_of.append(len(self.buckets))
            if (end - lo) * elem >= bucket_bytes:
                self.buckets.append((lo, end))
                lo = end
        if lo < flat.n or not self.buckets:
            self.buckets.append((lo, flat.n))
        self._bucket_of = [min(b, len(self.buckets) - 1) for b in self._bucket_of]
        self._need = [0] * len(self.buckets)
        for b in self._bucket_of:
            self._need[b] += 1
        self._seen = [0] * len(self.buckets)
        self._work: List[Optional[object]] = [None] * len(self.buckets)
        self._handles = [prm.register_post_accumulate_grad_hook(self._make_hook(b))
                         for prm, b in zip(flat._train_params, self._bucket_of)]

    def _make_hook(self, bucket: int):
        def hook(_param):
            self._seen[bucket] += 1
            if self._seen[bucket] == self._need[bucket]:
                self._launch(bucket)
        return hook

    def _launch(self, bucket: int) -> None:
        lo, hi = self.buckets[bucket]
        view = self.flat.g[lo:hi]
        if self.average and self.shards._nccl:
            self._work[bucket] = dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.shards.group, async_op=True)
        else:
            self._work[bucket] = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.shards.group, async_op=True)

    def finish(self) -> torch.Tensor:
        """Call after `loss.backward()`: completes the exchange and returns this rank's gradient shard."""
        for b in range(len(self.buckets)):
            if self._work[b] is None:                        # a parameter of this bucket got no gradient this pass
                self._launch(b)
        for b, w in enumerate(self._work):
            w.wait()
            self._work[b] = None
            self._seen[b] = 0
        if self.average and not self.shards._nccl:
            self.flat.g.div_(self.shards.world)
        return self.shards.local(self.flat.g)

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


# byte offset of sfr_select_state.thr_key: 5 x u64 (k, k_in_bin, count_gt, count_eq, tie_budget) + u32 prefix
_THR_KEY_I32_INDEX = capi.SelectState.thr_key.offset // 4


class ShardedHotPath(HotPath):
    """HotPath over this rank's shard, with the cross-rank reductions filled in."""

    def __init__(self, shards: ShardGroup, device, opt: OptConfig, **kw):
        super().__init__(shards.n_local, device, opt, **kw)
        self.shards = shards
        self._local_bins: Optional[torch.Tensor] = None
        self.select_graphs = False                 # the select's cross-rank reductions are torch.distributed calls
        self.xchg: Optional["PeerExchange"] = None
        self._g_red: dict = {}
        self._pending: Optional[dict] = None
        self._sumsq_late: Optional[torch.Tensor] = None

    def reduce_scalar_(self, t: torch.Tensor) -> None:
        self.shards.all_reduce_(t)

    def total_elements(self) -> int:
        return self.shards.n_total

    def reduce_bins_(self, bins: torch.Tensor, count: int) -> None:
        self.shards.all_reduce_(bins[:count])

    def keep_local_bins_(self, bins: torch.Tensor) -> Optional[torch.Tensor]:
        if self._local_bins is None:                                  # persistent: no allocation per select
            self._local_bins = torch.empty(capi.SELECT_BINS1, dtype=bins.dtype, device=bins.device)
        self._local_bins.copy_(bins[:capi.SELECT_BINS1])
        return self._local_bins

    def tie_base_(self, state: torch.Tensor, local_bins: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """Number of threshold-equal keys on lower ranks, entirely on the device (no host read of the select
        state): the threshold key sits at a fixed offset of the device struct (sfr_select_state.thr_key)."""
        thr_key = state.view(torch.int32)[_THR_KEY_I32_INDEX].to(torch.int64) & 0xFFFFFFFF
        mine = local_bins.index_select(0, (thr_key & 0xFFFF).reshape(1))          # local #keys == threshold
        gathered = torch.empty(self.shards.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(gathered, mine, group=self.shards.group)
        # (for select_all / select_none the value is unused: the apply kernels do not order ties then)
        return gathered[:self.shards.rank].sum().reshape(1)

    def ratio_mask(self, threshold: float, **kw) -> torch.Tensor:
        mask = super().ratio_mask(threshold, **kw)
        self.shards.all_reduce_(self.zero_count)
        return mask

    def ratio_masks(self, thresholds, **kw) -> torch.Tensor:
        masks = super().ratio_masks(thresholds, **kw)
        self.shards.all_reduce_(self.zero_count)
        return masks

    # ---- data-parallel steps over NVLink peer memory (csrc/peer.cu) --------------------------------------
    # The reference's DataParallel loops (DiT/forget.py:193,285-322; DiT/generate_fisher.py:173,216-291;
    # DDPM/runners/diffusion.py:1060,1126-1180,1270-1281) reduce every GPU's gradient onto GPU 0, update
    # there, and re-broadcast.  Here each rank pulls ITS shard of the summed gradient straight out of the
    # peers' gradient buffers inside the kernel that consumes it, and pushes its updated weights into
    # every rank's weight buffer from the kernel that produced them.
    def attach_exchange(self, xchg: "PeerExchange") -> None:
        if xchg.shards is not self.shards:
            raise ValueError("the exchange and the hot path must share one ShardGroup")
        self.xchg = xchg

    def _need_xchg(self) -> "PeerExchange":
        if self.xchg is None:
            raise capi.SfrError(capi.ERR_ARG, "ShardedHotPath", "no PeerExchange attached (attach_exchange)")
        return self.xchg

    def reduced(self, slot: str) -> torch.Tensor:
        """Local fp32 buffer for a reduced gradient shard (one per slot name, allocated once)."""
        t = self._g_red.get(slot)
        if t is None:
            t = torch.zeros(self.n, dtype=torch.float32, device=self.device)
            self._g_red[slot] = t
        return t

    # ---- span-level building blocks (one kernel launch each; no barrier) ----------------------------------
    def _span(self, t: Optional[torch.Tensor], i: int) -> Optional[torch.Tensor]:
        """Span i of a tensor that covers this rank's whole local shard."""
        if t is None:
            return None
        _, off, cnt = self.shards.spans[i]
        return t if (t.numel() == cnt and self.shards.parts == 1) else t[off:off + cnt]

    def _p_span(self, p, i: int) -> torch.Tensor:
        """p: the local fp32 shard (all spans concatenated), or one tensor per span (views of a full weight vector)."""
        return p[i] if isinstance(p, (list, tuple)) else self._span(p, i)

    def reduce_span(self, x: "PeerExchange", i: int, g: "SymBuffer", *, average: bool = True,
                    g_red: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
                    sumsq: Optional[torch.Tensor] = None, fisher: Optional[torch.Tensor] = None,
                    fisher_divisor: float = 1.0, max_ctas: int = 0, xp: Optional[int] = None) -> None:
        capi.peer_reduce(g.buf, g.tensor.dtype, x.span_geom(i), x.reduce_transport if xp is None else xp, average,
                         g_red=self._span(g_red, i), mask=self._span(mask, i), sumsq=sumsq,
                         fisher=self._span(fisher, i), fisher_divisor=fisher_divisor, max_ctas=max_ctas)

    def update_span(self, x: "PeerExchange", i: int, p, a: "capi.UpdateArgs", *, g, g_red, weights, weights_bf16, mask,
                    use_ema: bool, clip, average: bool = True, consts: str = "auto", max_ctas: int = 0,
                    xp: Optional[int] = None) -> None:
        """K3 (+ exchange) on span i.  consts: how the step's scalars (step counter, bias corrections, clip
        coefficient) reach the kernel —
          "auto"     one launch per optimizer step: by value unless a clip / device counter / device LR needs the
                     scalar-prep kernel;
          "prepare"  first of several launches of ONE optimizer step: always run the prep kernel (it leaves the
                     scalars in device scratch and advances the device step counter once);
          "reuse"    a later launch of the same step: read that scratch, advance nothing."""
        sgd = self.opt.kind == "sgd"
        prepare = consts != "reuse"
        need_consts = (consts != "auto" or clip is not None or self.step_dev is not None
                       or self.lr_table is not None)
        flags = a.flags
        if not prepare:
            a.flags = flags | capi.F_REUSE_CONSTS
        pushes = weights is not None or weights_bf16 is not None
        g_xp, bc_xp = x.reduce_transport, x.push_transport
        if g_red is None and pushes:
            g_xp = bc_xp = x.fused_transport              # one kernel, both directions
        if xp is not None:
            g_xp = bc_xp = xp
        try:
            capi.peer_fused_update(
                self._p_span(p, i), x.span_geom(i), a, g_red=self._span(g_red, i), g=g.buf if g_red is None else None,
                g_dtype=g.tensor.dtype if g_red is None else torch.float32, g_transport=g_xp,
                average=average, m=None if (sgd and self.opt.momentum == 0.0) else self._span(self.m, i),
                v=None if sgd else self._span(self.v, i), mask=self._span(mask, i),
                ema=self._span(self.slow, i) if use_ema else None, bc_f32=None if weights is None else weights.buf,
                bc_bf16=None if weights_bf16 is None else weights_bf16.buf, bc_transport=bc_xp,
                clip_sumsq=clip, step_counter=self.step_dev if prepare else None,
                consts_scratch=self._consts_dev if need_consts else None, max_ctas=max_ctas)
        finally:
            a.flags = flags

    def dp_fisher_accumulate(self, which: str, g: "SymBuffer", divisor: float, *, average: bool = True,
                             keep: Optional[str] = None, clip_max_norm: Optional[float] = None) -> None:
        """K1 on the data-parallel mean gradient: barrier, then ONE kernel (per span) that pulls this shard of every
        rank's `g`, averages and accumulates F += gbar**2 / divisor (keep="slot": the reduced shard is also left in
        `self.reduced(slot)` for a following step on the same gradients).  With `clip_max_norm` (DDPM Fisher of
        the clipped gradient) the norm has to be global first: reduce + sum of squares, summed across ranks on
        the barrier, then the shard-local clipped K1."""
        x = self._need_xchg()
        acc = self.buffer({"forget": "forget_fisher", "remain": "remain_fisher"}.get(which, which))
        spans = range(len(self.shards.spans))
        x.barrier()
        if clip_max_norm is None:
            for i in spans:
                self.reduce_span(x, i, g, average=average, fisher=acc, fisher_divisor=divisor,
                                 g_red=self.reduced(keep) if keep else None)
            self._t("fisher_accum")
            x.barrier()
            return
        red = self.reduced(keep or "_clip")
        self.sumsq.zero_()
        for i in spans:
            self.reduce_span(x, i, g, average=average, g_red=red, sumsq=self.sumsq)
        x.barrier(self.sumsq, self.sumsq)
        capi.fisher_accum(acc, red, divisor, clip_sumsq=self.sumsq, clip_max_norm=clip_max_norm)
        self._t("fisher_accum_clipped")

    def _step_flags(self, mask, mask_order: str) -> int:
        flags = 0
        if mask is not None:
            flags |= capi.F_MASK if mask_order == "mask_then_clip" else capi.F_MASK_AFTER_CLIP
        if self.step_dev is None and self.opt.kind == "sgd" and self.opt.momentum != 0.0 and not self.has("m"):
            flags |= capi.F_SGD_FIRST_STEP
        return flags

    def dp_step(self, p, g, *, weights: Optional["SymBuffer"] = None,
                weights_bf16: Optional["SymBuffer"] = None, mask: Optional[torch.Tensor] = None,
                mask_order: str = "mask_then_clip", max_norm: Optional[float] = None, ema: bool = False,
                lr: Optional[float] = None, average: bool = True, keep: str = "_step") -> None:
        """One optimizer step of the sharded data-parallel loop.

        p: this rank's fp32 weight shard (a view of `weights.tensor[lo:hi]`, the fp32 master shard in bf16 mode, or
           one tensor per span).
        g: a SymBuffer holding every rank's FULL gradient (reduced inside the kernels), or a local fp32
           tensor with this rank's already reduced shard (e.g. `self.reduced(slot)` left by dp_fisher_accumulate).
        weights / weights_bf16: full-vector symmetric buffers that receive the updated shard on EVERY rank.
        No clip: barrier -> reduce + K3 + push in one kernel -> barrier.
        Clip:    barrier -> reduce + masked sum of squares -> barrier carrying the norm -> K3 + push -> barrier."""
        x = self._need_xchg()
        from_peers = isinstance(g, SymBuffer)
        flags = self._step_flags(mask, mask_order)
        spans = range(len(self.shards.spans))
        clip = None
        g_red = None if from_peers else g
        if from_peers:
            x.barrier()                                   # every rank's backward has written its gradient
        if max_norm is not None:
            norm_mask = mask if (mask is not None and mask_order == "mask_then_clip") else None
            self.sumsq.zero_()
            if from_peers:
                g_red = self.reduced(keep)
                for i in spans:
                    self.reduce_span(x, i, g, average=average, g_red=g_red, mask=norm_mask, sumsq=self.sumsq)
            else:
                capi.masked_sumsq(g_red, norm_mask, self.sumsq)
            self._t("masked_sumsq")
            x.barrier(self.sumsq, self.sumsq)             # the clip norm's all-reduce rides on the barrier
            clip = self.sumsq
        self.step_count += 1
        a = self._args(flags, ema, max_norm, lr)
        use_ema = ema and self.ema_mode != "none"
        several = len(spans) > 1
        for k, i in enumerate(spans):
            self.update_span(x, i, p, a, g=g, g_red=g_red, weights=weights, weights_bf16=weights_bf16, mask=mask,
                             use_ema=use_ema, clip=clip, average=average,
                             consts="auto" if not several else ("prepare" if k == 0 else "reuse"))
        self._t("fused_update_ema" if use_ema else "fused_update")
        x.barrier()                                       # gradients may be overwritten; every weight store has landed

    # ---- the same step, pipelined against the backward pass (ShardGroup(parts=2) + OverlappedBackward) ----------
    def dp_begin_step(self, ov: "OverlappedBackward", p, g: "SymBuffer", *, weights: Optional["SymBuffer"] = None,
                      weights_bf16: Optional["SymBuffer"] = None, mask: Optional[torch.Tensor] = None,
                      mask_order: str = "mask_then_clip", max_norm: Optional[float] = None, ema: bool = False,
                      lr: Optional[float] = None, average: bool = True, keep: str = "_step") -> None:
        """Call BEFORE `loss.backward()`.  The flat vector is cut into `parts` pieces; backward finishes the LAST piece
        first.  As soon as a piece other than the first is final on this rank, this rank's span of it is exchanged
        on a side stream, by a few CTAs, while backward keeps producing the earlier pieces:
          unclipped step   barrier -> reduce + K3 [+ EMA] + weight push of the span (one kernel)
          clipped step     barrier -> reduce + masked sum of squares of the span (the update needs the whole norm)
        `dp_finish_step()` after backward does the first piece and whatever had to wait for the norm."""
        if self.shards.parts < 2:
            raise capi.SfrError(capi.ERR_ARG, "dp_begin_step", "needs a ShardGroup with parts >= 2")
        self._need_xchg()
        rec = dict(ov=ov, p=p, g=g, weights=weights, weights_bf16=weights_bf16, mask=mask, mask_order=mask_order,
                   max_norm=max_norm, ema=ema, lr=lr, average=average, keep=keep, a=None)
        self._pending = rec
        if self._sumsq_late is None:
            self._sumsq_late = torch.zeros(1, dtype=torch.float64, device=self.device)
        use_ema = ema and self.ema_mode != "none"
        norm_mask = mask if (mask is not None and mask_order == "mask_then_clip") else None
        last = self.shards.parts - 1

        def late(part: int):
            xs = ov.xchg
            xs.barrier()                                  # every rank's backward is past this piece
            if max_norm is not None:
                if part == last:
                    self._sumsq_late.zero_()
                self.reduce_span(xs, part, g, average=average, g_red=self.reduced(keep), mask=norm_mask,
                                 sumsq=self._sumsq_late, max_ctas=ov.max_ctas, xp=capi.XP_TMA)
            else:
                if part == last:                          # the first piece to go prepares the step's scalars
                    self.step_count += 1
                    rec["a"] = self._args(self._step_flags(mask, mask_order), ema, None, lr)
                self.update_span(xs, part, p, rec["a"], g=g, g_red=None, weights=weights, weights_bf16=weights_bf16,
                                 mask=mask, use_ema=use_ema, clip=None, average=average,
                                 consts="prepare" if part == last else "reuse", max_ctas=ov.max_ctas,
                                 xp=capi.XP_TMA)          # few CTAs beside backward: the copy-engine transport

        ov.arm(late)

    def dp_finish_step(self) -> None:
        """Call AFTER `loss.backward()`: the first piece, the norm (clipped steps), the join with the side stream."""
        rec, self._pending = self._pending, None
        if rec is None:
            raise capi.SfrError(capi.ERR_ARG, "dp_finish_step", "no step in flight (dp_begin_step)")
        x, ov = self._need_xchg(), rec["ov"]
        ov.flush()                                        # (a pass in which some late parameter got no gradient)
        g, p, mask, max_norm = rec["g"], rec["p"], rec["mask"], rec["max_norm"]
        use_ema = rec["ema"] and self.ema_mode != "none"
        x.barrier()                                       # every rank's backward has written ALL its gradients
        if max_norm is not None:
            norm_mask = mask if (mask is not None and rec["mask_order"] == "mask_then_clip") else None
            g_red = self.reduced(rec["keep"])
            self.sumsq.zero_()
            self.reduce_span(x, 0, g, average=rec["average"], g_red=g_red, mask=norm_mask, sumsq=self.sumsq)
            ov.join()
            self.sumsq.add_(self._sumsq_late)
            self._t("masked_sumsq")
            x.barrier(self.sumsq, self.sumsq)
            self.step_count += 1
            a = self._args(self._step_flags(mask, rec["mask_order"]), rec["ema"], max_norm, rec["lr"])
            for i in range(self.shards.parts):
                self.update_span(x, i, p, a, g=g, g_red=g_red, weights=rec["weights"],
                                 weights_bf16=rec["weights_bf16"], mask=mask, use_ema=use_ema, clip=self.sumsq,
                                 average=rec["average"], consts="prepare" if i == 0 else "reuse")
        else:
            ov.join()                                     # the last piece's prep kernel has set this step's scalars
            self.update_span(x, 0, p, rec["a"], g=g, g_red=None, weights=rec["weights"],
                             weights_bf16=rec["weights_bf16"], mask=mask, use_ema=use_ema, clip=None,
                             average=rec["average"], consts="reuse")
        self._t("fused_update_ema" if use_ema else "fused_update")
        x.barrier()

    def dp_forget_step(self, p, g, *, mask: Optional[torch.Tensor] = None, use_mask: bool = True, **kw) -> None:
        """grad *= mask ; clip ; step on the data-parallel mean gradient (DiT/forget.py:285-299 under DataParallel)."""
        if mask is None and use_mask:
            mask = self.require_mask()
        self.dp_step(p, g, mask=mask if use_mask else None, ema=False, **kw)

    def dp_remain_step(self, p, g, *, ema: bool = True, **kw) -> None:
        """[clip ;] step ; EMA on the data-parallel mean gradient (DiT/forget.py:310-322 under DataParallel)."""
        self.dp_step(p, g, mask=None, ema=ema, **kw)


class OverlappedBackward:
    """Fires a callback INSIDE `loss.backward()`, on a side stream, each time one more piece of the flat gradient is
    final.  The vector is cut into `shards.parts` equal pieces; autograd produces gradients roughly in reverse
    parameter order, so the pieces complete last-to-first, and a piece is final once every parameter that reaches
    into it has accumulated its gradient.  The callback (armed per backward pass by ShardedHotPath.dp_begin_step)
    launches the exchange of that piece with few CTAs (`max_ctas`), so it runs beside the rest of the backward
    pass instead of after it; piece 0 is left to `dp_finish_step`.  `xchg` is the exchange the side stream uses (its
    own barrier pad: PeerExchange.sibling()).  Works under CUDA-graph capture: the side stream forks from and
    joins the capturing stream."""

    def __init__(self, flat, shards: ShardGroup, xchg: "PeerExchange", max_ctas: int = 32):
        if shards.parts < 2:
            raise ValueError("OverlappedBackward pipelines parts >= 2: ShardGroup(..., parts=P)")
        self.xchg, self.max_ctas, self.parts = xchg, int(max_ctas), shards.parts
        self.device = flat.device
        self.side = torch.cuda.Stream(device=self.device)
        piece = flat.n_padded // shards.parts              # the same cut on every rank
        self._parts_of = {}                                # parameter -> the late pieces it reaches into
        self._need = [0] * shards.parts
        for seg, prm in zip(flat.layout, flat._train_params):
            first, last = seg.offset // piece, (seg.offset + seg.numel - 1) // piece
            mine = [k for k in range(max(first, 1), min(last, shards.parts - 1) + 1)]
            if mine:
                self._parts_of[prm] = mine
                for k in mine:
                    self._need[k] += 1
        self._seen = [0] * shards.parts
        self._next = shards.parts - 1                      # pieces fire in order P-1, P-2, ..., 1
        self._fn = None
        self._handles = [prm.register_post_accumulate_grad_hook(self._hook) for prm in self._parts_of]

    def arm(self, fn) -> None:
        self._fn, self._seen, self._next = fn, [0] * self.parts, self.parts - 1

    def _hook(self, param) -> None:
        if self._fn is None:
            return
        for k in self._parts_of[param]:
            self._seen[k] += 1
        while self._next >= 1 and self._seen[self._next] >= self._need[self._next]:
            self._fire(self._next)

    def _fire(self, part: int) -> None:
        self._next = part - 1
        main = torch.cuda.current_stream(self.device)
        self.side.wait_stream(main)                        # this piece's gradients precede this point of the backward pass
        with torch.cuda.stream(self.side):
            self._fn(part)

    def flush(self) -> None:
        """After backward: pieces whose hook count never completed (a parameter without a gradient this pass)."""
        while self._fn is not None and self._next >= 1:
            self._fire(self._next)

    def join(self) -> None:
        torch.cuda.current_stream(self.device).wait_stream(self.side)
        self._fn = None

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


class SymBuffer:
    """One symmetric buffer: the local tensor plus the addresses every rank's copy has in THIS process."""

    def __init__(self, tensor: torch.Tensor, buf: "capi.PeerBuf", handle, world: int):
        self.tensor, self.buf, self.handle, self.world = tensor, buf, handle, world

    @property
    def has_multicast(self) -> bool:
        return bool(self.buf.multicast)


class PeerExchange:
    """NVLink peer-memory plumbing of the data-parallel hot path (one process per GPU).

    Torch's symmetric-memory allocator provides what a C caller would get from cuMem* / cuMulticast*: buffers
    that every rank of the group maps (`buffer_ptrs`) and, on an NVSwitch fabric, one multicast address per
    buffer.  Nothing else of torch is on the data path: reduction, update, broadcast and the barrier are the
    library's kernels (csrc/peer.cu).

    transport: "auto", "p2p", "multimem", or "<reduce>+<push>" (e.g. "p2p+multimem") to choose separately how
    gradients are pulled and how weights are pushed.
    """

    def __init__(self, shards: ShardGroup, device, transport: str = "auto", timeout_s: float = 30.0):
        if shards.world > capi.MAX_PEERS:
            raise capi.SfrError(capi.ERR_ARG, "PeerExchange", f"at most {capi.MAX_PEERS} ranks (one NVSwitch box)")
        if any(c and g % 16 for g, _, c in shards.spans):
            raise capi.SfrError(capi.ERR_ALIGN, "PeerExchange", "shard starts must be multiples of 16 elements")
        self.shards = shards
        self.device = torch.device(device)
        self.group = shards.group if shards.group is not None else dist.group.WORLD
        self.timeout_ns = int(timeout_s * 1e9)
        self._want = transport
        # what "auto" means (measured, tools/xchg_bench.py -> profiles/r2_xchg_n*.jsonl): TMA bulk copies match or beat
        # the load/store transports and NCCL for every op at every world size tried; the ONE-kernel reduce + update +
        # push is the exception at 8 ranks — there the switch's fan-in / fan-out (multimem) halves the bytes per
        # direction (fp32: 5.85 vs 7.07 ms; at 4 ranks TMA still wins, 5.96 vs 6.21 ms)
        self._auto = (capi.XP_TMA, capi.XP_TMA)
        self._auto_fused = capi.XP_MULTIMEM if shards.world >= 8 else capi.XP_TMA
        self._buffers: List[SymBuffer] = []
        self._geoms = [capi.PeerGeom(shards.world, shards.rank, g, c) for g, _, c in shards.spans]
        self.geom = self._geoms[0]                          # (the only one unless the ShardGroup has parts > 1)
        self.pad = self.alloc(capi.peer_pad_bytes() // 8, torch.int64)         # zero-filled by alloc()
        self._vals = torch.zeros(8, dtype=torch.float64, device=self.device)
        self.barriers = 0
        torch.cuda.synchronize(self.device)
        dist.barrier(group=shards.group)                   # every rank's pad is zero before the first epoch

    # transport = "<reduce>" or "<reduce>+<push>" with each of auto | p2p | tma | multimem, e.g. "p2p+multimem":
    # gradients pulled through the mapped pointers, weights pushed through the multicast address.  "tma" moves
    # both directions with TMA bulk copies through shared memory (csrc/peer_tma.cu) and cannot be mixed inside one
    # fused kernel.
    def _pick(self, which: int) -> int:
        want = self._want.split("+")
        want = want[which] if len(want) > 1 else want[0]
        if want == "p2p":
            return capi.XP_P2P
        if want == "tma":
            return capi.XP_TMA
        data = self._buffers[1:] or self._buffers
        mc = all(b.has_multicast for b in data)
        if want == "multimem":
            if not mc:
                raise capi.SfrError(capi.ERR_ARG, "PeerExchange", "multimem requested but a buffer has no multicast address")
            return capi.XP_MULTIMEM
        if want != "auto":
            raise capi.SfrError(capi.ERR_ARG, "PeerExchange", f"unknown transport {want!r}")
        if self._auto[which] == capi.XP_MULTIMEM and not mc:
            return capi.XP_TMA
        return self._auto[which]

    @property
    def fused_transport(self) -> int:
        """Transport of the kernel that reduces, updates and pushes in one pass (both directions at once)."""
        if self._want != "auto":
            r, q = self._pick(0), self._pick(1)
            return r if r == q else capi.XP_TMA
        data = self._buffers[1:] or self._buffers
        if self._auto_fused == capi.XP_MULTIMEM and not all(b.has_multicast for b in data):
            return capi.XP_TMA
        return self._auto_fused

    @property
    def reduce_transport(self) -> int:
        return self._pick(0)

    @property
    def push_transport(self) -> int:
        return self._pick(1)

    @property
    def transport_name(self) -> str:
        names = {capi.XP_P2P: "p2p", capi.XP_MULTIMEM: "multimem", capi.XP_TMA: "tma"}
        r, p = names[self.reduce_transport], names[self.push_transport]
        return r if r == p else f"{r}+{p}"

    def span_geom(self, i: int) -> "capi.PeerGeom":
        return self._geoms[i]

    def sibling(self) -> "PeerExchange":
        """A second exchange over the same shards with its OWN barrier pad, for kernels on another stream
        (a pad serves one stream at a time).  Collective."""
        other = PeerExchange(self.shards, self.device, transport=self._want, timeout_s=self.timeout_ns / 1e9)
        other._auto = self._auto
        other._buffers.extend(self._buffers[1:])            # same data buffers: "auto" sees the same multicast support
        return other

    def alloc(self, numel: int, dtype: torch.dtype) -> SymBuffer:
        """Collective: every rank allocates the same buffer (zero-filled) and maps the others'."""
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(int(numel), dtype=dtype, device=self.device)
        h = symm.rendezvous(t, self.group)
        t.zero_()
        ptrs = [int(p) for p in h.buffer_ptrs]
        if ptrs[self.shards.rank] != t.data_ptr():
            raise capi.SfrError(capi.ERR_ARG, "PeerExchange.alloc", "symmetric buffer is not at the start of its allocation")
        b = SymBuffer(t, capi.peer_buf(ptrs, int(getattr(h, "multicast_ptr", 0) or 0)), h, self.shards.world)
        self._buffers.append(b)
        return b

    def barrier(self, vals: Optional[torch.Tensor] = None, sums: Optional[torch.Tensor] = None) -> None:
        """Cross-GPU barrier on the current stream (a kernel, not a host wait).  `vals` (device float64, <= 8):
        their sum over ranks, in rank order, is left in `sums` on every rank."""
        capi.peer_barrier(self.pad.buf, self.shards.world, self.shards.rank, vals, sums, self.timeout_ns)
        self.barriers += 1

    def check(self) -> None:
        """Raise if a barrier ever timed out on this rank (synchronises the device)."""
        if int(self.pad.tensor[1].item()) != 0:
            raise RuntimeError("sfr_peer_barrier timed out: a peer rank stopped participating")

    def all_gather_(self, shard: torch.Tensor, full: SymBuffer) -> None:
        """Push this rank's shard into every rank's copy of `full` (the all-gather alone), between two barriers."""
        capi.peer_broadcast(shard, full.buf, self.geom, self.push_transport)
