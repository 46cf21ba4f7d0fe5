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
    """

    def __init__(self, n_total: int, group=None, align: int = 16, padded_len: Optional[int] = None):
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
        self._nccl = dist.get_backend(group) == "nccl"

    @staticmethod
    def pad_multiple(world: int, align: int = 16) -> int:
        return world * align

    @property
    def n_local(self) -> int:
        return self.hi - self.lo

    def local(self, flat: torch.Tensor) -> torch.Tensor:
        return flat[self.lo:self.hi]

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

    def reduce_scalar_(self, t: torch.Tensor) -> None:
        self.shards.all_reduce_(t)

    def total_elements(self) -> int:
        return self.shards.n_total

    def reduce_bins_(self, bins: torch.Tensor, count: int) -> None:
        self.shards.all_reduce_(bins[:count])

    def keep_local_bins_(self, bins: torch.Tensor) -> Optional[torch.Tensor]:
        return bins.clone()

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
