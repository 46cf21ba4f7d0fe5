"""Flat parameter-vector layout: the HBM data layout of the SFR-on hot path.

The reference walks `model.named_parameters()` in Python for every stage (Fisher, mask,
mask-apply, clip, step, EMA: SURVEY.md §3.5).  Here every role of the path — weights `p`,
gradients `g`, optimizer state `m`/`v`, slow/EMA weights, the two Fisher accumulators and the
saliency mask — is ONE contiguous device buffer over the same element order, so each stage is
one kernel launch over `n` elements.

Order and packing: trainable parameters (those with requires_grad) in `named_parameters()`
order, TIGHTLY packed — flat index == index in the reference's
`torch.cat([t.flatten() for t in gradients.values()])` (DDPM/runners/diffusion.py:1009-1012),
which is what makes top-k tie-breaking by lowest flat index equal to a stable argsort of the
reference's vector.  There is no per-tensor padding, hence nothing to exclude from norms, counts
or top-k.  Tensor starts are 16-byte aligned whenever all preceding sizes are multiples of 4,
which holds for every tensor of the four reference models except a trailing 3- or 10-element
bias (verified by instantiating them).  Frozen parameters (DiT `pos_embed`) are kept
in a separate small segment list: only the reference's EMA loop touches them.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch


@dataclass(frozen=True)
class Segment:
    name: str
    shape: Tuple[int, ...]
    offset: int
    numel: int


class FlatLayout:
    """Name <-> offset table of a flat vector.  Pure host logic (no device, no kernels)."""

    def __init__(self, named_shapes: Iterable[Tuple[str, Sequence[int]]]):
        self.segments: List[Segment] = []
        off = 0
        for name, shape in named_shapes:
            shape = tuple(int(s) for s in shape)
            numel = 1
            for s in shape:
                numel *= s
            self.segments.append(Segment(name, shape, off, numel))
            off += numel
        self.numel = off
        self._by_name = {s.name: s for s in self.segments}
        if len(self._by_name) != len(self.segments):
            raise ValueError("duplicate parameter names in layout")

    @classmethod
    def from_named_tensors(cls, named: Iterable[Tuple[str, torch.Tensor]]) -> "FlatLayout":
        return cls((n, tuple(t.shape)) for n, t in named)

    def __len__(self) -> int:
        return len(self.segments)

    def __iter__(self) -> Iterator[Segment]:
        return iter(self.segments)

    def __contains__(self, name: str) -> bool:
        return name in self._by_name

    def segment(self, name: str) -> Segment:
        return self._by_name[name]

    @property
    def names(self) -> List[str]:
        return [s.name for s in self.segments]

    def misaligned(self, multiple: int = 4) -> List[Segment]:
        """Segments whose start is not a multiple of `multiple` elements (16 B for fp32)."""
        return [s for s in self.segments if s.offset % multiple]

    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        s = self._by_name[name]
        return flat[s.offset:s.offset + s.numel].view(s.shape)

    def views(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        if flat.numel() != self.numel:
            raise ValueError(f"flat vector has {flat.numel()} elements, layout has {self.numel}")
        return {s.name: flat[s.offset:s.offset + s.numel].view(s.shape) for s in self.segments}

    def flatten(self, tensors: Dict[str, torch.Tensor], *, dtype=None, device=None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Pack a name->tensor dict into one flat vector (missing names are an error)."""
        if out is None:
            first = tensors[self.segments[0].name] if self.segments else torch.empty(0)
            out = torch.empty(self.numel, dtype=dtype or first.dtype, device=device or first.device)
        for s in self.segments:
            out[s.offset:s.offset + s.numel].copy_(tensors[s.name].reshape(-1))
        return out

    # ---- shard partition (multi-GPU) --------------------------------------------------------
    def shard_bounds(self, world: int, rank: int, align: int = 16) -> Tuple[int, int]:
        return shard_bounds(self.numel, world, rank, align)


def shard_bounds(n: int, world: int, rank: int, align: int = 16) -> Tuple[int, int]:
    """Contiguous, `align`-element-aligned shard [lo, hi) of an n-element vector.

    Every shard start is a multiple of `align` elements (default 16: 16 bytes even for the
    1-byte mask stream) so the kernels' 128-bit accesses stay aligned; shard sizes differ by at
    most `align`; the ragged end (n % align) lands in the last non-empty shard.  No padding elements exist: shards
    partition [0, n) exactly, so norms / counts / top-k need no exclusion logic.
    """
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    units = (n + align - 1) // align
    per, extra = divmod(units, world)
    lo_u = rank * per + min(rank, extra)
    hi_u = lo_u + per + (1 if rank < extra else 0)
    return min(lo_u * align, n), min(hi_u * align, n)


class FlatParams:
    """Device buffers of the hot path for one model, with per-name views.

    The constructor re-points every trainable `param.data` (and `param.grad`) into flat buffers, so
    autograd reads weights from and accumulates gradients into the flat vectors directly: no gather
    before, no scatter after the kernels.

    fp32 model (the reference's only mode): `p` (fp32) IS the module weights, `g` (fp32) the grads.
    bf16 model (BASELINE config 3, an extension — the reference never uses reduced precision): the
    module weights are views into `p_work` (bf16) and the grads views into `g` (bf16); `p` is the fp32
    MASTER copy the kernels update (writing the bf16 working copy back in the same pass), and the
    optimizer state / Fisher / EMA stay fp32.

    pad_multiple: allocate the flat buffers with a length rounded up to this many elements (the views
    and every kernel still cover exactly n elements) so that equal-sized shards exist for
    reduce-scatter / all-gather; the tail [n, n_padded) is never part of the logical vector.

    alloc: optional `alloc(role, numel, dtype) -> zero-filled device tensor` for the buffers other ranks must
    reach — role "g" (gradients), "p" (fp32 weights) and "p_work" (bf16 working weights) — e.g. the tensors of
    `dist.PeerExchange.alloc`, so that the data-parallel kernels read / write them in place over NVLink.
    """

    def __init__(self, model: torch.nn.Module, device=None, *, grads_as_views: bool = True,
                 grad_dtype: Optional[torch.dtype] = None, pad_multiple: int = 1,
                 alloc=None):
        named = list(model.named_parameters())
        train = [(n, p) for n, p in named if p.requires_grad]
        frozen = [(n, p) for n, p in named if not p.requires_grad]
        self.all_names = [n for n, _ in named]
        self.layout = FlatLayout.from_named_tensors(train)
        self.frozen_layout = FlatLayout.from_named_tensors(frozen)
        if device is None:
            device = train[0][1].device if train else torch.device("cuda")
        self.device = torch.device(device)
        dtypes = {p.dtype for _, p in train}
        if len(dtypes) > 1:
            raise ValueError(f"mixed parameter dtypes {dtypes}")
        self.param_dtype = dtypes.pop() if dtypes else torch.float32
        if self.param_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError(f"parameter dtype {self.param_dtype} not supported (fp32 | bf16)")
        if grad_dtype is None:
            grad_dtype = self.param_dtype
        if not grads_as_views:
            grad_dtype = torch.float32          # the gather kernel widens whatever autograd produced
        n = self.layout.numel
        self.n = n
        self.n_padded = (n + pad_multiple - 1) // pad_multiple * pad_multiple
        if alloc is None:
            def alloc(_role, numel, dtype):
                return torch.zeros(numel, dtype=dtype, device=self.device)
        self.p_padded = alloc("p", self.n_padded, torch.float32)
        self.g_padded = alloc("g", self.n_padded, grad_dtype)
        for t_ in (self.p_padded, self.g_padded):
            if t_.numel() != self.n_padded or t_.device != self.device:
                raise ValueError("alloc() must return a tensor of the requested length on the model's device")
        self.p = self.p_padded[:n]
        self.g = self.g_padded[:n]
        self.p_work_padded = (alloc("p_work", self.n_padded, torch.bfloat16)
                              if self.param_dtype == torch.bfloat16 else None)
        self.p_work = None if self.p_work_padded is None else self.p_work_padded[:n]
        weights = self.p if self.p_work is None else self.p_work        # what the module sees
        self.frozen = torch.empty(self.frozen_layout.numel, dtype=torch.float32, device=self.device)
        self.frozen_work = (torch.empty(self.frozen_layout.numel, dtype=self.param_dtype, device=self.device)
                            if self.param_dtype != torch.float32 else None)
        views_ok = grads_as_views and grad_dtype == self.param_dtype
        with torch.no_grad():
            for seg, (_, prm) in zip(self.layout, train):
                sl = slice(seg.offset, seg.offset + seg.numel)
                self.p[sl].copy_(prm.detach().reshape(-1))              # fp32 master (exact widening)
                if self.p_work is not None:
                    self.p_work[sl].copy_(prm.detach().reshape(-1))
                prm.data = weights[sl].view(seg.shape)
                if views_ok:
                    prm.grad = self.g[sl].view(seg.shape)
            for seg, (_, prm) in zip(self.frozen_layout, frozen):
                sl = slice(seg.offset, seg.offset + seg.numel)
                self.frozen[sl].copy_(prm.detach().reshape(-1))
                if self.frozen_work is not None:
                    self.frozen_work[sl].copy_(prm.detach().reshape(-1))
                    prm.data = self.frozen_work[sl].view(seg.shape)
                else:
                    prm.data = self.frozen[sl].view(seg.shape)
        self.grads_as_views = views_ok
        self._train_params = [p for _, p in train]
        self._gather_tables = None

    # ---- buffers of the other roles, allocated on demand ------------------------------------
    def new_buffer(self, dtype=torch.float32, zero: bool = True) -> torch.Tensor:
        f = torch.zeros if zero else torch.empty
        return f(self.n, dtype=dtype, device=self.device)

    def zero_grad(self) -> None:
        """One memset instead of optimizer.zero_grad()'s per-tensor loop.  (The fused update can
        also zero `g` on its way out: SFR_F_ZERO_GRAD.)"""
        self.g.zero_()

    def collect_grads(self) -> torch.Tensor:
        """Make sure `g` holds this step's gradients.  With view-grads this is free; otherwise
        the per-tensor `.grad` tensors autograd allocated are gathered by ONE kernel."""
        if self.grads_as_views:
            return self.g
        from . import capi
        srcs, keep, dt = [], [], None
        for seg, prm in zip(self.layout, self._train_params):
            if prm.grad is None:
                raise RuntimeError(f"parameter {seg.name} has no gradient")
            gt = prm.grad
            if not gt.is_contiguous():
                gt = gt.contiguous()
                keep.append(gt)          # the temporary must outlive the gather launch (its block could be reused)
            dt = gt.dtype if dt is None else dt
            if gt.dtype != dt:
                raise RuntimeError("mixed gradient dtypes")
            srcs.append(gt.data_ptr())
        if self._gather_tables is None:
            offs = torch.tensor([s.offset for s in self.layout], dtype=torch.int64)
            sizes = torch.tensor([s.numel for s in self.layout], dtype=torch.int64)
            self._gather_tables = (offs.to(self.device), sizes.to(self.device))
            self._gather_srcs_host = None
            self._gather_srcs_dev = torch.empty(len(srcs), dtype=torch.int64, device=self.device)
        offs_d, sizes_d = self._gather_tables
        if srcs != self._gather_srcs_host:
            # the pointer table is persistent: autograd usually hands back the same gradient storage step after
            # step, so it is re-uploaded only when an address changed
            self._gather_srcs_dev.copy_(torch.tensor(srcs, dtype=torch.int64))
            self._gather_srcs_host = list(srcs)
        capi.gather_segments(self.g, self._gather_srcs_dev, offs_d, sizes_d,
                             capi.F32 if dt == torch.float32 else capi.BF16, self.n)
        if keep:
            # stream-ordered: the caching allocator only reuses these blocks for work queued after the gather
            for tmp in keep:
                tmp.record_stream(torch.cuda.current_stream(self.device))
        return self.g

    def named_views(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        return self.layout.views(flat)
