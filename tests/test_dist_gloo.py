"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): shard arithmetic, the collectives
the path uses, and the cross-rank protocol of the top-k select (histogram all-reduce -> scan ->
filtered histogram all-reduce -> scan -> per-rank tie bases), with the oracle's key function
standing in for the CUDA histogram kernels (which need a GPU)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

WORLD = 2


def _worker(rank, port, fn_name, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        globals()[fn_name](rank, tmp)
    finally:
        dist.destroy_process_group()


def _spawn(fn_name, tmp_path):
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(port, fn_name, str(tmp_path)), nprocs=WORLD, join=True)


# ------------------------------------------------------------------------------------------------
def _shards_and_collectives(rank, tmp):
    from sfron_b200.dist import ShardGroup
    n = 1000 + 37                                        # ragged: not a multiple of 16 * world
    sg = ShardGroup(n)
    assert sg.bounds[0][0] == 0 and sg.bounds[-1][1] == n and sg.bounds[0][1] == sg.bounds[1][0]
    assert sg.lo % 16 == 0
    # gradient all-reduce (mean over ranks) and the shard view
    g = torch.full((n,), float(rank + 1))
    shard = sg.reduce_gradients_(g, average=True)
    assert torch.allclose(g, torch.full((n,), 1.5)) and shard.numel() == sg.n_local
    assert shard.data_ptr() == g[sg.lo:].data_ptr()      # a view, not a copy
    # clip-norm scalar: sum of the per-shard sums of squares == the global one
    full = torch.arange(n, dtype=torch.float64)
    part = sg.local(full).pow(2).sum().reshape(1)
    sg.all_reduce_(part)
    assert part.item() == full.pow(2).sum().item()
    # all-gather of the updated weights, uneven shards
    p = torch.zeros(n)
    sg.local(p).fill_(rank + 1.0)
    sg.all_gather_params_(p)
    want = torch.cat([torch.full((hi - lo,), r + 1.0) for r, (lo, hi) in enumerate(sg.bounds)])
    assert torch.equal(p, want)


def _spans_of_a_parted_vector(rank, tmp):
    """ShardGroup(parts=P): every piece of the vector is split over the ranks; the spans of all ranks partition [0, n)
    exactly, start on 16-element boundaries, and `local()` is their concatenation."""
    from sfron_b200.dist import ShardGroup
    n = 10_000 + 5
    for parts in (2, 3, 4):
        unit = 16 * WORLD * parts
        n_pad = -(-n // unit) * unit
        sg = ShardGroup(n, padded_len=n_pad, parts=parts)
        assert len(sg.spans) == parts and all(g % 16 == 0 for g, _, c in sg.spans if c)
        assert [off for _, off, _ in sg.spans] == [sum(c for _, _, c in sg.spans[:i]) for i in range(parts)]
        mine = torch.zeros(n, dtype=torch.int32)
        for g, _, c in sg.spans:
            mine[g:g + c] += 1
        dist.all_reduce(mine)
        assert bool((mine == 1).all()), "spans of all ranks must cover every element exactly once"
        full = torch.arange(n, dtype=torch.float32)
        assert torch.equal(sg.local(full), torch.cat([full[g:g + c] for g, _, c in sg.spans]))
        assert sg.local(full).numel() == sg.n_local
        piece = n_pad // parts                                   # piece k starts at k * piece on every rank
        assert all(k * piece <= g < (k + 1) * piece or c == 0 for k, (g, _, c) in enumerate(sg.spans))


def _select_protocol(rank, tmp):
    from oracle import sfron_oracle as O
    from sfron_b200.dist import ShardGroup, scan_from_top, tie_bases
    g = torch.Generator().manual_seed(0)
    n = 40_003
    x = torch.randint(0, 40, (n,), generator=g).float() * 0.125        # heavy ties
    x[torch.rand(n, generator=g) < 0.2] = 0.0
    x *= torch.where(torch.rand(n, generator=g) < 0.5, -1.0, 1.0)
    sg = ShardGroup(n)
    mine = sg.local(x)
    key = O.select_key(mine)
    for k in (1, n // 7, n // 2, n - 5, n):
        # pass 0: histogram of key[30:16], all-reduced
        bins0 = torch.bincount(key >> 16, minlength=32768)
        sg.all_reduce_(bins0)
        prefix, above0 = scan_from_top(bins0, k)
        assert prefix >= 0
        k_in_bin = k - above0
        # pass 1: low 16 bits of the keys matching the prefix
        sel = (key >> 16) == prefix
        local_bins1 = torch.bincount(key[sel] & 0xFFFF, minlength=65536)
        bins1 = local_bins1.clone()
        sg.all_reduce_(bins1)
        b, above1 = scan_from_top(bins1, k_in_bin)
        thr = (prefix << 16) | b
        budget = k_in_bin - above1
        # tie bases: all-gather one count per rank
        eq_local = local_bins1[b].reshape(1)
        gathered = [torch.zeros_like(eq_local) for _ in range(WORLD)]
        dist.all_gather(gathered, eq_local)
        base = tie_bases([int(t) for t in gathered])[rank]
        # apply, shard-local
        tie = key == thr
        rank_in_ties = base + torch.cumsum(tie.to(torch.int64), 0) - 1
        local_mask = (key > thr) | (tie & (rank_in_ties < budget))
        want = sg.local(O.topk_mask_flat(x, k)).bool()
        assert torch.equal(local_mask, want), (rank, k)
        # the two-pass form the kernels use: pass 1 already wrote `prefix(key) > prefix` for every element and staged
        # the prefix-matching ones; after scan 1 only that list is walked (low bits above the threshold's -> 1, ties
        # ranked from the per-rank base), the vector is not read a third time
        provisional = (key >> 16) > prefix
        cand = torch.nonzero(sel).flatten()                       # staged (flat index, low 16 bits)
        low = key[cand] & 0xFFFF
        two_pass = provisional.clone()
        two_pass[cand[low > b]] = True
        ties = cand[low == b]
        two_pass[ties[(base + torch.arange(ties.numel())) < budget]] = True
        assert torch.equal(two_pass, want), (rank, k)


def _bucketed_gradient_exchange(rank, tmp):
    """BucketedGradReducer == one monolithic all-reduce: the FlatParams constructor needs no GPU (only kernels do), so
    the hook / bucket logic runs here on CPU tensors with gloo."""
    import torch.nn as nn
    from sfron_b200.dist import BucketedGradReducer, ShardGroup
    from sfron_b200.flat import FlatParams
    torch.manual_seed(0)                                  # identical weights on both ranks
    model = nn.Sequential(nn.Linear(7, 33), nn.Tanh(), nn.Linear(33, 19), nn.Tanh(), nn.Linear(19, 5))
    model[2].bias.requires_grad_(False)                   # a frozen parameter in the middle
    import copy
    plain = copy.deepcopy(model)                          # same weights, ordinary per-tensor gradients, no hooks
    flat = FlatParams(model, torch.device("cpu"))
    sg = ShardGroup(flat.n)
    red = BucketedGradReducer(flat, sg, bucket_bytes=200 * 4)
    assert len(red.buckets) >= 3 and red.buckets[0][0] == 0 and red.buckets[-1][1] == flat.n
    assert all(a[1] == b[0] for a, b in zip(red.buckets, red.buckets[1:]))
    for step in range(3):
        x = torch.randn(4, 7, generator=torch.Generator().manual_seed(10 * step + rank))    # different data per rank
        def loss_of(m):
            if step == 2:                                 # a pass that does not reach the first layer's parameters
                return m[4](torch.tanh(m[2](torch.tanh(x @ torch.ones(7, 33))))).pow(2).sum()
            return m(x).pow(2).sum()

        plain.zero_grad()
        loss_of(plain).backward()
        want = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1)
                          for p in plain.parameters() if p.requires_grad])
        dist.all_reduce(want)
        want /= WORLD
        flat.zero_grad()
        loss_of(model).backward()                         # buckets start their all-reduce while this runs
        shard = red.finish()
        assert torch.equal(flat.g, want), step
        assert shard.data_ptr() == flat.g[sg.lo:].data_ptr() and shard.numel() == sg.n_local
    red.remove()


def test_spans_of_a_parted_vector(tmp_path):
    _spawn("_spans_of_a_parted_vector", tmp_path)


def test_bucketed_gradient_exchange(tmp_path):
    _spawn("_bucketed_gradient_exchange", tmp_path)


def test_shards_and_collectives(tmp_path):
    _spawn("_shards_and_collectives", tmp_path)


def test_select_protocol_two_ranks(tmp_path):
    _spawn("_select_protocol", tmp_path)
