"""Probe: the whole K2b select (14 launches, all stream-ordered, no host read) captured in ONE CUDA graph and
replayed, against eager launches — at the DDPM size, where the select is launch-bound, and at DiT-XL/2 size."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfron_b200 as sfr

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=9):
    ts = []
    for _ in range(iters + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts[2:])[len(ts[2:]) // 2]


for n in (38_632_323, 675_129_632):
    x = torch.empty(n, device=dev).normal_(0, 1e-2, generator=torch.Generator(device=dev).manual_seed(0))
    hp = sfr.HotPath(n, dev, sfr.OptConfig())
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    k = n // 2
    for _ in range(3):
        hp.topk_mask(x, k, out=out)
    eager = timed(lambda: hp.topk_mask(x, k, out=out))
    ref = out.clone()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        hp.topk_mask(x, k, out=out)
    out.zero_()
    graph = timed(g.replay)
    print(json.dumps({"n": n, "eager_ms": round(eager, 4), "graph_ms": round(graph, 4), "identical": bool(torch.equal(out, ref)),
                      "selected": int(out.sum(dtype=torch.int64))}), flush=True)
    del x, hp, out, ref, g
