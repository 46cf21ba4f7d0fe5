#!/usr/bin/env python
"""Print the BASELINE.md §6 tables from the committed measurement files under profiles/."""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
rows = [json.loads(l) for l in open(os.path.join(P, "r1_sweep.jsonl"))]
# the K2b rows were re-measured after the select became two-pass: those lines supersede the first sweep's
two_pass = [json.loads(l) for l in open(os.path.join(P, "r1_sweep_select_two_pass.jsonl"))]
fresh = {(r["n"], r["kernel"]) for r in two_pass}
rows = two_pass + [r for r in rows if (r["n"], r["kernel"]) not in fresh]
def get(n, k, dt=None):
    for r in rows:
        if r["n"] == n and r["kernel"] == k and (dt is None or r["grad_dtype"] == dt):
            return r
    return None
def cell(r): return f"{r['GBps']:.0f}" if r else "—"
def frac(r): return f"{r['frac_of_measured_peak']:.3f}" if r else "—"
N3 = 675129632
print("| config / n | GPUs | kernel | B/elem | GB/s | frac of 6464.6 | parity |\n|---|---|---|---|---|---|---|")
spec = [("K1 `fisher_accum`", "fisher_accum", "f32", "bit-exact"), ("K1, 4 per-sample rows", "fisher_accum_rows4", "f32", "bit-exact"),
        ("K1, bf16 gradients", "fisher_accum", "bf16", "bit-exact (exact widening)"), ("K2a `ratio_mask`", "ratio_mask", None, "mask bit-exact"),
        ("K2a, 5 thresholds in one pass", "ratio_mask_x5_thresholds", None, "mask bit-exact"), ("clip norm `masked_sumsq`", "masked_sumsq", "f32", "1e-7 of exact"),
        ("K3 AdamW masked", "fused_update_adamw_masked_clip", "f32", "1e-6"), ("K3 AdamW + EMA", "fused_update_adamw_ema", "f32", "1e-6"),
        ("K3 AdamW masked, bf16 gradients", "fused_update_adamw_masked_clip", "bf16", "1e-6 on the fp32 master"),
        ("K3 AdamW + EMA, bf16 gradients", "fused_update_adamw_ema", "bf16", "1e-6 on the fp32 master"),
        ("K3 SGD-momentum masked", "fused_update_sgd_masked", "f32", "1e-6"), ("K3 SGD-momentum + slow-fast", "fused_update_sgd_slowfast", "f32", "1e-6"),
        ("K2b top-k select, total", "topk_select_total", None, "mask bit-exact (stable ties)"),
        ("K2b top-k select, three-read fallback form", "topk_select_total_three_reads", None, "mask bit-exact (identical bytes)"),
        ("K2b hist pass 0 / pass 1 + provisional mask / candidate-only apply", None, None, "")]
for name, k, dt, par in spec:
    if k is None:
        a, b, c = get(N3, "topk_hist_pass0"), get(N3, "topk_hist_pass1_with_mask"), get(N3, "topk_apply_candidates_only")
        print(f"| 3 | 1 | {name} | 4 / 5 / — | {cell(a)} / {cell(b)} / {c['ms']} ms | {frac(a)} / {frac(b)} / — | |")
        continue
    r = get(N3, k, dt)
    print(f"| 3 (N3 = 675,129,632) | 1 | {name} | {r['bytes_per_elem']} | {cell(r)} | {frac(r)} | {par} |")
for label, n in (("2 (N2 = 38,632,323)", 38632323), ("4 (N4 = 859,520,964)", 859520964), ("5 (n = 2e9)", 2000000000)):
    ks = [("fisher_accum", "f32"), ("ratio_mask", None), ("masked_sumsq", "f32"), ("fused_update_adamw_masked_clip", "f32"), ("fused_update_adamw_ema", "f32"), ("topk_select_total", None)]
    rs = [get(n, k, dt) for k, dt in ks]
    print(f"| {label} | 1 | K1 / K2a / norm / K3 masked / K3+EMA / K2b | 12 / 9 / 5 / 29 / 36 / 13 | " + " / ".join(cell(r) for r in rs) + " | " + " / ".join(frac(r) for r in rs) + " | as above |")
print()
print("| GPUs | GB/s (all ranks) | frac of N × 6464.6 | ms/step | steps/s per GPU | K2b select (ms, frac) | e2e from pinned host (GB/s) |\n|---|---|---|---|---|---|---|")
for n in (1, 2, 4, 8):
    d = json.loads(open(os.path.join(P, f"r1_bench_n{n}.json")).read())
    ex = d.get("extra_kernels", {}).get("topk_select_k_half", {})
    print(f"| {n} | {d['value']:,.0f} | {d['hot_path_frac_of_peak']:.3f} | {d['ms_per_step']:.2f} | {d['steps_per_s']:.1f} | {ex.get('ms','—')} , {ex.get('frac','—')} | {d['e2e']['value']:.0f} |")
d = json.loads(open(os.path.join(P, "r1_bench_n1.json")).read())
print("\ncpu_baseline:", d["cpu_baseline"])
for f in ("r1_dit_e2e_1gpu.jsonl", "r1_dit_e2e_2gpu.jsonl", "r1_dit_e2e_8gpu.jsonl"):
    for l in open(os.path.join(P, f)):
        r = json.loads(l)
        print(f, r["n_gpus"], {k: {a: round(b, 2) for a, b in v.items()} for k, v in r.items() if isinstance(v, dict)})
print()
for l in open(os.path.join(P, "r1_unet_e2e_1gpu.jsonl")):
    r = json.loads(l)
    print("unet_e2e", r["family"], r["params"], {k: v for k, v in r.items() if k in ("stock", "ours", "speedup", "mask_agreement")})
