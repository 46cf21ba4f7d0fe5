"""Drop-ins for the three ratio-mask scripts — same flags, same input and output files:

  python -m sfron_b200.methods.masks ddpm --ckpt_folder DIR [--threshold 1.0]
        = DDPM/generate_fisher_mask.py:17-48   (forget_fisher.pt, remain_fisher.pt -> fisher_{th}.pt)
  python -m sfron_b200.methods.masks sd   --ckpt_folder DIR [--threshold 1.0]
        = SD/train-scripts/generate_fisher_mask.py:17-48 (nude_forget.pt, nude_remain.pt -> nude_mask_{th}.pt)
  python -m sfron_b200.methods.masks dit  --mask-path DIR --forget-class C [C ...] [--thresholds ...]
        = DiT/generate_mask.py:16-57   (per class: forget_fisher.pt, remain_fisher.pt -> fisher_{th}.pt)

The DiT form computes all thresholds in ONE pass over the two Fisher vectors (sfr_ratio_mask_multi);
the reference re-reads both per threshold.
"""
from __future__ import annotations

import argparse
import os
from typing import Dict, List, Sequence

import torch

from .. import capi, formats
from ..flat import FlatLayout


def _layout_of(fisher: Dict[str, object]) -> FlatLayout:
    """Entries that never received a gradient are the int 0 placeholder (DiT pos_embed): they are
    not part of the flat vector and keep their placeholder in the output."""
    return FlatLayout((n, tuple(t.shape)) for n, t in fisher.items() if torch.is_tensor(t))


def ratio_masks_from_fishers(forget_fisher: Dict[str, object], remain_fisher: Dict[str, object],
                             thresholds: Sequence[float], device="cuda") -> List[Dict[str, object]]:
    layout = _layout_of(forget_fisher)
    names = list(forget_fisher.keys())
    ff = formats.dict_to_flat(layout, forget_fisher, device=device)
    rf = formats.dict_to_flat(layout, remain_fisher, device=device)
    out = []
    for lo in range(0, len(thresholds), capi.MAX_THRESHOLDS):
        ths = list(thresholds[lo:lo + capi.MAX_THRESHOLDS])
        stride = (layout.numel + 15) // 16 * 16
        masks = torch.empty(len(ths), stride, dtype=torch.uint8, device=device)
        zeros = torch.zeros(capi.MAX_THRESHOLDS, dtype=torch.int64, device=device)
        capi.ratio_mask_multi(ff, rf, [float(t) for t in ths], masks, zeros)
        zeros = zeros.cpu().tolist()
        for i, th in enumerate(ths):
            print(f"Total sparsity th:{th} weight:{formats.sparsity_percent(zeros[i], layout.numel)}")
            out.append(formats.ratio_mask_to_dict(layout, masks[i], all_names=names))
    return out


def generate_fisher_mask(ckpt_folder: str, threshold: float = 1.0, *, forget_name="forget_fisher.pt",
                         remain_name="remain_fisher.pt", out_fmt="fisher_{th}.pt", device="cuda") -> str:
    forget_fisher = formats.load_file(os.path.join(ckpt_folder, forget_name))
    remain_fisher = formats.load_file(os.path.join(ckpt_folder, remain_name))
    (mask,) = ratio_masks_from_fishers(forget_fisher, remain_fisher, [threshold], device)
    path = os.path.join(ckpt_folder, out_fmt.format(th=formats.threshold_tag(threshold)))
    torch.save(mask, path)
    return path


def generate_mask_dit(mask_path: str, forget_class: Sequence[int], thresholds: Sequence[float],
                      device="cuda") -> List[str]:
    written = []
    for cls in forget_class:
        folder = os.path.join(mask_path, str(cls))
        forget_fisher = formats.load_file(os.path.join(folder, "forget_fisher.pt"))
        remain_fisher = formats.load_file(os.path.join(folder, "remain_fisher.pt"))
        for name, v in forget_fisher.items():
            if not torch.is_tensor(v):
                print(f"{name} {v}")                      # the reference's except-branch print
        masks = ratio_masks_from_fishers(forget_fisher, remain_fisher, thresholds, device)
        for th, mask in zip(thresholds, masks):
            path = os.path.join(folder, f"fisher_{formats.threshold_tag(th)}.pt")
            torch.save(mask, path)
            written.append(path)
    return written


def build_parser() -> argparse.ArgumentParser:
    """One sub-command per family, each with the flags of the reference script it stands for:
    DDPM/generate_fisher_mask.py:17-24, SD/train-scripts/generate_fisher_mask.py:17-24, DiT/generate_mask.py:50-56."""
    ap = argparse.ArgumentParser(prog="sfron_b200.methods.masks")
    sub = ap.add_subparsers(dest="family", required=True)
    for fam in ("ddpm", "sd"):
        p = sub.add_parser(fam)
        p.add_argument("--ckpt_folder", type=str, required=True, help="Path to fisher ckpt path")
        p.add_argument("--threshold", type=float, default=1.0, help="Saliency map threshold, lambda in paper")
    p = sub.add_parser("dit")
    p.add_argument("--mask-path", required=True, type=str, default="./mask")
    p.add_argument("--forget-class", nargs="+", type=int, required=True)
    p.add_argument("--thresholds", nargs="+", type=float, default=[0.5, 1, 3, 5, 10])
    return ap


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.family == "ddpm":
        generate_fisher_mask(args.ckpt_folder, args.threshold)
    elif args.family == "sd":
        generate_fisher_mask(args.ckpt_folder, args.threshold, forget_name="nude_forget.pt",
                             remain_name="nude_remain.pt", out_fmt="nude_mask_{th}.pt")
    else:
        generate_mask_dit(args.mask_path, args.forget_class, args.thresholds)


if __name__ == "__main__":
    main()
