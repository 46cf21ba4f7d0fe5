"""Reference on-disk formats <-> flat vectors (SURVEY.md §8b "File formats").

The three stages of the reference talk to each other through files; a drop-in must read and
write them byte-compatibly:

  Fisher      torch.save(dict[name -> fp32 CPU tensor of param shape | int 0])
              forget_fisher.pt / remain_fisher.pt   sfron.py:293,320; runners/diffusion.py:1299,1364;
              DiT/generate_fisher.py:251,291; fisher/nude_{forget,remain}.pt SD generate_fisher.py:79,129
  ratio mask  dict[name -> bool tensor | int 0]   fisher_{th}.pt (DDPM/generate_fisher_mask.py:48,
              DiT/generate_mask.py:46), nude_mask_{th}.pt (SD generate_fisher_mask.py:48);
              `th` formatted by Python str(float)
  top-k mask  dict[name -> int64 0/1]             results/cifar10/mask/{label}/with_{ratio}.pt
              (runners/diffusion.py:1000-1036)
  FIM         pickle dict[name -> fp32 tensor]     fisher_dict.pkl (runners/diffusion.py:346-352)
  ckpt        DDPM list [model_sd, opt_sd, step, ema_shadow] (runners/diffusion.py:1188-1199);
              DiT dict {model, ema, opt, args} (DiT/forget.py:348-355)

Names: DDPM and DiT keys carry the DataParallel `module.` prefix; SD keys are U-Net-local;
Classification keys are bare.  Parameters that never get a gradient keep the reference's
int-0 placeholder.  Everything here is plain host-side torch: no kernels.
"""
from __future__ import annotations

import os
import pickle
from typing import Dict, List, Optional, Sequence

import torch

from .flat import FlatLayout


def _cpu(flat: torch.Tensor) -> torch.Tensor:
    return flat.detach().to("cpu")


def load_file(path):
    """torch.load of a reference-format dict, memory-mapped when the file allows it: the tensors are then
    copied to the device straight from the page cache instead of through a private host copy of the file."""
    try:
        return torch.load(path, weights_only=False, mmap=True)
    except (RuntimeError, ValueError):          # legacy (non-zip) files cannot be mapped
        return torch.load(path, weights_only=False)


def flat_to_dict(layout: FlatLayout, flat: torch.Tensor, *, all_names: Optional[Sequence[str]] = None,
                 prefix: str = "", dtype: Optional[torch.dtype] = None) -> Dict[str, object]:
    """name -> CPU tensor of the parameter's shape, in `all_names` order; names that are not in the
    layout (frozen parameters) get the reference's int `0` placeholder.  ONE device->host copy; the
    per-name tensors are views of that host vector (torch.save writes the shared storage once — at SD
    size a per-tensor clone would be another 3.4 GB of host memcpy per Fisher file)."""
    host = _cpu(flat)
    if dtype is not None:
        host = host.to(dtype)
    if host.data_ptr() == flat.data_ptr():
        host = host.clone()                  # flat already lived on the host: do not alias the caller's buffer
    out: Dict[str, object] = {}
    for name in (all_names if all_names is not None else layout.names):
        if name in layout:
            s = layout.segment(name)
            out[prefix + name] = host[s.offset:s.offset + s.numel].view(s.shape)
        else:
            out[prefix + name] = 0
    return out


def dict_to_flat(layout: FlatLayout, d: Dict[str, object], *, prefix: str = "", dtype=torch.float32,
                 device="cpu") -> torch.Tensor:
    """Pack a reference-format dict into a flat vector on `device` (int-0 placeholders are skipped).
    Every tensor is copied straight into its slice of the destination (dtype conversion included), without
    a packed host intermediate."""
    out = torch.zeros(layout.numel, dtype=dtype, device=device)
    for s in layout:
        v = d[prefix + s.name]
        if not torch.is_tensor(v):
            raise ValueError(f"{prefix + s.name}: placeholder {v!r} for a trainable parameter")
        if tuple(v.shape) != s.shape:
            raise ValueError(f"{prefix + s.name}: shape {tuple(v.shape)} != {s.shape}")
        out[s.offset:s.offset + s.numel].copy_(v.reshape(-1))
    return out


# ---- Fisher ----------------------------------------------------------------------------------------
def save_fisher(path: str, layout: FlatLayout, flat: torch.Tensor, *, all_names=None, prefix: str = "") -> None:
    torch.save(flat_to_dict(layout, flat, all_names=all_names, prefix=prefix), path)


def load_fisher(path: str, layout: FlatLayout, *, prefix: str = "", device="cpu") -> torch.Tensor:
    return dict_to_flat(layout, load_file(path), prefix=prefix, device=device)


def save_fim_pickle(path: str, layout: FlatLayout, flat: torch.Tensor, *, prefix: str = "") -> None:
    """`pickle.dump(fisher_dict)` of DDPM save_fim (runners/diffusion.py:346-352)."""
    with open(path, "wb") as f:
        pickle.dump(flat_to_dict(layout, flat, prefix=prefix), f)


# ---- masks -----------------------------------------------------------------------------------------
def threshold_tag(th) -> str:
    """How the reference formats the threshold into the file name: f"fisher_{th}.pt" with th a float
    for the argparse scripts (`fisher_1.0.pt`), whatever type the list holds for DiT (`fisher_1.pt`
    when the default int list is used, DiT/generate_mask.py:46,55)."""
    return str(th)


def ratio_mask_to_dict(layout: FlatLayout, mask_u8: torch.Tensor, *, all_names=None, prefix: str = ""):
    """bool tensors, as `weight_saliency >= th` produces."""
    return flat_to_dict(layout, mask_u8[:layout.numel], all_names=all_names, prefix=prefix, dtype=torch.bool)


def topk_mask_to_dict(layout: FlatLayout, mask_u8: torch.Tensor, *, all_names=None, prefix: str = ""):
    """int64 0/1 tensors, as `torch.zeros_like(ranks)` produces (runners/diffusion.py:1026-1030)."""
    return flat_to_dict(layout, mask_u8[:layout.numel], all_names=all_names, prefix=prefix, dtype=torch.int64)


def load_mask(path_or_dict, layout: FlatLayout, *, prefix: str = "", device="cpu") -> torch.Tensor:
    """Any reference mask file (bool or int64) -> flat uint8 0/1."""
    d = load_file(path_or_dict) if isinstance(path_or_dict, (str, os.PathLike)) else path_or_dict
    return dict_to_flat(layout, d, prefix=prefix, dtype=torch.uint8, device=device)


def sparsity_percent(zero_count: int, total: int) -> float:
    """The number behind the reference's `Total sparsity th:{th} weight:{...}` print."""
    return zero_count / total * 100


# ---- optimizer state / checkpoints -------------------------------------------------------------------
def adam_state_dict(layout: FlatLayout, m: torch.Tensor, v: torch.Tensor, step: int, *, lr: float,
                    betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                    amsgrad: bool = False, n_frozen_before: Optional[Sequence[int]] = None,
                    param_count: Optional[int] = None, decoupled: bool = False) -> dict:
    """A torch.optim.Adam/AdamW `state_dict()` built from the flat moments, loadable by the
    reference's `optimizer.load_state_dict` (DDPM ckpt slot 1, DiT ckpt["opt"]).

    torch indexes optimizer state by the position of the parameter in the list handed to the
    optimizer (all of `model.parameters()`, frozen ones included but without state);
    `n_frozen_before[i]` = number of frozen parameters preceding trainable parameter i."""
    mh, vh = _cpu(m), _cpu(v)
    state = {}
    for i, s in enumerate(layout):
        idx = i + (n_frozen_before[i] if n_frozen_before is not None else 0)
        state[idx] = {
            "step": torch.tensor(float(step)),
            "exp_avg": mh[s.offset:s.offset + s.numel].clone().view(s.shape),
            "exp_avg_sq": vh[s.offset:s.offset + s.numel].clone().view(s.shape),
        }
    total = param_count if param_count is not None else len(layout)
    group = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=amsgrad,
                 maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                 decoupled_weight_decay=decoupled, params=list(range(total)))
    return {"state": state, "param_groups": [group]}


def load_adam_state(layout: FlatLayout, sd: dict, *, n_frozen_before=None, device="cpu"):
    """Inverse of `adam_state_dict`: (m, v, step) flat vectors from a torch optimizer state dict."""
    m = torch.zeros(layout.numel)
    v = torch.zeros(layout.numel)
    step = 0
    for i, s in enumerate(layout):
        idx = i + (n_frozen_before[i] if n_frozen_before is not None else 0)
        st = sd["state"].get(idx)
        if st is None:
            continue
        m[s.offset:s.offset + s.numel].copy_(st["exp_avg"].reshape(-1))
        v[s.offset:s.offset + s.numel].copy_(st["exp_avg_sq"].reshape(-1))
        step = int(st["step"]) if not torch.is_tensor(st["step"]) else int(st["step"].item())
    return m.to(device), v.to(device), step


def ddpm_checkpoint(model_sd: dict, opt_sd: dict, step: int, ema_shadow: Optional[dict]) -> list:
    """`states = [model.state_dict(), optimizer.state_dict(), step, ema_helper.state_dict()]`
    (runners/diffusion.py:1188-1195)."""
    states: List[object] = [model_sd, opt_sd, step]
    if ema_shadow is not None:
        states.append(ema_shadow)
    return states


def dit_checkpoint(model_sd: dict, ema_sd: dict, opt_sd: dict, args) -> dict:
    """DiT/forget.py:348-353."""
    return {"model": model_sd, "ema": ema_sd, "opt": opt_sd, "args": args}
