"""CPU oracle of the SFR-on hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement, in plain CPU torch, of the reference's algorithm for the path
(K1nght/Unified-Unlearning-w-Remain-Geometry, paths below relative to its root).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import this module, and only as the checker / baseline.  The product path
(`sfron_b200`) never imports it and has no CPU fallback.

The reference's arithmetic on this path lives in PyTorch itself (elementwise ops,
torch.optim.{SGD,Adam,AdamW}, torch.nn.utils.clip_grad_norm_, torch.argsort), whose
source is not under /root/reference; reference pins: torch==1.12.0+cu113
(Classification/README.md:8), torch==2.0.1 (DDPM/requirements.txt:71), pytorch>=1.13
(DiT/environment.yml:7).  This oracle therefore calls the SAME torch functions in the
SAME order the reference does, per named tensor, on CPU fp32 — it does not re-derive
their arithmetic.

Pinning: the reference ships no tests or golden vectors.  The oracle is
pinned instead against outputs of the reference ITSELF executed in the build container
(tests/golden/make_golden.py imports /root/reference and records fixtures;
tests/test_oracle_golden.py replays them through this module).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------- a12 schedules
def cosine_lr_scheduler(base_lr: float, current_epoch: int, T_max: int) -> float:
    """sfron.py:45-46, DDPM/functions/losses.py:71-72, DiT/forget.py:35-36."""
    return base_lr * (1 + math.cos(math.pi * current_epoch / T_max)) / 2


# --------------------------------------------------------------------------- a1-a5 Fisher
def fisher_init(names: Iterable[str]) -> Dict[str, object]:
    """`forget_gradients[name] = 0` (sfron.py:274-276; DiT/generate_fisher.py:217-219)."""
    return {n: 0 for n in names}


def fisher_accumulate(acc: Dict[str, object], grads: Dict[str, Optional[Tensor]], divisor) -> None:
    """`F[name] += param.grad.data.cpu()**2 / len(loader)` for params that have a grad.

    sfron.py:288-291,315-318; DDPM/runners/diffusion.py:1277-1281,1342-1346;
    DiT/generate_fisher.py:236-239,276-279; SD/train-scripts/generate_fisher.py:73-76,123-126.
    """
    for name, g in grads.items():
        if g is not None:
            acc[name] += g.cpu() ** 2 / divisor


def clip_grad_norm(grads: Sequence[Tensor], max_norm: float) -> Tensor:
    """torch.nn.utils.clip_grad_norm_ on a list of gradient tensors (in place).

    sfron.py:205; DDPM/runners/diffusion.py:1131-1136,1169-1174,1270-1275; DiT/forget.py:293-298.
    """
    params = []
    for g in grads:
        p = torch.nn.Parameter(g.detach())  # shares storage; clip_grad_norm_ only touches .grad
        p.grad = g
        params.append(p)
    return torch.nn.utils.clip_grad_norm_(params, max_norm)


def fisher_accumulate_clipped(acc, grads: Dict[str, Tensor], divisor, max_norm: float) -> None:
    """DDPM generate_fisher: clip the batch gradient to `max_norm`, THEN square
    (DDPM/runners/diffusion.py:1270-1281)."""
    gs = {k: v.clone() for k, v in grads.items() if v is not None}
    clip_grad_norm(list(gs.values()), max_norm)
    fisher_accumulate(acc, gs, divisor)


def per_sample_fim(acc: Dict[str, Tensor], rows: List[Dict[str, Tensor]], dataset_len: int) -> None:
    """`fisher_dict[name] += tmp_i[name]**2 / len(dataset)` for every sample i, in order
    (DDPM/runners/diffusion.py:337-344)."""
    for name in acc:
        for r in rows:
            acc[name].data += (r[name].data ** 2) / dataset_len


# --------------------------------------------------------------------------- a6 ratio mask
def ratio_mask(forget_fisher: Dict[str, object], remain_fisher: Dict[str, object], th):
    """Per-name `((F_f + 1e-15) / (F_r + 1e-15)) >= th`, with the int-0 placeholder kept for
    entries that never received a gradient (DiT pos_embed).

    sfron.py:322-336; DDPM/generate_fisher_mask.py:36-48; DiT/generate_mask.py:27-46;
    SD/train-scripts/generate_fisher_mask.py:36-48.
    Returns (mask dict, zero count, total count).
    """
    masks: Dict[str, object] = {}
    total_cnt = 0
    w_cnt = 0
    for name in forget_fisher.keys():
        masks[name] = 0
        ff, rf = forget_fisher[name], remain_fisher[name]
        if not (torch.is_tensor(ff) and torch.is_tensor(rf)):
            continue  # the reference's `except: pass` branch: 0 has no .numel()
        weight_saliency = (ff + 1e-15) / (rf + 1e-15)
        w = weight_saliency >= th
        total_cnt += w.numel()
        w_cnt += int(w.numel() - torch.count_nonzero(w))
        masks[name] = w
    return masks, w_cnt, total_cnt


# --------------------------------------------------------------------------- a7 top-k mask
def topk_mask(gradients: Dict[str, Tensor], ratio: float, stable: bool = True) -> Dict[str, Tensor]:
    """SalUn global top-k mask: `ranks = argsort(argsort(-cat(|g|))); mask = ranks < int(N*ratio)`.

    DDPM/runners/diffusion.py:1009-1034; Classification/unlearn/salun.py:170-193.
    `gradients` holds the accumulated (already abs'ed) gradients.  The reference calls
    torch.argsort without `stable=True`, which leaves the order among equal values
    unspecified; `stable=True` (default here) is the deterministic contract the CUDA path
    implements (lowest flat index wins among ties).  With stable=False this is the
    reference's literal call.  Returns int64 0/1 tensors of the parameter shapes.
    """
    all_elements = -torch.cat([t.flatten() for t in gradients.values()])
    threshold_index = int(len(all_elements) * ratio)
    positions = torch.argsort(all_elements, stable=stable)
    ranks = torch.argsort(positions, stable=stable)
    out: Dict[str, Tensor] = {}
    start = 0
    for key, tensor in gradients.items():
        n = tensor.numel()
        tensor_ranks = ranks[start:start + n]
        threshold_tensor = torch.zeros_like(tensor_ranks)
        threshold_tensor[tensor_ranks < threshold_index] = 1
        out[key] = threshold_tensor.reshape(tensor.shape)
        start += n
    return out


def topk_mask_flat(values: Tensor, k: int) -> Tensor:
    """Flat form of `topk_mask` with an explicit k (stable ties), uint8 0/1."""
    all_elements = -values.abs().flatten()
    positions = torch.argsort(all_elements, stable=True)
    ranks = torch.argsort(positions, stable=True)
    return (ranks < k).to(torch.uint8)


def select_key(values: Tensor) -> Tensor:
    """The order-preserving integer key the CUDA select uses: 0 for NaN, else bits(|x|)+1
    (int64 to hold the +1 without sign trouble)."""
    bits = values.abs().contiguous().view(torch.int32).to(torch.int64) & 0x7FFFFFFF
    return torch.where(bits > 0x7F800000, torch.zeros_like(bits), bits + 1)


# --------------------------------------------------------------------------- a8-a11 update
def apply_mask_(params: Dict[str, torch.nn.Parameter], mask: Dict[str, object]) -> None:
    """`param.grad *= mask[name].to(param.grad.device)` for params with a grad.

    sfron.py:201-204; DDPM/runners/diffusion.py:1126-1129; DiT/forget.py:289-292;
    SD gradient_ascent.py:94-99 (and the intended behaviour of nsfw_removal.py:157-160).
    """
    for name, param in params.items():
        if param.grad is not None:
            param.grad *= mask[name].to(param.grad.device)


def ema_ddpm_(shadow: Dict[str, Tensor], params: Dict[str, torch.nn.Parameter], mu: float) -> None:
    """EMAHelper.update: `shadow = (1 - mu) * param + mu * shadow` (DDPM/models/ema.py:17-24)."""
    for name, param in params.items():
        if param.requires_grad:
            shadow[name].data = (1.0 - mu) * param.data + mu * shadow[name].data


def ema_dit_(ema: Dict[str, Tensor], params: Dict[str, Tensor], decay: float = 0.9999) -> None:
    """update_ema: `ema.mul_(decay).add_(param, alpha=1 - decay)` over ALL named parameters,
    frozen ones included (DiT/forget.py:52-62)."""
    for name, param in params.items():
        ema[name].mul_(decay).add_(param.data, alpha=1 - decay)


def slowfast_(params: Dict[str, torch.nn.Parameter], prev: Dict[str, Tensor], beta: float) -> None:
    """`update_parameters(model, ori_model, avg_fn)` then `ori_model = deepcopy(model)`:
    p <- (1 - beta) * p_prev + beta * p ; p_prev <- p     (sfron.py:30-37,126-127,255-257)."""
    with torch.no_grad():
        for name, p in params.items():
            p.copy_((1 - beta) * prev[name] + beta * p.detach())
            prev[name] = p.detach().clone()


def make_optimizer(kind: str, params: Sequence[torch.nn.Parameter], **kw) -> torch.optim.Optimizer:
    """The optimizers the reference constructs: SGD(momentum .9, wd 5e-4) sfron.py:167-170;
    Adam DDPM/functions/__init__.py:9-18; AdamW(wd 0) DiT/forget.py:199; Adam SD nsfw_removal.py:81."""
    if kind == "sgd":
        return torch.optim.SGD(params, kw["lr"], momentum=kw.get("momentum", 0.0),
                               weight_decay=kw.get("weight_decay", 0.0),
                               dampening=kw.get("dampening", 0.0))
    if kind == "adam":
        return torch.optim.Adam(params, lr=kw["lr"], weight_decay=kw.get("weight_decay", 0.0),
                                betas=(kw.get("beta1", 0.9), kw.get("beta2", 0.999)),
                                amsgrad=False, eps=kw.get("eps", 1e-8))
    if kind == "adamw":
        return torch.optim.AdamW(params, lr=kw["lr"], weight_decay=kw.get("weight_decay", 0.0),
                                 betas=(kw.get("beta1", 0.9), kw.get("beta2", 0.999)),
                                 eps=kw.get("eps", 1e-8))
    raise ValueError(kind)


class FlatReferenceLoop:
    """The forget-loop update sequence of the reference on explicit gradients.

    One instance = one model's parameters + ONE optimizer shared by the forget and the
    remain step (as in every reference loop), replayed on recorded / synthetic gradients:

      forget_step(g):  grad = g ; grad *= mask ; clip_grad_norm_ ; optimizer.step()
      remain_step(g):  grad = g ; [clip_grad_norm_] ; optimizer.step() ; EMA / slow-fast

    Order of mask and clip: "mask_then_clip" (SFR-on: sfron.py:201-206,
    runners/diffusion.py:1126-1138, DiT/forget.py:289-299) or "clip_then_mask"
    (SalUn-DDPM: runners/diffusion.py:579-590).
    """

    def __init__(self, shapes: Dict[str, Sequence[int]], theta0: Dict[str, Tensor], opt: str,
                 opt_kw: dict, ema_mode: str = "none", ema_a: float = 0.0):
        self.params = {n: torch.nn.Parameter(theta0[n].detach().clone().float().reshape(tuple(s)))
                       for n, s in shapes.items()}
        self.optimizer = make_optimizer(opt, list(self.params.values()), **opt_kw)
        self.ema_mode = ema_mode
        self.ema_a = ema_a
        if ema_mode in ("ddpm", "dit"):
            self.slow = {n: p.detach().clone() for n, p in self.params.items()}
        elif ema_mode == "slowfast":
            self.slow = {n: p.detach().clone() for n, p in self.params.items()}
        else:
            self.slow = {}

    def set_lr(self, lr: float) -> None:
        for group in self.optimizer.param_groups:
            group["lr"] = lr

    def _load_grads(self, grads: Dict[str, Tensor]) -> None:
        for n, p in self.params.items():
            p.grad = grads[n].detach().clone().reshape(p.shape)

    def forget_step(self, grads, mask=None, max_norm: Optional[float] = None,
                    order: str = "mask_then_clip") -> Optional[Tensor]:
        self._load_grads(grads)
        norm = None
        if order == "mask_then_clip":
            if mask is not None:
                apply_mask_(self.params, mask)
            if max_norm is not None:
                norm = torch.nn.utils.clip_grad_norm_(list(self.params.values()), max_norm)
        else:
            if max_norm is not None:
                norm = torch.nn.utils.clip_grad_norm_(list(self.params.values()), max_norm)
            if mask is not None:
                apply_mask_(self.params, mask)
        self.optimizer.step()
        return norm

    def remain_step(self, grads, max_norm: Optional[float] = None, ema: bool = True) -> Optional[Tensor]:
        self._load_grads(grads)
        norm = None
        if max_norm is not None:
            norm = torch.nn.utils.clip_grad_norm_(list(self.params.values()), max_norm)
        self.optimizer.step()
        if ema:
            self.slow_update()
        return norm

    def slow_update(self) -> None:
        if self.ema_mode == "ddpm":
            ema_ddpm_(self.slow, self.params, self.ema_a)
        elif self.ema_mode == "dit":
            with torch.no_grad():
                ema_dit_(self.slow, self.params, self.ema_a)
        elif self.ema_mode == "slowfast":
            slowfast_(self.params, self.slow, self.ema_a)

    # flat views for comparisons --------------------------------------------------------
    def flat(self, what: str = "p") -> Tensor:
        if what == "p":
            return torch.cat([p.detach().flatten() for p in self.params.values()])
        if what == "slow":
            return torch.cat([self.slow[n].detach().flatten() for n in self.params])
        st = self.optimizer.state
        key = {"m": "exp_avg", "v": "exp_avg_sq", "buf": "momentum_buffer"}[what]
        return torch.cat([st[p][key].detach().flatten() for p in self.params.values()])


# --------------------------------------------------------------------------- data-parallel gradient
def dp_reduce(grads: Sequence[Tensor], average: bool = True) -> Tensor:
    """The gradient a data-parallel step updates with, from the per-replica gradients.

    The reference wraps its models in `torch.nn.DataParallel` (DiT/forget.py:193, DiT/generate_fisher.py:173,
    DDPM/runners/diffusion.py:110,1060): backward of the replicated module ends in `ReduceAddCoalesced`
    (torch/nn/parallel/_functions.py), i.e. `comm.reduce_add`, which starts from the first replica's gradient and
    adds the others IN DEVICE ORDER.  With one process per GPU each rank's loss is the mean over its own batch,
    so the global-batch mean is that sum divided by the number of ranks (`average`)."""
    total = grads[0].detach().clone().float()
    for g in grads[1:]:
        total.add_(g.detach().float())
    if average:
        total = total / len(grads)
    return total


# --------------------------------------------------------------------------- flat reference-form ops
# The stock-torch op sequences of SURVEY.md §2.1 on ONE flat vector: the form timed as the
# CPU baseline (bench.py) and used for size-independent parity checks in the sweep.
def flat_fisher_accum(acc: Tensor, g: Tensor, divisor) -> Tensor:
    acc += g ** 2 / divisor
    return acc


def flat_ratio_mask(ff: Tensor, rf: Tensor, th) -> Tensor:
    return ((ff + 1e-15) / (rf + 1e-15)) >= th


def flat_masked_clip_(g: Tensor, mask: Optional[Tensor], max_norm: Optional[float]) -> None:
    if mask is not None:
        g *= mask
    if max_norm is not None:
        p = torch.nn.Parameter(g.detach())  # shares storage
        p.grad = g
        torch.nn.utils.clip_grad_norm_([p], max_norm)


# --------------------------------------------------------------------------- §8f n2 / n3 consumers
def ewc_penalty_grads(params: Dict[str, Tensor], params_mle: Dict[str, Tensor], fisher: Dict[str, Tensor],
                      lmbda: float):
    """Selective-Amnesia / EWC term of DDPM sa_forget, through autograd exactly as written:
    `_loss = fisher[name] * (param - params_mle[name]) ** 2 ; loss += lmbda * _loss.sum()`
    (DDPM/runners/diffusion.py:424-433).  Returns ({name: d loss / d param}, ewc_loss)."""
    leaves = {n: p.detach().clone().requires_grad_(True) for n, p in params.items()}
    loss = 0.0
    for name, param in leaves.items():
        _loss = fisher[name] * (param - params_mle[name]) ** 2
        loss = loss + lmbda * _loss.sum()
    loss.backward()
    return {n: p.grad for n, p in leaves.items()}, loss.detach()


def proximal_step_(params: Sequence[Tensor], init_params: Sequence[Tensor], ratio: int) -> Tensor:
    """The shrink of SD/train-scripts/proximal_gradient.py:151-183 on CPU tensors (in place):
    threshold = -topk(-|theta - theta0|, ratio)[0][-1], then soft-threshold theta - theta0."""
    n_params = sum(p.numel() for p in params)
    flat = torch.zeros(n_params)
    flat_init = torch.cat([p.reshape(-1) for p in init_params])
    cnt = 0
    for p in params:
        flat[cnt:cnt + p.numel()] = p.reshape(-1)
        cnt += p.numel()
    flat -= flat_init
    flat.abs_().neg_()
    threshold = -torch.topk(flat, ratio)[0][-1]
    for p, init_param in zip(params, init_params):
        p -= init_param
        larger = p > threshold
        smaller = p < -threshold
        between = ~(larger | smaller)
        p[larger] -= threshold
        p[smaller] += threshold
        p[between] = 0
        p += init_param
    return threshold
