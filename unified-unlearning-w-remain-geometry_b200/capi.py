"""ctypes binding of libsfron_b200.so (include/sfron_b200.h).

This is the ONLY compute path of the package: there is no eager / CPU fallback.  If the
shared library is missing, `load()` raises; if a call is made without an sm_100 GPU, the
library returns SFR_ERR_NO_DEVICE and the wrapper raises `SfrError`.

Torch is used here purely as the owner of device memory and streams: every wrapper turns
tensors into raw device pointers + element counts and passes torch's CURRENT stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsfron_b200.so")

# ---- constants mirrored from include/sfron_b200.h ---------------------------------------
ABI_VERSION = 3
OK, ERR_NULL, ERR_ALIGN, ERR_ARG, ERR_NO_DEVICE = 0, -1, -2, -3, -4
F32, BF16 = 0, 1
KEY_ABS, KEY_RATIO, KEY_ABSDIFF = 0, 1, 2
SELECT_BINS0, SELECT_BINS1 = 32768, 65536
SELECT_BINS_ALLOC = 65536 + 128         # u64 words of a bins buffer (histogram + the scan's partials / ticket)
MAX_THRESHOLDS = 8
MAX_PEERS = 8
XP_P2P, XP_MULTIMEM, XP_TMA = 1, 2, 3
OPT_SGD, OPT_ADAM, OPT_ADAMW = 0, 1, 2
EMA_NONE, EMA_DDPM, EMA_DIT, EMA_SLOWFAST = 0, 1, 2, 3
F_MASK, F_MASK_AFTER_CLIP, F_ZERO_GRAD, F_SGD_FIRST_STEP, F_WRITE_BF16, F_REUSE_CONSTS = 1, 2, 4, 8, 16, 32

EXPORTED_SYMBOLS = (
    "sfr_abi_version", "sfr_error_string", "sfr_device_info", "sfr_fisher_accum", "sfr_grad_accum",
    "sfr_ratio_mask", "sfr_ratio_mask_multi", "sfr_select_init", "sfr_select_hist", "sfr_select_hist1_mask",
    "sfr_select_scan", "sfr_select_scratch_elems", "sfr_select_apply", "sfr_masked_sumsq",
    "sfr_fused_update", "sfr_clipped_update", "sfr_ema_update", "sfr_gather_segments",
    "sfr_ewc_penalty", "sfr_select_threshold_value", "sfr_soft_threshold",
    "sfr_peer_pad_bytes", "sfr_peer_barrier", "sfr_peer_reduce", "sfr_peer_fused_update", "sfr_peer_broadcast",
)


class SfrError(RuntimeError):
    def __init__(self, code: int, where: str, message: str):
        super().__init__(f"{where}: {message} (code {code})")
        self.code = code
        _seen_devices.clear()          # a wrapper that fails half-way leaves no device record behind


class UpdateArgs(C.Structure):
    """struct sfr_update_args"""
    _fields_ = [
        ("opt", C.c_int32), ("ema_mode", C.c_int32), ("flags", C.c_uint32), ("g_dtype", C.c_int32),
        ("step", C.c_int64),
        ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
        ("weight_decay", C.c_double), ("momentum", C.c_double), ("dampening", C.c_double),
        ("ema_a", C.c_double), ("clip_max_norm", C.c_double),
        ("lr_table_dev", C.c_void_p), ("lr_index_dev", C.c_void_p),
    ]


class SelectState(C.Structure):
    """struct sfr_select_state (device-resident; this mirror is for read-back / tests)"""
    _fields_ = [
        ("k", C.c_ulonglong), ("k_in_bin", C.c_ulonglong), ("count_gt", C.c_ulonglong),
        ("count_eq", C.c_ulonglong), ("tie_budget", C.c_ulonglong),
        ("prefix", C.c_uint32), ("thr_key", C.c_uint32), ("select_all", C.c_uint32),
        ("select_none", C.c_uint32), ("reserved", C.c_ulonglong * 3),
    ]


class PeerBuf(C.Structure):
    """struct sfr_peer_buf: one symmetric buffer as this rank sees it"""
    _fields_ = [("ptrs", C.c_void_p * MAX_PEERS), ("multicast", C.c_void_p)]


class PeerGeom(C.Structure):
    """struct sfr_peer_geom"""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("lo", C.c_int64), ("n_local", C.c_int64)]


SELECT_STATE_BYTES = C.sizeof(SelectState)
assert SELECT_STATE_BYTES == 80, SELECT_STATE_BYTES

_lib: Optional[C.CDLL] = None


def load(path: Optional[str] = None) -> C.CDLL:
    """dlopen the C-ABI library and declare every prototype.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: build it with `python __graft_entry__.py build` "
            "(sfron_b200 has no CPU or eager fallback)")
    lib = C.CDLL(path)
    vp, i64, i32, f32, f64, u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_double, C.c_ulonglong
    lib.sfr_abi_version.restype = C.c_int
    lib.sfr_abi_version.argtypes = []
    lib.sfr_error_string.restype = C.c_char_p
    lib.sfr_error_string.argtypes = [C.c_int]
    lib.sfr_device_info.restype = C.c_int
    lib.sfr_device_info.argtypes = [C.POINTER(C.c_int)] * 3
    lib.sfr_fisher_accum.restype = C.c_int
    lib.sfr_fisher_accum.argtypes = [vp, vp, C.c_int, i64, i64, i64, f32, vp, f32, vp]
    lib.sfr_grad_accum.restype = C.c_int
    lib.sfr_grad_accum.argtypes = [vp, vp, C.c_int, i64, vp, f32, vp]
    lib.sfr_ratio_mask.restype = C.c_int
    lib.sfr_ratio_mask.argtypes = [vp, vp, i64, f32, f32, vp, vp, vp]
    lib.sfr_ratio_mask_multi.restype = C.c_int
    lib.sfr_ratio_mask_multi.argtypes = [vp, vp, i64, C.POINTER(f32), C.c_int, f32, vp, i64, vp, vp]
    lib.sfr_select_init.restype = C.c_int
    lib.sfr_select_init.argtypes = [vp, vp, u64, vp]
    lib.sfr_select_hist.restype = C.c_int
    lib.sfr_select_hist.argtypes = [vp, vp, C.c_int, f32, i64, C.c_int, vp, vp, vp, vp]
    lib.sfr_select_hist1_mask.restype = C.c_int
    lib.sfr_select_hist1_mask.argtypes = [vp, vp, C.c_int, f32, i64, vp, vp, vp, vp, vp]
    lib.sfr_select_scan.restype = C.c_int
    lib.sfr_select_scan.argtypes = [C.c_int, vp, vp, vp]
    lib.sfr_select_scratch_elems.restype = i64
    lib.sfr_select_scratch_elems.argtypes = [i64]
    lib.sfr_select_apply.restype = C.c_int
    lib.sfr_select_apply.argtypes = [vp, vp, C.c_int, f32, i64, vp, vp, vp, vp, vp]
    lib.sfr_masked_sumsq.restype = C.c_int
    lib.sfr_masked_sumsq.argtypes = [vp, C.c_int, vp, i64, vp, vp]
    lib.sfr_fused_update.restype = C.c_int
    lib.sfr_fused_update.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, C.POINTER(UpdateArgs), vp, vp, vp, vp]
    lib.sfr_clipped_update.restype = C.c_int
    lib.sfr_clipped_update.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, C.POINTER(UpdateArgs), vp, vp, vp]
    lib.sfr_ema_update.restype = C.c_int
    lib.sfr_ema_update.argtypes = [vp, vp, i64, C.c_int, f64, vp]
    lib.sfr_gather_segments.restype = C.c_int
    lib.sfr_gather_segments.argtypes = [vp, vp, vp, vp, i32, C.c_int, i64, vp]
    lib.sfr_ewc_penalty.restype = C.c_int
    lib.sfr_ewc_penalty.argtypes = [vp, vp, vp, vp, i64, f32, vp, vp]
    lib.sfr_select_threshold_value.restype = C.c_int
    lib.sfr_select_threshold_value.argtypes = [vp, vp, vp]
    lib.sfr_soft_threshold.restype = C.c_int
    lib.sfr_soft_threshold.argtypes = [vp, vp, i64, vp, vp]
    lib.sfr_peer_pad_bytes.restype = i64
    lib.sfr_peer_pad_bytes.argtypes = []
    lib.sfr_peer_barrier.restype = C.c_int
    lib.sfr_peer_barrier.argtypes = [C.POINTER(PeerBuf), C.c_int, C.c_int, vp, vp, C.c_int, C.c_uint64, vp]
    lib.sfr_peer_reduce.restype = C.c_int
    lib.sfr_peer_reduce.argtypes = [C.POINTER(PeerBuf), C.c_int, C.POINTER(PeerGeom), C.c_int, C.c_int, vp, vp, vp,
                                    vp, f32, C.c_int, vp]
    lib.sfr_peer_fused_update.restype = C.c_int
    lib.sfr_peer_fused_update.argtypes = [vp, vp, C.POINTER(PeerBuf), C.c_int, C.c_int, C.c_int, vp, vp, vp, vp,
                                          C.POINTER(PeerBuf), C.POINTER(PeerBuf), C.c_int, C.POINTER(PeerGeom),
                                          C.POINTER(UpdateArgs), vp, vp, vp, C.c_int, vp]
    lib.sfr_peer_broadcast.restype = C.c_int
    lib.sfr_peer_broadcast.argtypes = [vp, C.POINTER(PeerBuf), C.c_int, C.POINTER(PeerGeom), C.c_int, vp]
    if lib.sfr_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libsfron_b200 ABI {lib.sfr_abi_version()} != binding {ABI_VERSION}")
    if path == LIB_PATH:
        _lib = lib
    return lib


def _check(code: int, where: str) -> None:
    if code != OK:
        raise SfrError(code, where, load().sfr_error_string(code).decode())


# Devices of the tensors handed to `_ptr` since the last launch.  Every wrapper evaluates its `_ptr(...)` arguments
# left to right and `_stream()` last, so `_stream()` sees exactly the tensors of its own call: it insists that they
# share ONE device and returns that device's current stream.  The library (DeviceScope, csrc/api.cu) launches on
# the device that owns the buffers, so a `HotPath(device="cuda:1")` works whatever the current device is.
_seen_devices: list = []


def _stream() -> int:
    devs = set(_seen_devices)
    _seen_devices.clear()
    if len(devs) > 1:
        raise SfrError(ERR_ARG, "launch", f"tensors of one call live on different devices: {sorted(map(str, devs))}")
    if devs:
        return torch.cuda.current_stream(devs.pop()).cuda_stream
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor], dtype=None, what: str = "tensor") -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise SfrError(ERR_NO_DEVICE, what, "expected a CUDA tensor: this path has no CPU implementation")
    if not t.is_contiguous():
        raise SfrError(ERR_ARG, what, "expected a contiguous tensor")
    if dtype is not None and t.dtype not in (dtype if isinstance(dtype, tuple) else (dtype,)):
        raise SfrError(ERR_ARG, what, f"expected dtype {dtype}, got {t.dtype}")
    _seen_devices.append(t.device)
    return t.data_ptr()


def _gdtype(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise SfrError(ERR_ARG, "gradient", f"gradient dtype {t.dtype} not supported (fp32 | bf16)")


_MASK_DTYPES = (torch.uint8, torch.bool)


def device_info():
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    _check(load().sfr_device_info(C.byref(sm), C.byref(ma), C.byref(mi)), "sfr_device_info")
    return sm.value, ma.value, mi.value


# ---- K1 ------------------------------------------------------------------------------------
def fisher_accum(acc: torch.Tensor, g: torch.Tensor, divisor: float, *,
                 clip_sumsq: Optional[torch.Tensor] = None, clip_max_norm: float = 0.0) -> None:
    """acc += g**2 / divisor over a flat shard; `g` may be [rows, n] (per-sample FIM)."""
    n = acc.numel()
    if g.dim() == 2:
        rows, stride = g.shape[0], g.stride(0)
        if g.shape[1] != n or g.stride(1) != 1:
            raise SfrError(ERR_ARG, "fisher_accum", "g must be [rows, n] with unit inner stride")
        if not g.is_cuda:
            raise SfrError(ERR_NO_DEVICE, "fisher_accum", "expected a CUDA tensor")
        gptr = g.data_ptr()
        _seen_devices.append(g.device)
    else:
        if g.numel() != n:
            raise SfrError(ERR_ARG, "fisher_accum", f"g has {g.numel()} elements, acc has {n}")
        rows, stride, gptr = 1, n, _ptr(g, what="g")
    _check(load().sfr_fisher_accum(_ptr(acc, torch.float32, "acc"), gptr, _gdtype(g), rows, stride, n,
                                   float(divisor), _ptr(clip_sumsq, torch.float64, "clip_sumsq"),
                                   float(clip_max_norm), _stream()), "sfr_fisher_accum")


def grad_accum(acc: torch.Tensor, g: torch.Tensor, *, clip_sumsq: Optional[torch.Tensor] = None,
               clip_max_norm: float = 0.0) -> None:
    """acc += g [* clip coefficient]: SalUn's accumulation of (clipped) batch gradients, one pass."""
    if g.numel() != acc.numel():
        raise SfrError(ERR_ARG, "grad_accum", f"g has {g.numel()} elements, acc has {acc.numel()}")
    _check(load().sfr_grad_accum(_ptr(acc, torch.float32, "acc"), _ptr(g, what="g"), _gdtype(g), acc.numel(),
                                 _ptr(clip_sumsq, torch.float64, "clip_sumsq"), float(clip_max_norm), _stream()),
           "sfr_grad_accum")


# ---- K2a -----------------------------------------------------------------------------------
def ratio_mask(ff: torch.Tensor, rf: torch.Tensor, threshold: float, mask: torch.Tensor,
               zero_count: Optional[torch.Tensor] = None, eps: float = 1e-15) -> None:
    n = ff.numel()
    if rf.numel() != n or mask.numel() != n:
        raise SfrError(ERR_ARG, "ratio_mask", "size mismatch")
    _check(load().sfr_ratio_mask(_ptr(ff, torch.float32, "ff"), _ptr(rf, torch.float32, "rf"), n,
                                 float(threshold), float(eps), _ptr(mask, _MASK_DTYPES, "mask"),
                                 _ptr(zero_count, torch.int64, "zero_count"), _stream()), "sfr_ratio_mask")


def ratio_mask_multi(ff: torch.Tensor, rf: torch.Tensor, thresholds: Sequence[float], masks: torch.Tensor,
                     zero_counts: Optional[torch.Tensor] = None, eps: float = 1e-15) -> None:
    """masks: [T, stride] uint8/bool with stride >= n and stride % 16 == 0."""
    n, t = ff.numel(), len(thresholds)
    if masks.dim() != 2 or masks.shape[0] != t or masks.shape[1] < n:
        raise SfrError(ERR_ARG, "ratio_mask_multi", "masks must be [len(thresholds), >= n]")
    th = (C.c_float * t)(*[float(x) for x in thresholds])
    _check(load().sfr_ratio_mask_multi(_ptr(ff, torch.float32, "ff"), _ptr(rf, torch.float32, "rf"), n, th, t,
                                       float(eps), _ptr(masks, _MASK_DTYPES, "masks"), masks.stride(0),
                                       _ptr(zero_counts, torch.int64, "zero_counts"), _stream()),
           "sfr_ratio_mask_multi")


# ---- K2b -----------------------------------------------------------------------------------
def select_init(state: torch.Tensor, bins: torch.Tensor, k: int) -> None:
    if state.numel() * state.element_size() < SELECT_STATE_BYTES or bins.numel() < SELECT_BINS_ALLOC:
        raise SfrError(ERR_ARG, "select_init", "state / bins buffers too small")
    _check(load().sfr_select_init(_ptr(state, torch.int64, "state"), _ptr(bins, torch.int64, "bins"),
                                  int(k), _stream()), "sfr_select_init")


def select_hist(a: torch.Tensor, b: Optional[torch.Tensor], key_mode: int, pass_: int,
                state: torch.Tensor, bins: torch.Tensor, scratch: Optional[torch.Tensor] = None,
                eps: float = 1e-15, mask: Optional[torch.Tensor] = None) -> None:
    """`mask` (pass 1 only): also write the provisional mask there, so that `select_apply` on the same mask
    finishes from the staged candidates without a third pass over the vector."""
    if scratch is not None and scratch.numel() < select_scratch_elems(a.numel()):
        raise SfrError(ERR_ARG, "select_hist", "scratch too small")
    if mask is not None:
        if pass_ != 1 or mask.numel() != a.numel():
            raise SfrError(ERR_ARG, "select_hist", "mask is a pass-1 output of the same length as the keys")
        _check(load().sfr_select_hist1_mask(_ptr(a, torch.float32, "a"), _ptr(b, torch.float32, "b"), key_mode,
                                            float(eps), a.numel(), _ptr(state, torch.int64, "state"),
                                            _ptr(bins, torch.int64, "bins"), _ptr(scratch, torch.int64, "scratch"),
                                            _ptr(mask, _MASK_DTYPES, "mask"), _stream()), "sfr_select_hist1_mask")
        return
    _check(load().sfr_select_hist(_ptr(a, torch.float32, "a"), _ptr(b, torch.float32, "b"), key_mode,
                                  float(eps), a.numel(), pass_, _ptr(state, torch.int64, "state"),
                                  _ptr(bins, torch.int64, "bins"), _ptr(scratch, torch.int64, "scratch"),
                                  _stream()), "sfr_select_hist")


def select_scan(pass_: int, state: torch.Tensor, bins: torch.Tensor) -> None:
    _check(load().sfr_select_scan(pass_, _ptr(state, torch.int64, "state"), _ptr(bins, torch.int64, "bins"),
                                  _stream()), "sfr_select_scan")


def select_scratch_elems(n: int) -> int:
    return int(load().sfr_select_scratch_elems(int(n)))


def select_apply(a: torch.Tensor, b: Optional[torch.Tensor], key_mode: int, state: torch.Tensor,
                 tie_base: Optional[torch.Tensor], scratch: torch.Tensor, mask: torch.Tensor,
                 eps: float = 1e-15) -> None:
    if scratch.numel() < select_scratch_elems(a.numel()) or mask.numel() != a.numel():
        raise SfrError(ERR_ARG, "select_apply", "scratch too small or mask size mismatch")
    _check(load().sfr_select_apply(_ptr(a, torch.float32, "a"), _ptr(b, torch.float32, "b"), key_mode,
                                   float(eps), a.numel(), _ptr(state, torch.int64, "state"),
                                   _ptr(tie_base, torch.int64, "tie_base"), _ptr(scratch, torch.int64, "scratch"),
                                   _ptr(mask, _MASK_DTYPES, "mask"), _stream()), "sfr_select_apply")


def read_select_state(state: torch.Tensor) -> SelectState:
    """Synchronising read-back of the device select state (tests, multi-GPU tie bases)."""
    raw = state.detach().cpu().contiguous().view(torch.uint8).numpy().tobytes()[:SELECT_STATE_BYTES]
    return SelectState.from_buffer_copy(raw)


# ---- clip norm + K3 ------------------------------------------------------------------------
def masked_sumsq(g: torch.Tensor, mask: Optional[torch.Tensor], out: torch.Tensor) -> None:
    """out (device float64 scalar) += sum((g * mask)**2)."""
    if mask is not None and mask.numel() != g.numel():
        raise SfrError(ERR_ARG, "masked_sumsq", "mask size mismatch")
    _check(load().sfr_masked_sumsq(_ptr(g, what="g"), _gdtype(g), _ptr(mask, _MASK_DTYPES, "mask"), g.numel(),
                                   _ptr(out, torch.float64, "out"), _stream()), "sfr_masked_sumsq")


def fused_update(p: torch.Tensor, g: torch.Tensor, m: Optional[torch.Tensor], v: Optional[torch.Tensor],
                 mask: Optional[torch.Tensor], ema: Optional[torch.Tensor], args: UpdateArgs,
                 clip_sumsq: Optional[torch.Tensor] = None, p_bf16: Optional[torch.Tensor] = None,
                 step_counter: Optional[torch.Tensor] = None, consts_scratch: Optional[torch.Tensor] = None) -> None:
    """step_counter (device int64[1]) + consts_scratch (device uint8[>=128]): replay-safe launch for CUDA
    graphs — the optimizer step lives on the device and advances on every replay."""
    n = p.numel()
    for name, t in (("g", g), ("m", m), ("v", v), ("mask", mask), ("ema", ema), ("p_bf16", p_bf16)):
        if t is not None and t.numel() != n:
            raise SfrError(ERR_ARG, "fused_update", f"{name} has {t.numel()} elements, p has {n}")
    args.g_dtype = _gdtype(g)
    _check(load().sfr_fused_update(_ptr(p, torch.float32, "p"), _ptr(g, what="g"), _ptr(m, torch.float32, "m"),
                                   _ptr(v, torch.float32, "v"), _ptr(mask, _MASK_DTYPES, "mask"),
                                   _ptr(ema, torch.float32, "ema"), _ptr(p_bf16, torch.bfloat16, "p_bf16"), n,
                                   C.byref(args), _ptr(clip_sumsq, torch.float64, "clip_sumsq"),
                                   _ptr(step_counter, torch.int64, "step_counter"),
                                   _ptr(consts_scratch, torch.uint8, "consts_scratch"), _stream()),
           "sfr_fused_update")


def clipped_update(p: torch.Tensor, g: torch.Tensor, m: Optional[torch.Tensor], v: Optional[torch.Tensor],
                   mask: Optional[torch.Tensor], ema: Optional[torch.Tensor], args: UpdateArgs, sumsq: torch.Tensor,
                   p_bf16: Optional[torch.Tensor] = None, step_counter: Optional[torch.Tensor] = None) -> None:
    """Clip norm + fused update in one cooperative launch (small vectors); leaves the squared norm in `sumsq`."""
    n = p.numel()
    for name, t in (("g", g), ("m", m), ("v", v), ("mask", mask), ("ema", ema), ("p_bf16", p_bf16)):
        if t is not None and t.numel() != n:
            raise SfrError(ERR_ARG, "clipped_update", f"{name} has {t.numel()} elements, p has {n}")
    args.g_dtype = _gdtype(g)
    _check(load().sfr_clipped_update(_ptr(p, torch.float32, "p"), _ptr(g, what="g"), _ptr(m, torch.float32, "m"),
                                     _ptr(v, torch.float32, "v"), _ptr(mask, _MASK_DTYPES, "mask"),
                                     _ptr(ema, torch.float32, "ema"), _ptr(p_bf16, torch.bfloat16, "p_bf16"), n,
                                     C.byref(args), _ptr(sumsq, torch.float64, "sumsq"),
                                     _ptr(step_counter, torch.int64, "step_counter"), _stream()),
           "sfr_clipped_update")


def ema_update(p: torch.Tensor, ema: torch.Tensor, ema_mode: int, ema_a: float) -> None:
    if p.numel() != ema.numel():
        raise SfrError(ERR_ARG, "ema_update", "size mismatch")
    _check(load().sfr_ema_update(_ptr(p, torch.float32, "p"), _ptr(ema, torch.float32, "ema"), p.numel(),
                                 ema_mode, float(ema_a), _stream()), "sfr_ema_update")


def gather_segments(flat: torch.Tensor, srcs: torch.Tensor, offsets: torch.Tensor, sizes: torch.Tensor,
                    src_dtype: int, total: int) -> None:
    _check(load().sfr_gather_segments(_ptr(flat, torch.float32, "flat"), _ptr(srcs, torch.int64, "srcs"),
                                      _ptr(offsets, torch.int64, "offsets"), _ptr(sizes, torch.int64, "sizes"),
                                      srcs.numel(), src_dtype, int(total), _stream()), "sfr_gather_segments")


# ---- consumers next to the path -------------------------------------------
def ewc_penalty(p: torch.Tensor, p_star: torch.Tensor, fisher: torch.Tensor, g: torch.Tensor, lmbda: float,
                penalty: Optional[torch.Tensor] = None) -> None:
    """g += (lmbda*F) * (2*(p - p_star)); penalty (device float64) += lmbda * sum(F * (p - p_star)**2)."""
    n = p.numel()
    if not (p_star.numel() == fisher.numel() == g.numel() == n):
        raise SfrError(ERR_ARG, "ewc_penalty", "size mismatch")
    _check(load().sfr_ewc_penalty(_ptr(p, torch.float32, "p"), _ptr(p_star, torch.float32, "p_star"),
                                  _ptr(fisher, torch.float32, "fisher"), _ptr(g, torch.float32, "g"), n,
                                  float(lmbda), _ptr(penalty, torch.float64, "penalty"), _stream()),
           "sfr_ewc_penalty")


def select_threshold_value(state: torch.Tensor, out: torch.Tensor) -> None:
    _check(load().sfr_select_threshold_value(_ptr(state, torch.int64, "state"), _ptr(out, torch.float32, "out"),
                                             _stream()), "sfr_select_threshold_value")


def soft_threshold(p: torch.Tensor, p0: torch.Tensor, threshold: torch.Tensor) -> None:
    if p.numel() != p0.numel():
        raise SfrError(ERR_ARG, "soft_threshold", "size mismatch")
    _check(load().sfr_soft_threshold(_ptr(p, torch.float32, "p"), _ptr(p0, torch.float32, "p0"), p.numel(),
                                     _ptr(threshold, torch.float32, "threshold"), _stream()), "sfr_soft_threshold")


# ---- cross-GPU exchange fused with the kernels (csrc/peer.cu) -----------------------------------------
def peer_buf(ptrs: Sequence[int], multicast: int = 0) -> PeerBuf:
    """struct sfr_peer_buf from the mapped device addresses of one symmetric buffer (ptrs[r] = rank r's copy)."""
    if not 1 <= len(ptrs) <= MAX_PEERS:
        raise SfrError(ERR_ARG, "peer_buf", f"1..{MAX_PEERS} ranks supported, got {len(ptrs)}")
    b = PeerBuf()
    for r, p in enumerate(ptrs):
        b.ptrs[r] = int(p)
    b.multicast = int(multicast) or None
    return b


def peer_buf_offset(b: PeerBuf, world: int, byte_offset: int) -> PeerBuf:
    """The same symmetric buffer seen from `byte_offset` (a multiple of 16) onwards."""
    out = PeerBuf()
    for r in range(world):
        out.ptrs[r] = b.ptrs[r] + byte_offset
    out.multicast = (b.multicast + byte_offset) if b.multicast else None
    return out


def peer_pad_bytes() -> int:
    return int(load().sfr_peer_pad_bytes())


def peer_barrier(pad: PeerBuf, world: int, rank: int, vals: Optional[torch.Tensor] = None,
                 sums: Optional[torch.Tensor] = None, timeout_ns: int = 0) -> None:
    """Cross-GPU barrier on the current stream; `vals` (device float64[k<=8]) are summed over ranks, in rank
    order, into `sums` on every rank."""
    nvals = 0 if vals is None else vals.numel()
    if nvals and (sums is None or sums.numel() < nvals):
        raise SfrError(ERR_ARG, "peer_barrier", "sums must hold as many doubles as vals")
    _check(load().sfr_peer_barrier(C.byref(pad), world, rank, _ptr(vals, torch.float64, "vals"),
                                   _ptr(sums, torch.float64, "sums"), nvals, int(timeout_ns), _stream()),
           "sfr_peer_barrier")


def peer_reduce(g: PeerBuf, g_dtype: torch.dtype, geom: PeerGeom, transport: int, average: bool, *,
                g_red: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
                sumsq: Optional[torch.Tensor] = None, fisher: Optional[torch.Tensor] = None,
                fisher_divisor: float = 1.0, max_ctas: int = 0) -> None:
    for name, t in (("g_red", g_red), ("mask", mask), ("fisher", fisher)):
        if t is not None and t.numel() != geom.n_local:
            raise SfrError(ERR_ARG, "peer_reduce", f"{name} has {t.numel()} elements, the shard has {geom.n_local}")
    gd = F32 if g_dtype == torch.float32 else BF16 if g_dtype == torch.bfloat16 else None
    if gd is None:
        raise SfrError(ERR_ARG, "peer_reduce", f"gradient dtype {g_dtype} not supported (fp32 | bf16)")
    _check(load().sfr_peer_reduce(C.byref(g), gd, C.byref(geom), transport, int(bool(average)),
                                  _ptr(g_red, torch.float32, "g_red"), _ptr(mask, _MASK_DTYPES, "mask"),
                                  _ptr(sumsq, torch.float64, "sumsq"), _ptr(fisher, torch.float32, "fisher"),
                                  float(fisher_divisor), int(max_ctas), _stream()), "sfr_peer_reduce")


def peer_fused_update(p: torch.Tensor, geom: PeerGeom, args: UpdateArgs, *, g_red: Optional[torch.Tensor] = None,
                      g: Optional[PeerBuf] = None, g_dtype: torch.dtype = torch.float32, g_transport: int = XP_P2P,
                      average: bool = True, m: Optional[torch.Tensor] = None, v: Optional[torch.Tensor] = None,
                      mask: Optional[torch.Tensor] = None, ema: Optional[torch.Tensor] = None,
                      bc_f32: Optional[PeerBuf] = None, bc_bf16: Optional[PeerBuf] = None,
                      bc_transport: int = XP_P2P, clip_sumsq: Optional[torch.Tensor] = None,
                      step_counter: Optional[torch.Tensor] = None,
                      consts_scratch: Optional[torch.Tensor] = None, max_ctas: int = 0) -> None:
    n = geom.n_local
    for name, t in (("p", p), ("g_red", g_red), ("m", m), ("v", v), ("mask", mask), ("ema", ema)):
        if t is not None and t.numel() != n:
            raise SfrError(ERR_ARG, "peer_fused_update", f"{name} has {t.numel()} elements, the shard has {n}")
    if (g is None) == (g_red is None):
        raise SfrError(ERR_ARG, "peer_fused_update", "exactly one gradient source: g_red (local) or g (peers)")
    gd = F32 if g_dtype == torch.float32 else BF16
    args.g_dtype = gd
    _check(load().sfr_peer_fused_update(
        _ptr(p, torch.float32, "p"), _ptr(g_red, torch.float32, "g_red"), C.byref(g) if g is not None else None, gd,
        g_transport, int(bool(average)), _ptr(m, torch.float32, "m"), _ptr(v, torch.float32, "v"),
        _ptr(mask, _MASK_DTYPES, "mask"), _ptr(ema, torch.float32, "ema"),
        C.byref(bc_f32) if bc_f32 is not None else None, C.byref(bc_bf16) if bc_bf16 is not None else None,
        bc_transport, C.byref(geom), C.byref(args), _ptr(clip_sumsq, torch.float64, "clip_sumsq"),
        _ptr(step_counter, torch.int64, "step_counter"), _ptr(consts_scratch, torch.uint8, "consts_scratch"),
        int(max_ctas), _stream()), "sfr_peer_fused_update")


def peer_broadcast(src: torch.Tensor, dst: PeerBuf, geom: PeerGeom, transport: int) -> None:
    if src.numel() != geom.n_local or src.element_size() not in (2, 4):
        raise SfrError(ERR_ARG, "peer_broadcast", "src must be this rank's shard of 2- or 4-byte elements")
    _check(load().sfr_peer_broadcast(_ptr(src, what="src"), C.byref(dst), src.element_size(), C.byref(geom), transport,
                                     _stream()), "sfr_peer_broadcast")
