"""B200-native SFR-on hot path (Fisher diagonal -> saliency mask -> fused masked update).

The directory name carries the reference's name and is not a Python identifier; import the
package through its alias:  `import sfron_b200`.

Only what the path needs lives here:
  csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/sfron_b200.h)
  capi.py    ctypes binding of libsfron_b200.so  (the ONLY compute path; no CPU fallback)
  flat.py    flat parameter-vector layout, shard partition
  engine.py  host sequencing of the kernels in the reference's order of operations
  formats.py reference Fisher / mask / checkpoint file formats
  dist.py    one-process-per-GPU sharding + NCCL collectives of the path
  methods/   drop-in mirrors of the reference's entry points
"""
from . import capi, flat, formats  # noqa: F401
from .engine import HostGradientFeeder, HotPath, OptConfig  # noqa: F401
from .flat import FlatLayout, FlatParams, shard_bounds  # noqa: F401

__all__ = ["capi", "flat", "formats", "HotPath", "HostGradientFeeder", "OptConfig", "FlatLayout", "FlatParams", "shard_bounds"]
