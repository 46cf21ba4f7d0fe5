import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with `-m gpu` under gpurun")


@pytest.fixture(scope="session", autouse=True)
def _library_is_built():
    """A fresh checkout has no libsfron_b200.so (build artefacts are git-ignored): build it before the first test if
    the sources changed or it is missing (nvcc cross-compiles without a GPU; a no-op when the stamp matches)."""
    import __graft_entry__ as entry
    builder = entry._load_build_module()
    if not builder.is_fresh():
        builder.build()
    yield


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def unflat(vec, names, shapes):
    out, off = {}, 0
    for n in names:
        numel = 1
        for s in shapes[n]:
            numel *= s
        out[n] = vec[off:off + numel].reshape(shapes[n]).clone()
        off += numel
    assert off == vec.numel()
    return out


def bits_equal(a, b):
    """Bit-for-bit equality of two fp32 tensors (distinguishes -0.0 / NaN payloads)."""
    return torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32))


def max_rel_err(a, b, floor=1e-30):
    a, b = a.double(), b.double()
    return ((a - b).abs() / b.abs().clamp_min(floor)).max().item() if a.numel() else 0.0
