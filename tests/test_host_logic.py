"""CPU tests of the host side: the C-ABI library loads and exports what include/sfron_b200.h declares,
flat layout / shard arithmetic, reference file formats, and the no-fallback guarantees."""
import os
import re

import pytest
import torch
from hypothesis import given, settings, strategies as st

from conftest import ROOT, load_golden

import sfron_b200 as sfr
from sfron_b200 import capi, formats
from sfron_b200.dist import scan_from_top, tie_bases


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "sfron_b200.h")).read()
    declared = set(re.findall(r"^SFR_API [\w\s\*]+?\b(sfr_\w+)\(", header, flags=re.M))
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    lib = capi.load()
    for sym in declared:
        assert hasattr(lib, sym)
    assert lib.sfr_abi_version() == capi.ABI_VERSION
    assert b"no CPU fallback" in lib.sfr_error_string(capi.ERR_NO_DEVICE)


def test_update_args_struct_matches_header():
    header = open(os.path.join(ROOT, "include", "sfron_b200.h")).read()
    body = re.search(r"typedef struct sfr_update_args \{(.*?)\} sfr_update_args;", header, flags=re.S).group(1)
    fields = re.findall(r"^\s*(?:int32_t|uint32_t|int64_t|double|const double\*|const long long\*)\s+(\w+);", body,
                        flags=re.M)
    assert fields == [f[0] for f in capi.UpdateArgs._fields_]
    import ctypes
    assert ctypes.sizeof(capi.UpdateArgs) == 4 * 4 + 8 + 9 * 8 + 2 * 8        # no padding surprises
    body = re.search(r"typedef struct sfr_select_state \{(.*?)\} sfr_select_state;", header, flags=re.S).group(1)
    fields = re.findall(r"^\s*(?:unsigned long long|uint32_t)\s+(\w+)", body, flags=re.M)
    assert fields == [f[0] for f in capi.SelectState._fields_]
    for struct, cls in (("sfr_peer_buf", capi.PeerBuf), ("sfr_peer_geom", capi.PeerGeom)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), header, flags=re.S).group(1)
        fields = re.findall(r"^\s*(?:void\*|int32_t|int64_t)\s+(\w+)", body, flags=re.M)
        assert fields == [f[0] for f in cls._fields_], struct


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a GPU")
def test_no_cpu_fallback_anywhere():
    with pytest.raises(capi.SfrError) as e:
        capi.device_info()
    assert e.value.code == capi.ERR_NO_DEVICE
    with pytest.raises(capi.SfrError):
        sfr.HotPath(16, "cpu", sfr.OptConfig())
    with pytest.raises(capi.SfrError):
        capi.fisher_accum(torch.zeros(16), torch.zeros(16), 1.0)
    with pytest.raises(capi.SfrError):
        capi.masked_sumsq(torch.zeros(16), None, torch.zeros(1, dtype=torch.float64))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "unified-unlearning-w-remain-geometry_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_flat_layout_offsets_and_views():
    L = sfr.FlatLayout([("a.weight", (3, 4)), ("a.bias", (4,)), ("s", ()), ("b", (2, 2, 2))])
    assert L.numel == 12 + 4 + 1 + 8 and L.names == ["a.weight", "a.bias", "s", "b"]
    assert [s.offset for s in L] == [0, 12, 16, 17]
    assert [s.name for s in L.misaligned(4)] == ["b"]
    flat = torch.arange(L.numel, dtype=torch.float32)
    v = L.views(flat)
    assert v["a.bias"].tolist() == [12, 13, 14, 15] and v["s"].shape == () and v["b"].shape == (2, 2, 2)
    assert torch.equal(L.flatten(v), flat)
    v["s"].fill_(-1)                                  # views alias the flat vector
    assert flat[16] == -1
    with pytest.raises(ValueError):
        sfr.FlatLayout([("x", (1,)), ("x", (2,))])


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 10_000_000), world=st.integers(1, 16), align=st.sampled_from([1, 4, 16]))
def test_shard_bounds_partition(n, world, align):
    bounds = [sfr.shard_bounds(n, world, r, align) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == n
    for (lo, hi), (lo2, _) in zip(bounds, bounds[1:]):
        assert hi == lo2 and lo <= hi
    for lo, hi in bounds:
        assert lo % align == 0 or lo == n
    sizes = [hi - lo for lo, hi in bounds]
    assert max(sizes) - min(sizes) < 2 * align      # one unit, plus the ragged last unit


def test_formats_round_trip_reference_files(tmp_path):
    fx = load_golden("dit_ratio_mask.pt")
    all_names = list(fx["forget"].keys())
    layout = sfr.FlatLayout([(n, tuple(t.shape)) for n, t in fx["forget"].items() if torch.is_tensor(t)])
    ff = formats.dict_to_flat(layout, fx["forget"])
    back = formats.flat_to_dict(layout, ff, all_names=all_names)
    assert list(back.keys()) == all_names
    for n in all_names:
        if torch.is_tensor(fx["forget"][n]):
            assert torch.equal(back[n], fx["forget"][n]) and back[n].dtype == torch.float32
        else:
            assert back[n] == 0 and isinstance(back[n], int)            # pos_embed placeholder survives
    # files written by us load like the reference's own
    formats.save_fisher(tmp_path / "forget_fisher.pt", layout, ff, all_names=all_names)
    again = torch.load(tmp_path / "forget_fisher.pt", weights_only=False)
    assert all(torch.equal(again[n], fx["forget"][n]) for n in layout.names)
    ref_mask = fx["masks"]["1"]
    flat_mask = formats.load_mask(ref_mask, layout)
    assert flat_mask.dtype == torch.uint8
    d = formats.ratio_mask_to_dict(layout, flat_mask, all_names=all_names)
    assert all(d[n].dtype == torch.bool and torch.equal(d[n], ref_mask[n]) for n in layout.names)
    d64 = formats.topk_mask_to_dict(layout, flat_mask)
    assert all(v.dtype == torch.int64 for v in d64.values())
    assert formats.threshold_tag(1.0) == "1.0" and formats.threshold_tag(1) == "1"
    with pytest.raises(ValueError):
        formats.dict_to_flat(sfr.FlatLayout([("module.pos_embed", (1, 6, 8))]), fx["forget"])


def test_adam_state_dict_loads_into_torch_optimizer():
    model = torch.nn.Linear(5, 3)
    layout = sfr.FlatLayout.from_named_tensors(model.named_parameters())
    m = torch.randn(layout.numel)
    v = torch.rand(layout.numel)
    sd = formats.adam_state_dict(layout, m, v, 7, lr=1e-4)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    opt.load_state_dict(sd)
    st0 = opt.state[model.weight]
    assert torch.equal(st0["exp_avg"].reshape(-1), m[:15]) and float(st0["step"]) == 7
    m2, v2, step = formats.load_adam_state(layout, opt.state_dict())
    assert torch.equal(m2, m) and torch.equal(v2, v) and step == 7


@settings(max_examples=100, deadline=None)
@given(data=st.data())
def test_scan_from_top_matches_bruteforce(data):
    nb = data.draw(st.sampled_from([4, 64, 1024]))
    counts = data.draw(st.lists(st.integers(0, 5), min_size=nb, max_size=nb))
    bins = torch.tensor(counts, dtype=torch.int64)
    total = int(bins.sum())
    want = data.draw(st.integers(0, total + 2))
    b, above = scan_from_top(bins, want)
    if want == 0 or want > total:
        assert b == -1
    else:
        assert above < want <= above + counts[b] and above == sum(counts[b + 1:])
    assert tie_bases([3, 0, 2]) == [0, 3, 3]


def test_format_converters_single_copy_semantics(tmp_path):
    """flat_to_dict hands out views of ONE host copy (never of the caller's buffer); a saved file reloads —
    memory-mapped — into the same dict; dict_to_flat packs bool / int64 / fp32 entries straight into the target."""
    layout = sfr.FlatLayout([("a", (2, 3)), ("b", (5,)), ("c", ())])
    flat = torch.arange(12, dtype=torch.float32)
    d = formats.flat_to_dict(layout, flat, all_names=["frozen", "a", "b", "c"], prefix="module.")
    assert list(d) == ["module.frozen", "module.a", "module.b", "module.c"] and d["module.frozen"] == 0
    d["module.a"][0, 0] = 99.0
    assert flat[0] == 0.0                                           # the caller's vector is not aliased
    assert d["module.a"].untyped_storage().data_ptr() == d["module.b"].untyped_storage().data_ptr()
    path = tmp_path / "forget_fisher.pt"
    torch.save(d, path)
    back = formats.load_file(str(path))
    assert back["module.frozen"] == 0 and torch.equal(back["module.b"], d["module.b"])
    assert torch.equal(formats.load_fisher(str(path), layout, prefix="module."), torch.cat([d[k].reshape(-1) for k in list(d)[1:]]))
    bits = (flat > 3).to(torch.uint8)
    for to_dict, dtype in ((formats.ratio_mask_to_dict, torch.bool), (formats.topk_mask_to_dict, torch.int64)):
        m = to_dict(layout, bits)
        assert all(t.dtype == dtype for t in m.values())
        assert torch.equal(formats.load_mask(m, layout), bits)


def test_cli_flags_match_the_reference():
    """The drop-in command lines accept the reference scripts' flags with the same defaults, types, `required`, `nargs`
    and store_true actions.  The fixture was read statically from the reference's sources
    (tests/golden/make_cli_flags.py); the one deliberate difference is listed here."""
    import argparse
    import json
    import os
    from sfron_b200.methods import dit, masks, sd
    doc = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cli_flags.json")))
    sub = {a.dest: a for a in masks.build_parser()._actions}["family"].choices
    ours = {"dit_forget": dit.forget_parser(), "dit_generate_fisher": dit.generate_fisher_parser(),
            "dit_generate_mask": sub["dit"], "sd_generate_fisher": sd.generate_fisher_parser(),
            "sd_nsfw_removal": sd.nsfw_removal_parser(), "sd_generate_fisher_mask": sub["sd"],
            "ddpm_generate_fisher_mask": sub["ddpm"]}
    # SD/train-scripts/nsfw_removal.py declares --lr with type=int (any `--lr 1e-5` on its command line fails to parse);
    # the default is the float 1e-5 and the value feeds Adam's lr, so the flag is a float here
    deliberate = {("sd_nsfw_removal", "--lr", "type")}
    assert set(ours) == set(doc)
    for name, parser in ours.items():
        mine = {a.option_strings[0]: a for a in parser._actions if a.option_strings and a.dest != "help"}
        theirs = {f["flags"][0]: f for f in doc[name]["flags"]}
        assert set(mine) == set(theirs), (name, set(mine) ^ set(theirs))
        for flag, ref in theirs.items():
            a = mine[flag]
            assert list(a.option_strings) == ref["flags"], (name, flag)
            if ref.get("action") == "store_true":
                assert isinstance(a, argparse._StoreTrueAction), (name, flag)
                continue
            if (name, flag, "type") not in deliberate:
                assert (a.type.__name__ if a.type else None) == ref.get("type"), (name, flag, "type")
            if not (isinstance(ref.get("default"), str) and ref["default"].startswith("<expr>")):
                assert a.default == ref.get("default"), (name, flag, "default", a.default)
            assert bool(a.required) == bool(ref.get("required", False)), (name, flag, "required")
            assert a.nargs == ref.get("nargs"), (name, flag, "nargs")
            if isinstance(ref.get("choices"), list):
                assert list(a.choices) == ref["choices"], (name, flag, "choices")


def test_ctypes_signatures_match_the_header():
    """Every prototype of include/sfron_b200.h against the argtypes / restype capi.py binds it with: same argument count,
    and per argument the same class (pointer, or the scalar's C type).  No call is made."""
    import ctypes as C
    hdr = open(os.path.join(ROOT, "include", "sfron_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = re.findall(r"SFR_API\s+([\w\s\*]+?)\s*\b(sfr_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S)
    assert len(protos) == 26, [p[1] for p in protos]
    scalar = {"int": C.c_int, "int32_t": C.c_int, "int64_t": C.c_int64, "float": C.c_float, "double": C.c_double,
              "unsigned long long": C.c_uint64, "uint64_t": C.c_uint64, "long long": C.c_int64, "unsigned int": C.c_uint, "uint32_t": C.c_uint}

    def klass(decl):
        decl = decl.strip()
        if "*" in decl or decl.startswith("sfr_stream_t"):
            return "ptr"
        words = [w for w in decl.replace("const", " ").split()]
        ty = " ".join(words[:-1]) if len(words) > 1 else words[0]
        return scalar[ty]

    def ctypes_klass(t):
        if t is C.c_void_p or t is C.c_char_p or isinstance(t, type) and issubclass(t, (C._Pointer,)):
            return "ptr"
        return t

    lib = capi.load()
    for ret, name, args in protos:
        fn = getattr(lib, name)
        want = [] if args.strip() in ("", "void") else [klass(a) for a in args.split(",")]
        got = [ctypes_klass(t) for t in (fn.argtypes or [])]
        assert len(got) == len(want), (name, len(got), len(want))
        for i, (g, w) in enumerate(zip(got, want)):
            if w == "ptr":
                assert g == "ptr", (name, i, g)
            else:
                assert g != "ptr" and C.sizeof(g) == C.sizeof(w) and (g in (C.c_float, C.c_double)) == (w in (C.c_float, C.c_double)), (name, i, g, w)
        rk = "ptr" if "*" in ret else scalar[ret.strip()]
        gr = ctypes_klass(fn.restype)
        assert (gr == "ptr") == (rk == "ptr") and (rk == "ptr" or C.sizeof(gr) == C.sizeof(rk)), (name, "restype")


def test_c_abi_from_plain_c(tmp_path):
    """tests/c/abi_consumer.c: the header compiles as pedantic C99, the library links from C, the struct sizes agree with
    the ctypes mirrors, and without a device the compute entry points return SFR_ERR_NO_DEVICE (no host fallback)."""
    import ctypes as C
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    capi.load()
    pkg = os.path.dirname(capi.LIB_PATH)
    exe = str(tmp_path / "abi_consumer")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_consumer.c"), "-o", exe, "-L", pkg,
                    "-l:" + os.path.basename(capi.LIB_PATH), "-Wl,-rpath," + pkg], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert f"abi {capi.ABI_VERSION}" in out
    sizes = dict(zip(re.findall(r"(sfr_\w+) \d+", out), map(int, re.findall(r"sfr_\w+ (\d+)", out))))
    assert sizes == {"sfr_update_args": C.sizeof(capi.UpdateArgs), "sfr_select_state": C.sizeof(capi.SelectState),
                     "sfr_peer_buf": C.sizeof(capi.PeerBuf), "sfr_peer_geom": C.sizeof(capi.PeerGeom)}, out
    if not torch.cuda.is_available():
        assert "no device: info -4 k1 -4 k2a -4" in out and "no CPU fallback" in out


def _bench(argv, env_extra=None, timeout=300):
    import subprocess
    import sys
    env = dict(os.environ, **(env_extra or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + argv, capture_output=True, text=True,
                          env=env, timeout=timeout, cwd=ROOT)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours): exactly one JSON line on stdout with our
    arm's metric / unit / config, the bounded sample named, no GPU work claimed; under a multi-rank launch only rank 0
    works and prints."""
    import json
    r = _bench(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-elems", str(1 << 18)])
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sfron_hot_path_GBps" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32" and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["elements"] == 675_129_632 and d["config"]["elements_timed"] == 1 << 18
    assert d["config"]["bytes_per_element"] == 103 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and str(1 << 18) in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
    # rank 1 of a 2-rank launch: exits 0 without work and without output
    r = _bench(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--ref-elems", str(1 << 18)],
               {"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without a GPU")
def test_bench_our_arm_refuses_to_run_without_a_gpu():
    """No CPU fallback behind the headline number: without a CUDA device our arm stops with a message, it does not print
    a line."""
    r = _bench(["--steps", "1"])
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CUDA device" in (r.stderr + r.stdout)
