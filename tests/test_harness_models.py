"""The harness networks of tools/ (stand-ins for reference models that cannot travel to the GPU box) have
the reference's state-dict names, shapes, ORDER and parameter counts — the flat element order of the hot
path is `named_parameters()` order, so a mask or Fisher file written by either side fits the other.
Compared against the reference modules when /root/reference is present (build container); the recorded
counts are checked everywhere.  CPU only."""
import argparse
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
REF = os.environ.get("SFR_REFERENCE", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")


def _sig(model):
    return [(n, tuple(p.shape)) for n, p in model.named_parameters()]


def test_ddpm_unet_counts():
    from ddpm_unet import DDPMCondUNet
    m = DDPMCondUNet()
    sig = _sig(m)
    assert len(sig) == 334 and sum(p.numel() for p in m.parameters()) == 38_632_323
    assert sig[0][0] == "null_classes_emb" and sig[-1] == ("conv_out.bias", (3,))


@needs_ref
def test_ddpm_unet_matches_reference_module():
    from ddpm_unet import DDPMCondUNet
    import yaml
    sys.path.insert(0, os.path.join(REF, "DDPM"))
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
    try:
        from models.diffusion import Conditional_Model
    finally:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path.remove(os.path.join(REF, "DDPM"))

    def d2n(d):
        ns = argparse.Namespace()
        for k, v in d.items():
            setattr(ns, k, d2n(v) if isinstance(v, dict) else v)
        return ns

    config = d2n(yaml.safe_load(open(os.path.join(REF, "DDPM/configs/cifar10_sfron.yml"))))
    torch.manual_seed(0)
    ref = Conditional_Model(config).eval()
    ours = DDPMCondUNet().eval()
    assert _sig(ours) == _sig(ref)
    ours.load_state_dict(ref.state_dict(), strict=True)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 32, 32, generator=g)
    t = torch.tensor([3.0, 977.0])
    c = torch.tensor([0, 7])
    with torch.no_grad():
        for kw in (dict(mode="train", cond_drop_prob=0.0), dict(mode="test", cond_scale=2.0)):
            a, b = ref(x, t, c, **kw), ours(x, t, c, **kw)
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-4), (kw, (a - b).abs().max())


def test_sd_unet_counts():
    from sd_unet import SDUNet
    with torch.device("meta"):
        m = SDUNet()
    sig = _sig(m)
    assert len(sig) == 686 and sum(p.numel() for p in m.parameters()) == 859_520_964
    xattn = [(n, s) for n, s in sig if "attn2" in n]           # nsfw_removal.py --train_method xattn
    assert len(xattn) == 80 and sum(torch.Size(s).numel() for _, s in xattn) == 43_962_560
    assert sig[0][0] == "time_embed.0.weight" and sig[-1] == ("out.2.bias", (4,))


@needs_ref
def test_sd_unet_matches_reference_module():
    """Names / shapes / order at full size (meta tensors), outputs on a narrow instance of both."""
    import types
    from sd_unet import SDUNet
    if "omegaconf" not in sys.modules:              # the only import of the reference file that is absent here
        oc = types.ModuleType("omegaconf")
        lc = types.ModuleType("omegaconf.listconfig")
        lc.ListConfig = list
        oc.listconfig = lc
        sys.modules["omegaconf"], sys.modules["omegaconf.listconfig"] = oc, lc
    sys.path.insert(0, os.path.join(REF, "SD"))
    try:
        from ldm.modules.diffusionmodules.openaimodel import UNetModel
    finally:
        sys.path.remove(os.path.join(REF, "SD"))
    kw = dict(image_size=32, in_channels=4, out_channels=4, attention_resolutions=[4, 2, 1], num_res_blocks=2,
              channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True, transformer_depth=1,
              context_dim=768, use_checkpoint=False, legacy=False)
    with torch.device("meta"):
        ref_full = UNetModel(model_channels=320, **kw)
        ours_full = SDUNet()
    assert _sig(ours_full) == _sig(ref_full)
    torch.manual_seed(0)
    ref = UNetModel(model_channels=64, **kw).eval()
    ours = SDUNet(model_channels=64).eval()
    assert _sig(ours) == _sig(ref)
    with torch.no_grad():                          # the constructor zeroes every block's last conv
        for p in ref.parameters():
            if not p.any():
                p.normal_(std=0.05)
    ours.load_state_dict(ref.state_dict(), strict=True)
    g = torch.Generator().manual_seed(1)
    z, t, ctx = torch.randn(2, 4, 16, 16, generator=g), torch.tensor([5, 900]), torch.randn(2, 77, 768, generator=g)
    with torch.no_grad():
        a, b = ref(z, t, context=ctx), ours(z, t, ctx)
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-4), (a - b).abs().max()


def test_dit_harness_structure_matches_reference_class():
    """tools/dit_xl2.py at the size of the recorded reference run (tests/golden/dit_scripts.pt was produced by the
    reference's own `DiT` class): same names, shapes and order, frozen `pos_embed` first; full-size counts."""
    from conftest import load_golden
    from dit_xl2 import DiTXL2Harness
    fx = load_golden("dit_scripts.pt")
    tiny = DiTXL2Harness(input_size=8, patch=8, width=8, depth=1, heads=2, classes=10)
    assert [("module." + n, list(p.shape)) for n, p in tiny.named_parameters()] == [(n, fx["shapes"][n]) for n in fx["names"]]
    assert ["module." + n for n, p in tiny.named_parameters() if p.requires_grad] == fx["train_names"]
    with torch.device("meta"):
        full = DiTXL2Harness()
    assert sum(p.numel() for p in full.parameters()) == 675_129_632 and len(list(full.parameters())) == 292
    assert sum(p.numel() for p in full.parameters() if p.requires_grad) == 674_834_720


def test_resnet18_harness_counts():
    from resnet18_cifar import ResNet18Harness
    m = ResNet18Harness()
    assert len(_sig(m)) == 62 and sum(p.numel() for p in m.parameters()) == 11_173_962


@needs_ref
def test_resnet18_harness_matches_reference_module():
    import importlib.util
    from resnet18_cifar import ResNet18Harness
    spec = importlib.util.spec_from_file_location("ref_resnet", os.path.join(REF, "Classification/models/resnet.py"))
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    torch.manual_seed(0)
    ref, ours = ref_mod.ResNet18(10).eval(), ResNet18Harness().eval()
    assert _sig(ours) == _sig(ref)
    ours.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        assert torch.allclose(ref(x), ours(x), rtol=1e-5, atol=1e-5)


@needs_ref
def test_dit_harness_matches_reference_class_outputs():
    """The reference's own `DiT` class (DiT/models.py) needs timm's PatchEmbed / Attention / Mlp; with the stand-ins
    that tests/golden/make_golden.py uses for them, a 2-block model of both implementations agrees on names, shapes,
    order and — same weights — outputs (embedders, adaLN modulation, final layer, unpatchify, fixed pos_embed)."""
    import importlib.util
    from dit_xl2 import DiTXL2Harness
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    before = set(sys.modules)
    saved_models = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
    sys.path.insert(0, os.path.join(REF, "DiT"))
    try:
        mg._install_dit_stubs()
        import models as ref_models
        torch.manual_seed(0)
        ref = ref_models.DiT(input_size=8, patch_size=2, hidden_size=32, depth=2, num_heads=4, num_classes=10)
        with torch.no_grad():
            for p in ref.parameters():
                if p.requires_grad and not p.any():
                    p.normal_(std=0.05)                  # adaLN / final layers are zero-initialised
        ours = DiTXL2Harness(input_size=8, patch=2, width=32, depth=2, heads=4, classes=10)
        assert _sig(ours) == _sig(ref)
        ours.load_state_dict(ref.state_dict(), strict=True)
        ref.eval(), ours.eval()
        g = torch.Generator().manual_seed(1)
        x, t, y = torch.randn(3, 4, 8, 8, generator=g), torch.tensor([0, 500, 999]), torch.tensor([1, 7, 10])
        with torch.no_grad():
            a, b = ref(x, t, y), ours(x, t, y)
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5), (a - b).abs().max()
    finally:
        sys.path.remove(os.path.join(REF, "DiT"))
        for k in set(sys.modules) - before:
            del sys.modules[k]
        sys.modules.update(saved_models)
