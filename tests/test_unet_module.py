"""Host-side checks of the drop-in ``Unet`` module (no GPU): parameter manifest, seeded-init
parity with the reference (via frozen fingerprints), argument validation."""
import hashlib

import pytest
import torch

from conftest import CONFIGS, seeded_state_dict
from flocoder_b200.unet import Unet


def fingerprint(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().contiguous().cpu().numpy().tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("name", list(CONFIGS))
def test_seeded_init_matches_reference_fingerprint(goldens, name):
    g = goldens[name]
    model, sd = seeded_state_dict(g["n_classes"], g["model_seed"])
    assert list(sd.keys()) == g["sd_names"]
    assert sum(p.numel() for p in model.parameters()) == g["n_params"]
    if torch.__version__ != g["torch_version"]:
        pytest.skip("torch version differs from the one the goldens were frozen with")
    assert fingerprint(sd) == g["sd_sha256"]


def test_param_counts_match_survey():
    # 2,619,172 / 2,573,092 / 2,607,396 parameters, 298 / 293 / 298 tensors
    expect = {102: (2619172, 298), 0: (2573092, 293), 10: (2607396, 298)}
    for n_cls, (n_par, n_tensors) in expect.items():
        m = Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=n_cls)
        assert sum(p.numel() for p in m.parameters()) == n_par
        assert len(m.state_dict()) == n_tensors


def test_cond_handling_matches_the_reference():
    with pytest.raises(TypeError):
        Unet.split_cond(torch.arange(4))
    assert Unet.split_cond(None) is None
    assert Unet.split_cond({"class_cond": None}) is None
    # unet.py:298: a mask is consumed only by a module built with mask_cond=True
    m = Unet(dim=16, channels=4, n_classes=0)
    assert m.mask_of({"mask_cond": torch.ones(1, 4, 16, 16)}) is None
    mm = Unet(dim=16, channels=4, n_classes=0, mask_cond=True)
    mask = torch.rand(1, 4, 16, 16)
    assert mm.mask_of({"mask_cond": mask}) is mask and mm.mask_of({"mask_cond": None}) is None and mm.mask_of(None) is None


@pytest.mark.parametrize("name", ["inpaint_16", "midi_inpainting"])
def test_inpainting_unet_and_mask_encoder_seeded_init_match_reference(name):
    """mask_cond=True: the mask-fusion parameters sit where the reference registers them (unet.py:214-235) and draw the
    same RNG stream; MaskEncoder (inpainting.py:180-245) likewise."""
    import os
    from conftest import GOLDEN_DIR
    from flocoder_b200.inpainting import MaskEncoder
    g = torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)
    torch.manual_seed(g["model_seed"])
    m = Unet(dim=g["dim"], channels=4, dim_mults=[1, 2, 4, 8], n_classes=0, mask_cond=True)
    assert list(m.state_dict().keys()) == g["sd_names"]
    assert sum(p.numel() for p in m.parameters()) == g["n_params"]
    torch.manual_seed(g["model_seed"] + 1)
    enc = MaskEncoder().eval()
    assert list(enc.state_dict().keys()) == g["enc_sd_names"]
    if torch.__version__ != g["torch_version"]:
        pytest.skip("torch version differs from the one the goldens were frozen with")
    assert fingerprint(m.state_dict()) == g["sd_sha256"]
    assert fingerprint(enc.state_dict()) == g["enc_sd_sha256"]
    with torch.no_grad():
        lat = enc(g["mask_pixels"])
    assert torch.allclose(lat, g["mask_latents"], atol=1e-6, rtol=1e-6)


def test_no_cpu_fallback():
    m = Unet(dim=16, channels=4, n_classes=0)
    with pytest.raises((RuntimeError, ImportError)):
        m(torch.zeros(1, 4, 16, 16), torch.zeros(1))


def test_load_reference_state_dict_roundtrip():
    a = Unet(dim=16, channels=4, n_classes=10)
    b = Unet(dim=16, channels=4, n_classes=10)
    missing = b.load_state_dict(a.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)


@pytest.mark.parametrize("kw", [dict(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=102), dict(dim=8, channels=4, dim_mults=[1, 2, 4], n_classes=0),
                                dict(dim=32, channels=3, dim_mults=[1, 2], n_classes=10)])
def test_unet_from_checkpoint_infers_the_constructor_arguments(kw, tmp_path):
    """generate_samples.py:78-108: a flow checkpoint ({'model_state_dict': ...}) -> U-Net; here dim_mults and n_classes are
    recovered from the tensors themselves instead of a Hydra config."""
    from flocoder_b200.unet import Unet, infer_unet_config, unet_from_checkpoint
    torch.manual_seed(3)
    src = Unet(**kw)
    assert infer_unet_config(src.state_dict()) == {"dim": kw["dim"], "channels": kw["channels"], "dim_mults": kw["dim_mults"], "n_classes": kw["n_classes"]}
    path = tmp_path / "flow_test.pt"
    torch.save({"model_state_dict": src.state_dict(), "epoch": 3}, path)
    m = unet_from_checkpoint(str(path))
    assert m.load_report["missing"] == [] and m.load_report["unexpected"] == [] and not m.training
    for (ka, va), (kb, vb) in zip(src.state_dict().items(), m.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    # strict=False like the reference: a checkpoint with an extra buffer still loads and is reported
    sd = dict(src.state_dict()); sd["ema.decay"] = torch.tensor(0.999)
    assert unet_from_checkpoint(sd).load_report["unexpected"] == ["ema.decay"]
