"""the inpainting U-Net (``Unet(mask_cond=True)``, unet.py:214-235,298-305,336-340,360-364) on the CUDA
path, through the C ABI (``flo_unet_set_mask`` + the fp32 op program), against outputs of the UNMODIFIED reference frozen by
``oracle/make_golden.py inpaint`` -- for a flowers-sized net (dim 16, 16x16 latents) and for the shape of
``configs/midi_inpainting.yaml`` (latent (4, 8, 8), ``Unet(dim=8)``: two channels per GroupNorm group, a 1x1 bottleneck).
Bar: fp32, rel-L2 <= 1e-5 per velocity and per final latent (BASELINE.json north_star)."""
import os

import pytest
import torch

import oracle
from oracle.unet_oracle import OracleModel, UnetSpec, unet_forward
from conftest import GOLDEN_DIR, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-5
NAMES = ["inpaint_16", "midi_inpainting"]
_cache = {}


def case(name):
    if name not in _cache:
        from flocoder_b200.unet import Unet
        g = torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)
        torch.manual_seed(g["model_seed"])
        m = Unet(dim=g["dim"], channels=4, dim_mults=[1, 2, 4, 8], n_classes=0, mask_cond=True, compute_dtype="fp32")
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        _cache[name] = (g, m.cuda().eval(), sd, UnetSpec(dim=g["dim"], dim_mults=(1, 2, 4, 8), channels=4, groups=4, n_classes=0))
    return _cache[name]


@pytest.mark.parametrize("name", NAMES)
def test_forward_with_mask_matches_reference_golden(name):
    g, m, _, _ = case(name)
    x0, t = g["x0"].cuda(), g["fwd_t"].cuda()
    shape = tuple(x0.shape)
    v = m(x0, t, cond={"mask_cond": g["mask_latents"].cuda()})
    assert rel_l2(v, g["fwd_v_mask"]) <= TOL
    # no mask: every mask branch is skipped (key_usable(cond, 'mask_cond') is false, unet.py:298,336,360)
    assert rel_l2(m(x0, t), g["fwd_v_nomask"]) <= TOL
    assert rel_l2(m(x0, t, cond={"mask_cond": None}), g["fwd_v_nomask"]) <= TOL
    # all-ones mask: only the first fusion is bypassed (torch.allclose, unet.py:301); the per-scale fusions still run
    assert rel_l2(m(x0, t, cond={"mask_cond": torch.ones(shape, device="cuda")}), g["fwd_v_ones"]) <= TOL
    # the bypass test is batch-wide: half the samples all ones does NOT bypass
    assert rel_l2(m(x0, t, cond={"mask_cond": g["mask_half"].cuda()}), g["fwd_v_half"]) <= TOL
    # and back to the real mask: the per-batch-size mask state follows the argument
    assert rel_l2(m(x0, t, cond={"mask_cond": g["mask_latents"].cuda()}), g["fwd_v_mask"]) <= TOL
    assert abs(rel_l2(g["fwd_v_mask"], g["fwd_v_nomask"])) > 1e-2, "the golden must depend on the mask"


@pytest.mark.parametrize("name", NAMES)
def test_integrators_with_mask_match_reference_golden(name):
    from flocoder_b200 import sampling
    g, m, _, _ = case(name)
    x0 = g["x0"].cuda()
    shape = tuple(x0.shape)
    cond = {"mask_cond": g["mask_latents"].cuda()}
    x1, nfe = sampling.generate_latents_rk4(m, shape, n_steps=10, cond=cond, source=x0)
    assert nfe == 40 and rel_l2(x1, g["rk4_10_mask"]) <= TOL
    x1, nfe = sampling.euler_sampler(m, shape, 10, cond=cond, source=x0)
    assert nfe == 10 and rel_l2(x1, g["euler_10_mask"]) <= TOL
    # the reference's generic composition (rk4_step over forward()) through OUR forward gives the same trajectory
    ts = sampling.time_grid(10, device="cuda")
    t_vec = torch.zeros(shape[0], device="cuda")
    y = x0.clone()
    for i in range(len(ts) - 1):
        y = sampling.rk4_step(lambda yy, tt: sampling.v_func_cfg(m, cond, 3.0, t_vec, yy, tt), y, ts[i], ts[i + 1] - ts[i])
    assert rel_l2(y, g["rk4_10_mask"]) <= TOL


def test_mask_pyramid_and_fusion_on_fresh_inputs_vs_oracle():
    """Fresh (unseeded-by-the-goldens) inputs at a batch size that is not a multiple of anything: random soft masks, so the
    bilinear down-sizing to 8x8 / 4x4 / 2x2 (F.interpolate, unet.py:338,362) is exercised with non-trivial weights."""
    g, m, sd, spec = case("inpaint_16")
    gen = torch.Generator().manual_seed(99)
    x = torch.randn(5, 4, 16, 16, generator=gen)
    mask = torch.rand(5, 4, 16, 16, generator=gen)
    t = torch.rand(5, generator=gen) * 999
    with torch.no_grad():
        want = unet_forward(sd, spec, x, t, cond={"mask_cond": mask})
    got = m(x.cuda(), t.cuda(), cond={"mask_cond": mask.cuda()})
    assert rel_l2(got, want) <= TOL


def test_mask_encoder_to_sampler_flow():
    """The flow of train_flow.py:239,340-341 / sampling.py:221: pixel mask -> MaskEncoder -> cond['mask_cond'] -> sampler."""
    from flocoder_b200 import sampling
    from flocoder_b200.inpainting import MaskEncoder, mask_blending
    g, m, sd, spec = case("midi_inpainting")
    torch.manual_seed(g["model_seed"] + 1)
    enc = MaskEncoder().cuda().eval()
    with torch.no_grad():
        lat = enc(g["mask_pixels"].cuda())
    assert torch.allclose(lat.cpu(), g["mask_latents"], atol=2e-6, rtol=1e-5)
    src = mask_blending(torch.zeros_like(g["x0"]).cuda(), lat, noise=g["x0"].cuda())

    class Codec(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(1))

        def decode(self, z):
            return z[:, :3] * self.w

    shape = tuple(g["x0"].shape)
    lat_out, dec, nfe = sampling.sampler(m, Codec().cuda(), batch_size=shape[0], n_steps=6, cond={"mask_cond": lat},
                                         latent_shape=shape[1:], source=src)
    model = OracleModel(sd, spec)
    want, _ = oracle.generate_latents_rk4(model, shape, n_steps=6, cond={"mask_cond": lat.cpu()}, source=src.cpu())
    assert rel_l2(lat_out, want) <= TOL and dec.shape == (shape[0], 3) + shape[2:]


def test_mask_cond_needs_the_fp32_path_and_a_latent_shaped_mask():
    from flocoder_b200.unet import Unet
    m = Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=0, mask_cond=True, compute_dtype="bf16").cuda().eval()
    with pytest.raises(NotImplementedError, match="mask_cond"):
        m(torch.zeros(2, 4, 16, 16, device="cuda"), torch.zeros(2, device="cuda"))
    g, mm, _, _ = case("inpaint_16")
    with pytest.raises(ValueError, match="mask_cond"):
        mm(g["x0"].cuda(), g["fwd_t"].cuda(), cond={"mask_cond": torch.ones(8, 1, 16, 16, device="cuda")})
