"""Round-2 GPU tests (through the C ABI): oracle parity at BASELINE batch sizes, stricter bf16 bounds, stream ordering of
plan creation, class-id validation, checkpoint ingest on the GPU, FLOP accounting, and the NCCL path of the sharded sampler."""
import os
import socket
import sys

import pytest
import torch

import oracle
from oracle.unet_oracle import FP32, FUSED_BF16, OracleModel, UnetSpec, unet_forward
from conftest import CONFIGS, rel_l2, seeded_state_dict

pytestmark = pytest.mark.gpu

SHAPE = (8, 4, 16, 16)


def spec_for(n_classes):
    return UnetSpec(dim=16, dim_mults=(1, 2, 4, 8), channels=4, groups=4, n_classes=n_classes)


_models = {}


def gpu_model(n_classes, compute_dtype):
    key = (n_classes, compute_dtype)
    if key not in _models:
        from flocoder_b200.unet import Unet
        torch.manual_seed(1234)
        _models[key] = Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=n_classes, compute_dtype=compute_dtype).cuda().eval()
    return _models[key]


# ---------------------------------------------------------------------------------------------------------------------
# parity against the ORACLE (not the sibling fp32 CUDA path) at the batch sizes that select the other kernel variants
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B", [256, 300, 1024])
@pytest.mark.parametrize("compute_dtype,tol", [("fp32", 1e-5), ("fp16", 2e-3), ("bf16", 8e-3)])
def test_oracle_parity_at_baseline_batch_sizes(compute_dtype, tol, B):
    """B = 256 (clusters of four on the 2x2 stages), 300 (clusters of two), 1024 (no split; several CTAs per SM): the CUDA
    forward of the whole batch, checked on a 16-sample slice spread over the batch (first / middle / last CTAs and cluster
    groups) against the CPU fp32 oracle run on exactly those samples -- samples never see each other.
    fp32: <= 1e-5; fp16: the per-step bar 2e-3; bf16: the format floor (~4.8e-3 per forward) with headroom."""
    n_classes = 102
    m = gpu_model(n_classes, compute_dtype)
    _, sd = seeded_state_dict(n_classes)
    gen = torch.Generator().manual_seed(900 + B)
    x = torch.randn(B, 4, 16, 16, generator=gen)
    t = torch.rand(B, generator=gen) * 999
    cls = torch.randint(0, n_classes, (B,), generator=gen)
    idx = torch.tensor(sorted(set([0, 1, 2, 7, 8, 95, 96, 97, B // 2, B // 2 + 1, B - 9, B - 8, B - 3, B - 2, B - 1, B // 3])))
    for cond_gpu, cond_cpu in ((None, None), ({"class_cond": cls.cuda()}, {"class_cond": cls[idx]})):
        v = m(x.cuda(), t.cuda(), cond_gpu).float().cpu()
        with torch.no_grad():
            v_ref = unet_forward(sd, spec_for(n_classes), x[idx], t[idx], cond_cpu, FP32)
        assert torch.isfinite(v).all()
        per_sample = (v[idx] - v_ref).flatten(1).norm(dim=1) / v_ref.flatten(1).norm(dim=1)
        print(compute_dtype, B, "cond" if cond_gpu else "uncond", "worst sample", float(per_sample.max()), "slice", rel_l2(v[idx], v_ref))
        assert rel_l2(v[idx], v_ref) <= tol
        assert float(per_sample.max()) <= 1.6 * tol


@pytest.mark.parametrize("name", list(CONFIGS))
def test_bf16_every_stage_of_rk4_50_is_at_the_format_floor(goldens, name):
    """bf16, teacher-forced on EVERY stage input of an fp32-oracle RK4 n_steps=50 trajectory (196 evaluations), all three
    BASELINE U-Nets: the CUDA per-step velocity error vs fp32 must stay within 15 % of the error the bf16 FORMAT itself
    causes (the oracle that rounds at the same points) -- i.e. the kernels add nothing measurable on top of
    8-bit mantissas.  The stated 2e-3 bar is below that floor; fp16 (next test) meets it."""
    g = goldens[name]
    n = g["n_classes"]
    m = gpu_model(n, "bf16")
    _, sd = seeded_state_dict(n)
    trace = []
    oracle.generate_latents_rk4(OracleModel(sd, spec_for(n), FP32), SHAPE, n_steps=50, source=g["x0"].clone(), trace=trace)
    assert len(trace) == 196
    emul = OracleModel(sd, spec_for(n), FUSED_BF16)
    worst_ratio, worst_cuda, worst_fmt = 0.0, 0.0, 0.0
    for k, (x_stage, t_stage, v_ref) in enumerate(trace):
        t_vec = torch.full((SHAPE[0],), float(t_stage)) * 999
        e_cuda = rel_l2(m(x_stage.cuda(), t_vec.cuda()), v_ref)
        worst_cuda = max(worst_cuda, e_cuda)
        if k % 4 == 0 or e_cuda > 6e-3:            # the emulating oracle costs a CPU forward: every interval's first stage + outliers
            e_fmt = rel_l2(emul(x_stage, t_vec), v_ref)
            worst_fmt = max(worst_fmt, e_fmt)
            worst_ratio = max(worst_ratio, e_cuda / e_fmt)
            assert e_cuda <= 1.15 * e_fmt + 1e-4, (k, e_cuda, e_fmt)
    print(name, "bf16 worst per-step", worst_cuda, "format floor", worst_fmt, "worst ratio", worst_ratio)
    assert worst_cuda <= 8e-3


@pytest.mark.parametrize("name", list(CONFIGS))
def test_fp16_every_stage_of_rk4_50_meets_the_step_bar(goldens, name):
    """fp16 operands (same kernels, same speed): per-step velocity rel-L2 <= 2e-3 on every one of the 196 stage inputs."""
    g = goldens[name]
    n = g["n_classes"]
    m = gpu_model(n, "fp16")
    _, sd = seeded_state_dict(n)
    trace = []
    oracle.generate_latents_rk4(OracleModel(sd, spec_for(n), FP32), SHAPE, n_steps=50, source=g["x0"].clone(), trace=trace)
    worst = 0.0
    for x_stage, t_stage, v_ref in trace:
        t_vec = torch.full((SHAPE[0],), float(t_stage)) * 999
        worst = max(worst, rel_l2(m(x_stage.cuda(), t_vec.cuda()), v_ref))
    print(name, "fp16 worst per-step velocity rel-L2 over 196 stages:", worst)
    assert worst <= 2e-3


# ---------------------------------------------------------------------------------------------------------------------
# host-side contracts
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("compute_dtype", ["fp32", "bf16"])
def test_first_call_on_a_side_stream(goldens, compute_dtype):
    """Plan creation (workspace memset, tensor maps, graph capture) happens inside the first call for a batch size; on a
    non-blocking side stream its memset must be ordered before the kernels of that same call."""
    from flocoder_b200 import sampling
    from flocoder_b200.unet import Unet
    g = goldens["midi_vqgan"]
    torch.manual_seed(1234)
    m = Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=0, compute_dtype=compute_dtype).cuda().eval()
    side = torch.cuda.Stream()
    x0 = g["x0"].cuda()
    t = g["fwd_t"].cuda()
    torch.cuda.synchronize()
    tol = 1e-5 if compute_dtype == "fp32" else 8e-3
    with torch.cuda.stream(side):
        v = m(x0, t)                                  # first forward of this handle at B=8: creates the plan
        x1, _ = sampling.generate_latents_rk4(m, SHAPE, n_steps=10, source=x0)
        v7 = m(x0[:7], t[:7])                         # another new plan on the side stream
    side.synchronize()
    assert rel_l2(v, g["fwd_v"]) <= tol
    assert rel_l2(v7, g["fwd_v"][:7]) <= tol
    assert rel_l2(x1, g["rk4_10"]) <= (1e-5 if compute_dtype == "fp32" else 1e-2)
    # switching back to the default stream on the same handle is safe (the handle waits for the side stream)
    v2 = m(x0, t)
    torch.cuda.synchronize()
    assert torch.equal(v2, v)


def test_out_of_range_class_ids_raise_like_nn_embedding():
    from flocoder_b200 import sampling
    m = gpu_model(10, "fp32")
    x = torch.randn(4, 4, 16, 16).cuda()
    t = torch.full((4,), 100.0).cuda()
    m(x, t, {"class_cond": torch.tensor([0, 9, 3, 4]).cuda()})
    for bad in ([0, 10, 1, 2], [-1, 0, 1, 2]):
        with pytest.raises(IndexError):
            m(x, t, {"class_cond": torch.tensor(bad).cuda()})
        with pytest.raises(IndexError):
            sampling.generate_latents_rk4(m, (4, 4, 16, 16), n_steps=3, cond={"class_cond": torch.tensor(bad).cuda()}, source=x)


def test_plan_cache_is_bounded():
    """More than 8 distinct batch sizes on one handle: the least recently used plans are released, results stay right."""
    m = gpu_model(0, "fp16")
    eng = m.engine(16, 16)
    x = torch.randn(12, 4, 16, 16).cuda()
    t = torch.full((12,), 321.0).cuda()
    ref = m(x, t)
    free0 = torch.cuda.mem_get_info()[0]
    for b in (1, 2, 3, 4, 5, 6, 7, 9, 10, 11):
        assert torch.allclose(m(x[:b], t[:b]), ref[:b], atol=2e-2)
    assert torch.allclose(m(x, t), ref, atol=0)               # B=12 was evicted and rebuilt: bit-identical
    torch.cuda.synchronize()
    assert free0 - torch.cuda.mem_get_info()[0] < (64 << 20)
    assert eng.launch_count() > 0


def test_conv_flop_accounting_matches_the_survey():
    """flo_unet_op_info: the algorithmic conv FLOPs of the launches of one forward sum to SURVEY 8d's 66,846,720 per sample
    (the N-split stages used to count only the per-CTA slice)."""
    for cd in ("bf16", "fp32"):
        m = gpu_model(102, cd)
        info = m.engine(16, 16).op_info()
        total = sum(fl for (_, kind, fl, _) in info if kind in (0, 1, 5, 6, 7))
        assert total == 66_846_720, (cd, total)


def test_checkpoint_ingest_runs_on_the_gpu(goldens, tmp_path):
    """N2: a train_flow.py-style checkpoint ({'model_state_dict': ...}, general.py:120-137) -> unet_from_checkpoint ->
    the CUDA forward and a trajectory equal the frozen reference outputs (generate_samples.py:76-104 flow)."""
    from flocoder_b200 import sampling
    from flocoder_b200.unet import unet_from_checkpoint
    g = goldens["flowers_sd"]
    _, sd = seeded_state_dict(102)
    path = tmp_path / "flow_99.pt"
    torch.save({"epoch": 99, "model_state_dict": sd, "optimizer_state_dict": {}, "extra.unused": torch.zeros(1)}, path)
    m = unet_from_checkpoint(str(path), device="cuda")
    assert m.load_report["config"] == {"dim": 16, "channels": 4, "dim_mults": [1, 2, 4, 8], "n_classes": 102}
    assert not m.load_report["missing"]
    assert rel_l2(m(g["x0"].cuda(), g["fwd_t"].cuda()), g["fwd_v"]) <= 1e-5
    v = m(g["x0"].cuda(), g["fwd_t"].cuda(), cond={"class_cond": g["cls"].cuda()})
    assert rel_l2(v, g["fwd_v_cls"]) <= 1e-5
    x1, _ = sampling.generate_latents_rk4(m, SHAPE, n_steps=10, source=g["x0"].cuda())
    assert rel_l2(x1, g["rk4_10"]) <= 1e-5
    m16 = unet_from_checkpoint(str(path), device="cuda", compute_dtype="fp16")
    assert rel_l2(m16(g["x0"].cuda(), g["fwd_t"].cuda()), g["fwd_v"]) <= 2e-3


# ---------------------------------------------------------------------------------------------------------------------
# the sharded sampler over NCCL (two GPUs; skipped on a one-GPU box)
# ---------------------------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, batch, out):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch.distributed as dist
    from flocoder_b200 import sampling
    from flocoder_b200.dist import generate_latents_sharded
    from flocoder_b200.unet import Unet
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(1234)
    m = Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=10, compute_dtype="fp32").to(dev).eval()
    x0 = torch.randn(batch, 4, 16, 16, generator=torch.Generator().manual_seed(5678))
    cls = torch.arange(batch) % 10
    init = torch.randn(batch, 4, 16, 16, generator=torch.Generator().manual_seed(99))
    res = {}
    for tag, kw in (("plain", {}), ("init", {"init_latents": init.to(dev), "init_strength": 0.3})):
        full, nfe = generate_latents_sharded(m, (batch, 4, 16, 16), n_steps=6, cond={"class_cond": cls.to(dev)}, cfg_strength=2.0,
                                             source=x0.to(dev), **kw)
        local, _ = generate_latents_sharded(m, (batch, 4, 16, 16), n_steps=6, cond={"class_cond": cls.to(dev)}, cfg_strength=2.0,
                                            source=x0.to(dev), gather=False, **kw)
        res[tag] = (full.cpu(), local.cpu(), nfe)
    if rank == 0:
        ref = {}
        for tag, kw in (("plain", {}), ("init", {"init_latents": init.to(dev), "init_strength": 0.3})):
            ref[tag] = sampling.generate_latents_rk4(m, (batch, 4, 16, 16), n_steps=6, cond={"class_cond": cls.to(dev)},
                                                     cfg_strength=2.0, source=x0.to(dev), **kw)[0].cpu()
        torch.save({"res": res, "ref": ref}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("batch", [9, 16])
def test_sharded_sampler_over_nccl_uneven_shards(tmp_path, batch):
    """generate_latents_sharded on two GPUs over NCCL: uneven (9 = 5 + 4, padded gather) and even shards, class-conditional CFG,
    and the init_latents branch whose per-sample tensor must be sliced with the noise: equals the unsharded run (fp32: to
    summation order)."""
    import torch.multiprocessing as mp
    from flocoder_b200.dist import shard_bounds
    out = str(tmp_path / "nccl.pt")
    mp.spawn(_nccl_worker, args=(2, _free_port(), batch, out), nprocs=2, join=True)
    got = torch.load(out)
    for tag in ("plain", "init"):
        full, local, nfe = got["res"][tag]
        ref = got["ref"][tag]
        assert full.shape == ref.shape
        assert rel_l2(full, ref) <= 2e-6, tag
        lo, hi = shard_bounds(batch, 2, 0)
        assert torch.equal(local, full[lo:hi])


def test_generate_samples_flow_checkpoint_to_decoded_batch(tmp_path):
    """SURVEY.md 8(f) N2, generate_samples.py:61-118,141-159: a {'model_state_dict': ...} checkpoint on disk ->
    load_models_once (U-Net arguments from the tensors + the config's flow section) -> generate_batch -> sampler ->
    latents and decoded images; the latents match the CPU oracle on the same noise."""
    import flocoder_b200.generate as gen
    from flocoder_b200.unet import Unet
    torch.manual_seed(4321)
    src = Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=0)
    sd = {k: v.detach().clone() for k, v in src.state_dict().items()}
    path = tmp_path / "flow_test.pt"
    torch.save({"model_state_dict": sd, "epoch": 1}, path)

    class Codec(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.full((1,), 0.5))

        def decode(self, z):
            return torch.nn.functional.interpolate(z[:, :3] * self.w, scale_factor=8, mode="nearest")

    config = {"image_size": 128, "codec": {"choice": "sd"}, "flow": {"dim_mults": [1, 2, 4, 8], "unet": {"n_classes": 0}}}
    codec, vmodel = gen.load_models_once(str(path), config, device="cuda", codec=Codec(), compute_dtype="fp32")
    assert gen.load_models(str(path), config, "cuda")[1] is vmodel          # cached: same path and config
    latent_shape = gen.infer_latent_shape(config)
    assert latent_shape == (4, 16, 16)
    torch.manual_seed(7)
    decoded, latents, nfe = gen.generate_batch(vmodel, codec, latent_shape, "rk4", 8, 3.0, "cuda", 6)
    torch.manual_seed(7)
    noise = torch.randn((6,) + latent_shape, device="cuda")
    want, _ = oracle.generate_latents_rk4(OracleModel(sd, spec_for(0)), (6,) + latent_shape, n_steps=8, source=noise.cpu())
    assert nfe == 32 and decoded.shape == (6, 3, 128, 128)
    assert rel_l2(latents, want) <= 1e-5
    assert torch.allclose(decoded.cpu(), torch.nn.functional.interpolate(latents.cpu()[:, :3] * 0.5, scale_factor=8, mode="nearest"))


# ---------------------------------------------------------------------------------------------------------------------
# classifier-free guidance as one 2B-sample forward per evaluation (SF_CFG_2B) vs the two-pass form and the reference
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("compute_dtype", ["bf16", "fp16"])
def test_single_pass_cfg_matches_two_pass_and_reference(goldens, compute_dtype, monkeypatch):
    from flocoder_b200 import sampling
    g = goldens["flowers_sd"]
    m = gpu_model(102, compute_dtype)
    cond = {"class_cond": g["cls"].cuda()}
    x0 = g["x0"].cuda()
    one, _ = sampling.generate_latents_rk4(m, SHAPE, n_steps=10, cond=cond, cfg_strength=3.0, source=x0)
    assert rel_l2(one, g["rk4_10_cfg3"]) <= 1e-2                       # north-star final-latent bar for 16-bit operands
    monkeypatch.setenv("FLO_CFG_TWO_PASS", "1")
    two, _ = sampling.generate_latents_rk4(m, SHAPE, n_steps=10, cond=cond, cfg_strength=3.0, source=x0)
    monkeypatch.delenv("FLO_CFG_TWO_PASS")
    assert rel_l2(two, g["rk4_10_cfg3"]) <= 1e-2
    # the two forms run the same kernels on the same rows (a sample's arithmetic does not depend on its batch index)
    assert rel_l2(one, two) <= (3e-3 if compute_dtype == "bf16" else 5e-4)
    # repeated calls re-use the captured graphs and the re-armed arrival counters
    again, _ = sampling.generate_latents_rk4(m, SHAPE, n_steps=10, cond=cond, cfg_strength=3.0, source=x0)
    assert torch.equal(again, one)
    # class conditioning without guidance: B rows through the same in-graph time embedding
    x1, _ = sampling.generate_latents_rk4(m, SHAPE, n_steps=10, cond=cond, cfg_strength=0, source=x0)
    assert rel_l2(x1, g["rk4_10_cls_nocfg"]) <= 1e-2
    monkeypatch.setenv("FLO_CFG_TWO_PASS", "1")
    x2, _ = sampling.generate_latents_rk4(m, SHAPE, n_steps=10, cond=cond, cfg_strength=0, source=x0)
    monkeypatch.delenv("FLO_CFG_TWO_PASS")
    assert torch.equal(x1, x2)                   # same kernels, same rows, same FiLM table: bit-identical


@pytest.mark.parametrize("B", [5, 37, 300])
def test_single_pass_cfg_ragged_batches_vs_fp32_path(B):
    """Odd and multi-wave batches (2B = 10 / 74 / 600 rows per forward; halves of a sample land in different CTAs and waves),
    Euler-grid and RK4 stages, against the fp32 CUDA path (itself pinned to the reference at <= 1e-5)."""
    from flocoder_b200 import sampling
    m16, m32 = gpu_model(10, "fp16"), gpu_model(10, "fp32")
    gen = torch.Generator().manual_seed(B)
    x = torch.randn(B, 4, 16, 16, generator=gen).cuda()
    cls = (torch.arange(B) % 10).cuda()
    a, _ = sampling.generate_latents_rk4(m16, (B, 4, 16, 16), n_steps=5, cond={"class_cond": cls}, cfg_strength=2.5, source=x)
    b, _ = sampling.generate_latents_rk4(m32, (B, 4, 16, 16), n_steps=5, cond={"class_cond": cls}, cfg_strength=2.5, source=x)
    assert rel_l2(a, b) <= 2e-3
    # per-sample independence: every sample of the batch equals the same sample integrated alone
    solo, _ = sampling.generate_latents_rk4(m16, (1, 4, 16, 16), n_steps=5, cond={"class_cond": cls[B // 2: B // 2 + 1]}, cfg_strength=2.5,
                                            source=x[B // 2: B // 2 + 1])
    assert rel_l2(a[B // 2: B // 2 + 1], solo) <= 1e-3


@pytest.mark.parametrize("compute_dtype,tol", [("fp32", 1e-5), ("fp16", 2e-3)])
def test_three_channel_latents_midi_vqgan_3d_gray(compute_dtype, tol):
    """configs/midi_vqgan_3d_gray.yaml: vq_embedding_dim = 3 -> latents (3, 16, 16), Unet(dim=16, channels=3): the init / final 1x1
    convolutions with a channel count that is not 4, against the CPU oracle."""
    from flocoder_b200.unet import Unet
    torch.manual_seed(77)
    m = Unet(dim=16, channels=3, dim_mults=[1, 2, 4, 8], n_classes=0, compute_dtype=compute_dtype)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.cuda().eval()
    spec = UnetSpec(dim=16, dim_mults=(1, 2, 4, 8), channels=3, groups=4, n_classes=0)
    gen = torch.Generator().manual_seed(78)
    x = torch.randn(6, 3, 16, 16, generator=gen)
    t = torch.rand(6, generator=gen) * 999
    with torch.no_grad():
        want = unet_forward(sd, spec, x, t)
    assert rel_l2(m(x.cuda(), t.cuda()), want) <= tol
    from flocoder_b200 import sampling
    x1, _ = sampling.generate_latents_rk4(m, (6, 3, 16, 16), n_steps=5, source=x.cuda())
    want1, _ = oracle.generate_latents_rk4(OracleModel(sd, spec), (6, 3, 16, 16), n_steps=5, source=x)
    assert rel_l2(x1, want1) <= tol


# ---------------------------------------------------------------------------------------------------------------------
# the small-batch plan (quarter-filled k_attn_small tiles with four threads per (row, head); st.async hand-off of the N-split
# stages; transposed K / V projections in k_attn) against the full-tile / fence-and-arrive / row forms of the same kernels:
# the arithmetic is the same in the same order
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B", [37, 256])
def test_plan_switches_are_bit_identical(B, monkeypatch):
    from flocoder_b200.unet import Unet

    def forward_with(env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        torch.manual_seed(1234)
        m = Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=10, compute_dtype="bf16").cuda().eval()   # a fresh engine reads the switches
        x = torch.randn(B, 4, 16, 16, generator=torch.Generator().manual_seed(7)).cuda()
        t = torch.linspace(10.0, 900.0, B).cuda()
        cls = (torch.arange(B) % 10).cuda()
        with torch.no_grad():
            v = m(x, t, {"class_cond": cls}).clone()
        m.invalidate()
        for k in env:
            monkeypatch.delenv(k)
        return v

    base = forward_with({})
    assert torch.isfinite(base).all()
    assert torch.equal(base, forward_with({"FLO_SMALL_DIV": "1"}))        # full k_attn_small tiles, one thread per (row, head)
    assert torch.equal(base, forward_with({"FLO_SMALL_DIV": "2"}))        # half tiles
    assert torch.equal(base, forward_with({"FLO_TX_HANDOFF": "0"}))       # stores -> fence -> barrier -> release-arrive hand-off
    assert torch.equal(base, forward_with({"FLO_ATTN_NO_KTRANS": "1"}))   # K / V projected with pixels as lanes (shuffle softmax)
