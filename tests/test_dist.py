"""Multi-rank host logic on CPU: world_size-2 gloo.  The sharded sampler is driven with the CPU oracle
as the (foreign) model, so the slicing / padding / all-gather code is exactly what runs on the GPUs --
only the per-rank integrator differs (NCCL + flo_integrate there)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flocoder_b200.dist import shard_bounds, shard_cond


def test_shard_bounds_cover_the_batch():
    for b in (0, 1, 7, 8, 256, 8192, 8191):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_bounds(b, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            for (lo0, hi0), (lo1, hi1) in zip(spans, spans[1:]):
                assert hi0 == lo1
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)
    assert shard_cond(None, 0, 4) is None
    c = shard_cond({"class_cond": torch.arange(8)}, 2, 5)
    assert c["class_cond"].tolist() == [2, 3, 4]
    assert shard_cond({"class_cond": None}, 0, 2) == {"class_cond": None}
    # the inpainting mask is per sample too (sampling.py:221) and is sliced with the noise
    m = torch.arange(8 * 4 * 2 * 2, dtype=torch.float32).reshape(8, 4, 2, 2)
    c = shard_cond({"class_cond": torch.arange(8), "mask_cond": m}, 5, 8)
    assert torch.equal(c["mask_cond"], m[5:8]) and c["class_cond"].tolist() == [5, 6, 7]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, n_classes, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    from conftest import seeded_state_dict
    from oracle.unet_oracle import OracleModel, UnetSpec
    from flocoder_b200.dist import generate_latents_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, sd = seeded_state_dict(n_classes)
    model = OracleModel(sd, UnetSpec(dim=16, dim_mults=(1, 2, 4, 8), channels=4, groups=4, n_classes=n_classes))
    x0 = torch.randn(batch, 4, 16, 16, generator=torch.Generator().manual_seed(5678))
    cond = {"class_cond": torch.arange(batch) % n_classes} if n_classes else None
    full, nfe = generate_latents_sharded(model, (batch, 4, 16, 16), n_steps=3, cond=cond, cfg_strength=2.0, source=x0)
    local, _ = generate_latents_sharded(model, (batch, 4, 16, 16), n_steps=3, cond=cond, cfg_strength=2.0, source=x0,
                                        gather=False)
    if rank == 0:
        torch.save({"full": full, "nfe": nfe, "local": local}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch,n_classes", [(6, 0), (5, 10)])
def test_sharded_equals_unsharded_gloo(tmp_path, batch, n_classes):
    """Rows (iii) of SURVEY.md section 4: sharded == single-process output, slice for slice; uneven split too."""
    import oracle
    from conftest import seeded_state_dict
    from oracle.unet_oracle import OracleModel, UnetSpec
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), batch, n_classes, out), nprocs=2, join=True)
    got = torch.load(out)
    _, sd = seeded_state_dict(n_classes)
    model = OracleModel(sd, UnetSpec(dim=16, dim_mults=(1, 2, 4, 8), channels=4, groups=4, n_classes=n_classes))
    x0 = torch.randn(batch, 4, 16, 16, generator=torch.Generator().manual_seed(5678))
    cond = {"class_cond": torch.arange(batch) % n_classes} if n_classes else None
    ref, nfe = oracle.generate_latents_rk4(model, (batch, 4, 16, 16), n_steps=3, cond=cond, cfg_strength=2.0,
                                           source=x0.clone())
    assert got["nfe"] == nfe == 12
    assert got["full"].shape == ref.shape
    assert float((got["full"] - ref).abs().max()) <= 2e-6          # per-sample ops: batch size only changes blocking
    lo, hi = shard_bounds(batch, 2, 0)
    assert torch.equal(got["local"], got["full"][lo:hi])
