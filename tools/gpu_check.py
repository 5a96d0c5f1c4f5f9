"""Developer diagnostics for a GPU box (not part of the product or the tests).

    python tools/gpu_check.py selftest | fp32 | bf16 | time [B]

Each mode is meant to run in its own process so a faulting kernel cannot hide the other results.
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import CONFIGS, rel_l2, seeded_state_dict  # noqa: E402
from oracle.unet_oracle import (BF16_MATCHED, FP32, FUSED_BF16, FUSED_FP16, OracleModel, UnetSpec,  # noqa: E402
                                unet_forward)
import oracle  # noqa: E402


def spec_for(n):
    return UnetSpec(dim=16, dim_mults=(1, 2, 4, 8), channels=4, groups=4, n_classes=n)


def load_golden(name):
    return torch.load(os.path.join(ROOT, "tests", "golden", f"{name}.pt"), weights_only=False)


LAYERWISE = False


def model(n_classes, cd):
    from flocoder_b200 import _lib
    from flocoder_b200.unet import Unet
    torch.manual_seed(1234)
    m = Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=n_classes, compute_dtype=cd).cuda().eval()
    if LAYERWISE:
        m.engine_flags = _lib.FLO_FLAG_LAYERWISE
    return m


def layer_table(m, n_classes, x, t, prec, flags_extra=0):
    from flocoder_b200 import _lib
    _, sd = seeded_state_dict(n_classes)
    eng = _lib.Engine(dim=m.dim, channels=m.channels, dim_mults=m.dim_mults, groups=m.groups, n_classes=m.n_classes,
                      height=16, width=16, compute_dtype=m._resolved_compute_dtype(), device=x.device,
                      state_dict=m.state_dict(), flags=_lib.FLO_FLAG_NO_BUFFER_REUSE | flags_extra | m.engine_flags)
    v = eng.forward(x.float().contiguous(), t.float().contiguous(), None)
    torch.cuda.synchronize()
    trace = {}
    with torch.no_grad():
        vref = unet_forward(sd, spec_for(n_classes), x.cpu(), t.cpu(), None, prec, trace)
    for name in trace:
        if trace[name].dim() != 4:
            continue
        try:
            a = eng.read_activation(name, x.shape[0])
        except ValueError:
            continue
        print(f"  {name:42s} rel_l2={rel_l2(a, trace[name]):.3e}")
    print(f"  {'OUTPUT v':42s} rel_l2={rel_l2(v, vref):.3e}")
    eng.close()


def run_selftest():
    from flocoder_b200 import _lib
    rc, report = _lib.selftest_umma()
    print(report)
    print("selftest rc =", rc, "| last error:", _lib.lib().flo_last_error())
    return rc


def run_path(cd):
    from flocoder_b200 import sampling
    prec = FP32 if cd == "fp32" else (BF16_MATCHED if LAYERWISE else (FUSED_FP16 if cd == "fp16" else FUSED_BF16))
    g = load_golden("midi_vqgan")
    m = model(0, cd)
    print(f"== per-layer ({cd}) vs {'fp32 oracle' if cd == 'fp32' else 'matched oracle'}")
    layer_table(m, 0, g["x0"].cuda(), g["fwd_t"].cuda(), prec)
    for name, n in CONFIGS.items():
        g = load_golden(name)
        m = model(n, cd)
        v = m(g["x0"].cuda(), g["fwd_t"].cuda())
        print(f"{name}: forward vs fp32 reference {rel_l2(v, g['fwd_v']):.3e}")
        x1, _ = sampling.generate_latents_rk4(m, (8, 4, 16, 16), n_steps=10, source=g["x0"].cuda())
        print(f"{name}: rk4_10 vs reference {rel_l2(x1, g['rk4_10']):.3e}")
        x1, _ = sampling.generate_latents_rk4(m, (8, 4, 16, 16), n_steps=50, source=g["x0"].cuda())
        print(f"{name}: rk4_50 vs reference {rel_l2(x1, g['rk4_50']):.3e}")
        x1, _ = sampling.euler_sampler(m, (8, 4, 16, 16), 10, source=g["x0"].cuda())
        print(f"{name}: euler_10 vs reference {rel_l2(x1, g['euler_10']):.3e}")
        if n:
            x1, _ = sampling.generate_latents_rk4(m, (8, 4, 16, 16), n_steps=10, cond={"class_cond": g["cls"].cuda()},
                                                  cfg_strength=3.0, source=g["x0"].cuda())
            print(f"{name}: rk4_10 cfg3 vs reference {rel_l2(x1, g['rk4_10_cfg3']):.3e}")


def run_time(cd, B):
    from flocoder_b200 import sampling
    m = model(0, cd)
    x0 = torch.randn(B, 4, 16, 16, generator=torch.Generator().manual_seed(5678)).cuda()
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        x1, _ = sampling.generate_latents_rk4(m, (B, 4, 16, 16), n_steps=50, source=x0)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"{cd} RK4-50 B={B}: {dt*1e3:.1f} ms -> {B/dt:.1f} samples/s ({dt/196*1e6:.1f} us/forward)")
    eng = m.engine(16, 16)
    print("ops per forward:", eng.launches_per_forward(B), "workspace MB:", eng.workspace_bytes(B) / 2**20)


if __name__ == "__main__":
    if "layerwise" in sys.argv:
        LAYERWISE = True
        sys.argv.remove("layerwise")
    mode = sys.argv[1]
    if mode == "selftest":
        sys.exit(1 if run_selftest() else 0)
    elif mode in ("fp32", "bf16", "fp16"):
        run_path(mode)
    elif mode == "time":
        run_time(sys.argv[2], int(sys.argv[3]))
