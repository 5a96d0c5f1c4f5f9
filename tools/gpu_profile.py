"""Per-op CUDA-event profile of one forward (developer tool): python tools/gpu_profile.py <dtype> <B> [layerwise]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flocoder_b200.unet import Unet  # noqa: E402

KINDS = ["init", "conv", "gn", "linattn", "midattn", "final", "chain", "attn"]


def main():
    cd, B = sys.argv[1], int(sys.argv[2])
    torch.manual_seed(1234)
    m = Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=0, compute_dtype=cd).cuda().eval()
    if "layerwise" in sys.argv:
        from flocoder_b200 import _lib
        m.engine_flags = _lib.FLO_FLAG_LAYERWISE
    eng = m.engine(16, 16)
    info = eng.op_info()
    ms = eng.profile_ops(B, reps=10)
    tot = sum(ms)
    print(f"== {cd} B={B}: sum of kernels {tot*1e3:.1f} us over {len(ms)} ops")
    by_kind = {}
    for (name, kind, fl, by), t in zip(info, ms):
        k = by_kind.setdefault(KINDS[kind], [0, 0.0, 0.0, 0.0])
        k[0] += 1; k[1] += t; k[2] += fl * B; k[3] += by * B
    for k, (n, t, fl, by) in by_kind.items():
        print(f"  {k:8s} n={n:3d} time={t*1e3:8.1f} us ({100*t/tot:5.1f}%)  {fl/(t*1e-3)/1e12 if t else 0:8.2f} TFLOP/s  "
              f"{by/(t*1e-3)/1e9 if t else 0:8.1f} GB/s")
    for i, ((name, kind, fl, by), t) in enumerate(zip(info, ms)):
        print(f"  {i:3d} {KINDS[kind]:8s} {name:32s} {t*1e3:8.2f} us  {fl*B/(t*1e-3)/1e12 if t else 0:7.2f} TF/s  "
              f"{by*B/(t*1e-3)/1e9 if t else 0:8.1f} GB/s")


if __name__ == "__main__":
    main()
