"""Two RK4 intervals (8 forwards) at a given batch -- the short command used under ncu.
usage: python tools/one_forward.py <dtype> <B> [nograph]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flocoder_b200 import _lib, sampling  # noqa: E402
from flocoder_b200.unet import Unet  # noqa: E402

cd, B = sys.argv[1], int(sys.argv[2])
torch.manual_seed(1234)
m = Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=102, compute_dtype=cd).cuda().eval()
x0 = torch.randn(B, 4, 16, 16, generator=torch.Generator().manual_seed(5678)).cuda()
x1, _ = sampling.generate_latents_rk4(m, (B, 4, 16, 16), n_steps=3, source=x0)
torch.cuda.synchronize()
print("ok", float(x1.norm()))
