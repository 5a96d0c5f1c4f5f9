"""Quick throughput probe (developer tool): python tools/quick_time.py bf16 256 [n_classes]
CUDA-event timing of whole RK4-50 trajectories (3 warm-ups, 7 timed, L2 flushed between), plus a checksum of the result so that
runs with different FLO_* switches can be compared for bit equality."""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flocoder_b200 import sampling            # noqa: E402
from flocoder_b200.unet import Unet           # noqa: E402

cd, B = sys.argv[1], int(sys.argv[2])
ncls = int(sys.argv[3]) if len(sys.argv) > 3 else 102
torch.manual_seed(1234)
m = Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=ncls, compute_dtype=cd).cuda().eval()
x0 = torch.randn(B, 4, 16, 16, generator=torch.Generator().manual_seed(5678)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    x1, _ = sampling.generate_latents_rk4(m, (B, 4, 16, 16), n_steps=50, source=x0)
torch.cuda.synchronize()
ms = []
for _ in range(7):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    x1, _ = sampling.generate_latents_rk4(m, (B, 4, 16, 16), n_steps=50, source=x0)
    b.record()
    torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
ms.sort()
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("FLO_"))
print(f"{cd} B={B} [{tag}]: best {B / ms[0] * 1e3:.0f} median {B / ms[len(ms) // 2] * 1e3:.0f} samples/s "
      f"({ms[len(ms) // 2] / 196 * 1e3:.1f} us/forward)  sha {hashlib.sha256(x1.cpu().numpy().tobytes()).hexdigest()[:12]} "
      f"norm {float(x1.double().norm()):.6f}", flush=True)
