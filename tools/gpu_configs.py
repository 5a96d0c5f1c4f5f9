"""Other U-Net shapes through the fused path vs the fp32 CUDA path: either they match or they fail with a clean error."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flocoder_b200.unet import Unet

cases = [(16, [1, 2, 4, 8], 16, 0), (16, [1, 2], 16, 0), (16, [1, 2, 4], 16, 10), (32, [1, 2, 4], 16, 0), (32, [1, 2, 4, 8], 16, 0),
         (16, [1, 2, 4, 8], 32, 0), (16, [1, 2, 4], 8, 0), (64, [1, 2], 16, 0), (16, [1, 1, 2, 2], 16, 0)]
for dim, mults, hw, ncls in cases:
    tag = f"dim={dim} mults={mults} {hw}x{hw} n_classes={ncls}"
    try:
        torch.manual_seed(7)
        m32 = Unet(dim=dim, channels=4, dim_mults=mults, n_classes=ncls, compute_dtype="fp32").cuda().eval()
        m16 = Unet(dim=dim, channels=4, dim_mults=mults, n_classes=ncls, compute_dtype="fp16").cuda().eval()
        m16.load_state_dict(m32.state_dict())
        B = 5
        x = torch.randn(B, 4, hw, hw).cuda(); t = torch.rand(B).cuda() * 999
        v32 = m32(x, t); v16 = m16(x, t)
        e = float((v16.double() - v32.double()).norm() / v32.double().norm())
        print(f"{tag}: fused fp16 vs fp32 path rel-L2 {e:.3e}  {'OK' if e < 3e-3 else 'MISMATCH'}")
    except Exception as ex:          # noqa: BLE001
        print(f"{tag}: {type(ex).__name__}: {str(ex)[:160]}")
