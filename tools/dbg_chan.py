"""Per-channel / per-sample error of one fused-path tensor vs the fp32 oracle (developer tool).
usage: python tools/dbg_chan.py <tensor-name> [dtype]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flocoder_b200 import _lib
from flocoder_b200.unet import Unet
from oracle.unet_oracle import UnetSpec, unet_forward, FP32

name = sys.argv[1]; cd = sys.argv[2] if len(sys.argv) > 2 else "fp16"
torch.manual_seed(1234)
m = Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=0, compute_dtype=cd).cuda().eval()
sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
spec = UnetSpec(dim=16, dim_mults=(1, 2, 4, 8), channels=4, groups=4, n_classes=0)
B = int(os.environ.get("B", "8"))
x = torch.randn(B, 4, 16, 16, generator=torch.Generator().manual_seed(5678))
t = torch.full((B,), 0.25 * 999)
eng = _lib.Engine(dim=16, channels=4, dim_mults=[1, 2, 4, 8], groups=4, n_classes=0, height=16, width=16, compute_dtype=cd,
                  device=torch.device("cuda"), state_dict=m.state_dict(), flags=_lib.FLO_FLAG_NO_BUFFER_REUSE)
eng.forward(x.cuda(), t.cuda(), None)
torch.cuda.synchronize()
trace = {}
with torch.no_grad():
    unet_forward(sd, spec, x, t, None, FP32, trace)
a = eng.read_activation(name, B).float().cpu(); ref = trace[name]
print("ref[0,0:4,0,0]", ref[0,0:4,0,0].tolist(), "ref[0,16:20,0,0]", ref[0,16:20,0,0].tolist(), "got", a[0,0:4,0,0].tolist(), a[0,16:20,0,0].tolist())
print("shape", tuple(a.shape), "total rel", float((a - ref).norm() / ref.norm()))
print("per channel:", " ".join(f"{float((a[:, c] - ref[:, c]).norm() / (ref[:, c].norm() + 1e-12)):.1e}" for c in range(a.shape[1])))
print("per sample :", " ".join(f"{float((a[b] - ref[b]).norm() / (ref[b].norm() + 1e-12)):.1e}" for b in range(B)))
if a.shape[1] == 32 and os.environ.get("XCORR"):
    # which reference channel does each computed channel resemble?
    A = a.permute(1, 0, 2, 3).reshape(32, -1); R = ref.permute(1, 0, 2, 3).reshape(32, -1)
    A = A - A.mean(1, keepdim=True); R = R - R.mean(1, keepdim=True)
    corr = (A @ R.T) / (A.norm(dim=1, keepdim=True) * R.norm(dim=1)[None, :] + 1e-12)
    print("best matching ref channel per computed channel:", corr.argmax(1).tolist())
    print("corr with own:", [round(float(corr[i, i]), 2) for i in range(32)])
    print("mean a / mean ref per channel:", [round(float(a[:, c].mean()), 3) for c in range(16, 32)], [round(float(ref[:, c].mean()), 3) for c in range(16, 32)])
if name == "downs.1.3" and os.environ.get("SUBSETS"):
    import itertools
    import torch.nn.functional as F
    xin = F.pixel_unshuffle(trace["downs.1.2"], 2)            # [B, 64, 4, 4], ref channel order c*4 + (p1*2+p2)
    W = sd["downs.1.3.1.weight"][:, :, 0, 0]; bvec = sd["downs.1.3.1.bias"]
    tgt = a[:, 16:32]
    best = []
    for mask in range(16):
        for use_b in (0, 1):
            sel = [c * 4 + s for c in range(16) for s in range(4) if (mask >> s) & 1]
            y = torch.einsum("oc,bchw->bohw", W[16:32][:, sel], xin[:, sel]) if sel else torch.zeros_like(tgt)
            if use_b: y = y + bvec[16:32].view(1, -1, 1, 1)
            best.append((float((y - tgt).norm() / tgt.norm()), mask, use_b))
    best.sort()
    print("subset fits (err, quadrant mask, bias):", best[:4])
