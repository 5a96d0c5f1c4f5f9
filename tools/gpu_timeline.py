"""clock64 timeline of CTA 0 of every fused chain stage (developer tool): python tools/gpu_timeline.py bf16 256"""
import os
import sys

import torch

os.environ["FLO_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flocoder_b200.unet import Unet  # noqa: E402

cd, B = sys.argv[1], int(sys.argv[2])
torch.manual_seed(1234)
m = Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=0, compute_dtype=cd).cuda().eval()
x = torch.randn(B, 4, 16, 16, device="cuda")
t = torch.full((B,), 300.0, device="cuda")
for _ in range(3):
    m(x, t)
torch.cuda.synchronize()
eng = m.engine(16, 16)
names = eng.op_names()
prev_end = None
print("stage                 entry-after-prev-end  prologue   body(cta0)  last-cta-end-after-cta0-end   [us, globaltimer]")
for i, name in enumerate(names):
    tl = eng.read_timeline(B, i)
    if tl[100] == 0:
        continue
    gap = (tl[100] - prev_end) / 1e3 if prev_end else float("nan")
    print(f"{i:3d} {name:18s} {gap:10.2f} {(tl[101]-tl[100])/1e3:10.2f} {(tl[102]-tl[101])/1e3:10.2f} {(tl[103]-tl[102])/1e3:10.2f}")
    prev_end = tl[103]
if "gaps" in sys.argv:
    sys.exit(0)
for i, name in enumerate(names):
    tl = eng.read_timeline(B, i)
    start, end = tl[64], tl[65]
    if start == 0:
        continue
    print(f"== stage {i} {name}: kernel body {end - start} cycles")
    if not name.startswith("S_") and tl[5]:
        d = lambda k: (tl[k] - start) if tl[k] else -1   # noqa: E731
        print(f"   attn_small: griddep_wait_done +{d(5)} input_loaded +{d(6)} | mma0_start +{d(0)} qkv_issued +{d(1)} | epiA_start +{d(2)} epiA_end +{d(4)} | "
              f"epiB_start +{d(10)} epiB_end +{d(11)} | mma_out_start +{d(24)} issued +{d(25)} | epiC_start +{d(26)} stats_done +{d(27)} "
              f"epiC_end +{d(28)} end +{end - start}")
        continue
    if not name.startswith("S_"):
        d = lambda k: (tl[k] - start) if tl[k] else -1   # noqa: E731
        print(f"   attn: load+mma0_start +{d(0)} kvconv_issued +{d(1)} | epi0_start +{d(2)} maxpass_done +{d(3)} epi0_end +{d(4)} | "
              f"mma1_start +{d(8)} ctx+q_issued +{d(9)} | epi1_start +{d(10)} epi1_end +{d(11)} | mma2_start +{d(16)} | "
              f"epi2_start +{d(18)} epi2_end +{d(19)} | mma3_start +{d(24)} | epi3_start +{d(26)} stats_own_done +{d(27)} stats_barrier +{d(28)} ms_done +{d(30)} written +{d(29)} end +{end - start}")
        continue
    if tl[105]:
        print("   weight chunks 8..15 (cycles rel. kernel start): " + " | ".join(
            f"issue {tl[105+3*k]-start} landed {tl[104+3*k]-start if tl[104+3*k] else -1} seen {tl[106+3*k]-start}" for k in range(8) if tl[105+3*k]))
    for s in range(8):
        row = tl[s * 8: s * 8 + 8]
        if row[2] == 0:
            continue
        print(f"   step {s}: mma_start +{row[0]-start:7d} issue {row[1]-row[0] if row[1] else 0:6d} | epi_start +{row[2]-start:7d} "
              f"(after issue {row[2]-row[1] if row[1] else 0:6d}) pass1 {row[3]-row[2] if row[3] else 0:6d} reduce {row[4]-row[3] if row[4] else 0:6d} "
              f"pass2 {row[5]-(row[4] if row[4] else row[2]):6d}"
              + (f" | fast head: tmem +{row[6]-row[3]} scatter +{row[7]-row[6]} barrier +{tl[80+s]-row[7]} finish +{row[4]-tl[80+s]}" if row[6] and tl[80 + s] else ""))
