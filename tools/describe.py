"""Print the fused-stage plan (host only, no GPU): python tools/describe.py [B] [n_classes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flocoder_b200 import _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = _lib.make_cfg(dim=16, channels=4, dim_mults=[1, 2, 4, 8], groups=4, n_classes=int(sys.argv[2]) if len(sys.argv) > 2 else 0,
                    height=16, width=16, compute_dtype="bf16", device_index=0)
for line in _lib.describe_plan(cfg, B).splitlines():
    if line.startswith(("stage", "fused")) or "--steps" in sys.argv and line.startswith("      step"):
        print(line)
