"""Experiment: one batch as K independent sub-batches on K streams (K engines), so that the low-resolution stages of one
sub-batch (8..128 CTAs, latency chains) overlap the full-resolution stages of another.

    python tools/exp_substreams.py [B] [dtype]        # prints samples/s for K = 1, 2, 3, 4
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flocoder_b200 import _lib, sampling          # noqa: E402
from flocoder_b200.unet import Unet               # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    dev = torch.device("cuda", 0)
    x0 = torch.randn(B, 4, 16, 16, generator=torch.Generator().manual_seed(5678)).to(dev)
    ts = sampling.warp_time(torch.linspace(0, 1, 50, dtype=torch.float32)).tolist()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = None
    for K in (1, 2, 3, 4, 2, 1):
        models = []
        for _ in range(K):
            torch.manual_seed(1234)
            models.append(Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=102, compute_dtype=dtype).to(dev).eval())
        engs = [m.engine(16, 16) for m in models]
        streams = [torch.cuda.Stream(dev) for _ in range(K)]
        # sub-batch bounds: multiples of 32 samples (the largest sample group of any stage kernel)
        per = -(-B // K)
        per = -(-per // 32) * 32
        bounds = [(min(i * per, B), min((i + 1) * per, B)) for i in range(K)]
        bounds = [b for b in bounds if b[1] > b[0]]

        def run():
            state = x0.clone()
            cur = torch.cuda.current_stream(dev)
            ev0 = torch.cuda.Event(); ev0.record(cur)
            for (lo, hi), e, s in zip(bounds, engs, streams):
                s.wait_event(ev0)
                with torch.cuda.stream(s):
                    e.integrate(state[lo:hi], ts, _lib.FLO_RK4)
                ev = torch.cuda.Event(); ev.record(s)
                cur.wait_event(ev)
            return state

        for _ in range(3):
            out = run()
        torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        err = float((out - ref).norm() / ref.norm())
        steps = 5
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        t0.record()
        for _ in range(steps):
            flush.zero_()
            run()
        t1.record()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        ms = t0.elapsed_time(t1) / steps
        print(f"B={B} {dtype} K={K} bounds={bounds}: {ms:.2f} ms/step  {B / ms * 1e3:.0f} samples/s  (wall {1e3 * (w1 - w0) / steps:.2f} ms)"
              f"  rel diff vs K=1 {err:.2e}", flush=True)
        for m in models:
            m.invalidate()


if __name__ == "__main__":
    main()
