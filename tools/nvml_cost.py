"""How much does NVML sampling perturb a launch-heavy step?  (developer tool)"""
import os, sys, time, threading
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flocoder_b200 import sampling
from flocoder_b200.unet import Unet
import pynvml as nv

torch.manual_seed(1234)
m = Unet(dim=16, channels=4, dim_mults=[1, 2, 4, 8], n_classes=102, compute_dtype="bf16").cuda().eval()
B = 256
x0 = torch.randn(B, 4, 16, 16, device="cuda")
def step():
    sampling.generate_latents_rk4(m, (B, 4, 16, 16), n_steps=50, source=x0)
for _ in range(3): step()
torch.cuda.synchronize()
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
def timed(n=5):
    torch.cuda.synchronize(); t=time.time()
    for _ in range(n): step()
    torch.cuda.synchronize(); return (time.time()-t)/n*1e3
print("no sampler: %.1f ms/step" % timed())
calls = {"clock": lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
         "reasons": lambda: nv.nvmlDeviceGetCurrentClocksThrottleReasons(h),
         "power": lambda: nv.nvmlDeviceGetPowerUsage(h)}
for name, fn in calls.items():
    t=time.time(); 
    for _ in range(5): fn()
    print(f"{name}: {(time.time()-t)/5*1e3:.2f} ms per call (idle GPU)")
for name, fn in calls.items():
    for period in (0.25, 1.0):
        stop=[False]; cnt=[0]; tt=[0.0]
        def run():
            while not stop[0]:
                t=time.time(); fn(); tt[0]+=time.time()-t; cnt[0]+=1; time.sleep(period)
        th=threading.Thread(target=run, daemon=True); th.start()
        ms=timed(); stop[0]=True; th.join()
        print(f"sampler {name} every {period}s: {ms:.1f} ms/step  ({cnt[0]} calls, {tt[0]/max(cnt[0],1)*1e3:.1f} ms each)")
