#!/usr/bin/env python3
"""How long does each NVML query take, and how much does it delay a busy GPU?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml as nv
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
x = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
def busy(n=40):
    for _ in range(n): x @ x
calls = {
    "clock_sm": lambda: nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
    "reasons": lambda: (getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons)(h),
    "power": lambda: nv.nvmlDeviceGetPowerUsage(h),
    "none": lambda: None,
}
busy(); torch.cuda.synchronize()
for name, fn in calls.items():
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); busy(); b.record()
        time.sleep(0.01)
        t0 = time.perf_counter(); r = fn(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"{name:9s} call {1e3 * (t1 - t0):7.2f} ms  region {a.elapsed_time(b):7.2f} ms  -> {r}")
