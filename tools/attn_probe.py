#!/usr/bin/env python
"""Time the attention kernels alone (CUDA events).  python tools/attn_probe.py [--batch 64] [--variant tc|tc:<poly eighths>|mma]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kutil as K
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64); ap.add_argument("--ntok", type=int, default=1370)
ap.add_argument("--heads", type=int, default=16); ap.add_argument("--precision", default="bf16")
ap.add_argument("--variant", default="tc"); ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dt = K.TORCH_DT[a.precision]
qkv = (torch.randn(a.batch * a.ntok, 3 * a.heads * 64, device="cuda")).to(dt)
for _ in range(2): K.attention(a.precision, qkv, a.batch, a.ntok, a.heads, a.variant)
torch.cuda.synchronize()
ts = []
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); K.attention(a.precision, qkv, a.batch, a.ntok, a.heads, a.variant); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
fl = 4.0 * a.batch * a.heads * a.ntok * a.ntok * 64
print(f"attention[{a.variant}] B={a.batch} N={a.ntok} H={a.heads}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")
