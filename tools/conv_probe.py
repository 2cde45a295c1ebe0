#!/usr/bin/env python
"""Time single implicit-GEMM 3x3 convolution launches (CUDA events).
    python tools/conv_probe.py [--cases out1,rn148,rcu148] [--reps 5]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kutil as K

ap = argparse.ArgumentParser()
ap.add_argument("--cases", default="out1,rcu148,rn74"); ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--precision", default="bf16"); ap.add_argument("--batch", type=int, default=64)
a = ap.parse_args()
dt = K.TORCH_DT[a.precision]
CASES = {"out1": (296, 296, 256, 128), "rcu148": (148, 148, 256, 256), "rn74": (74, 74, 512, 256), "rcu296": (296, 296, 256, 256)}
for case in a.cases.split(","):
    H, W, cin, cout = CASES[case]
    B = a.batch
    x = (torch.randn(B, H, W, cin, device="cuda") * 0.5).to(dt)
    w = K.pack_conv3x3(torch.randn(cout, cin, 3, 3, device="cuda") * (9 * cin) ** -0.5, dt)
    bias = torch.randn(cout, device="cuda")
    out = torch.empty(B, H, W, cout, dtype=dt, device="cuda")
    ep = K.epilogue(bias=bias, out=out, ld_out=cout)
    for _ in range(2): K.conv3x3(a.precision, x, w, cout, ep)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); K.conv3x3(a.precision, x, w, cout, ep); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{case:8s} B={B} {H}x{W} {cin}->{cout}  {ms:8.3f} ms  {2.0 * B * H * W * cout * 9 * cin / ms / 1e9:8.1f} TFLOP/s", flush=True)
