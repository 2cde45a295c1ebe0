#!/usr/bin/env python
"""Time every attention kernel / polynomial share in one process (CUDA events, median of --reps).
python tools/attn_sweep.py [--batch 64] [--ntok 1370] [--precisions fp16,bf16] [--variants tc:1,tc:2,...]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kutil as K
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64); ap.add_argument("--ntok", type=int, default=1370)
ap.add_argument("--heads", type=int, default=16); ap.add_argument("--precisions", default="fp16,bf16")
ap.add_argument("--variants", default=",".join(f"tc:{p}" for p in (0, 1, 2, 3, 4)))
ap.add_argument("--reps", type=int, default=7)
a = ap.parse_args()
fl = 4.0 * a.batch * a.heads * a.ntok * a.ntok * 64
for prec in a.precisions.split(","):
    dt = K.TORCH_DT[prec]
    torch.manual_seed(0)
    qkv = (torch.randn(a.batch * a.ntok, 3 * a.heads * 64, device="cuda")).to(dt)
    ref = None
    for v in a.variants.split(","):
        for _ in range(2): out = K.attention(prec, qkv, a.batch, a.ntok, a.heads, v)
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); K.attention(prec, qkv, a.batch, a.ntok, a.heads, v); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        if ref is None: ref = out.float()
        err = float((out.float() - ref).abs().max())
        print(f"attention[{v}] {prec} B={a.batch} N={a.ntok} H={a.heads}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s  max|d| vs first variant {err:.2e}", flush=True)
