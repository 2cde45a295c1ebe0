#!/usr/bin/env python
"""Time single tensor-core GEMM launches (CUDA events, L2-cold operands by rotating buffers).
    python tools/gemm_probe.py [--m 87680] [--cases qkv,proj,fc1,fc2] [--reps 5]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kutil as K

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=87680); ap.add_argument("--d", type=int, default=1024)
ap.add_argument("--cases", default="qkv,proj,fc1,fc2,plain1024,plain4096"); ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--precision", default="bf16"); ap.add_argument("--out", default="")
a = ap.parse_args()
dt = K.TORCH_DT[a.precision]; M, D = a.m, a.d
dev = "cuda"
def mk(r, c): return (torch.randn(r, c, device=dev) * 0.5).to(dt)
lines = []
for case in a.cases.split(","):
    if case == "qkv":   n, k, kw = 3 * D, D, dict(bias=True, out=True)
    elif case == "proj": n, k, kw = D, D, dict(bias=True, gamma=True, x=True)
    elif case == "fc1":  n, k, kw = 4 * D, D, dict(bias=True, act=1, out=True)
    elif case == "fc2":  n, k, kw = D, 4 * D, dict(bias=True, gamma=True, x=True)
    elif case == "plain1024": n, k, kw = 4 * D, D, dict(out=True)
    elif case == "plain4096": n, k, kw = D, 4 * D, dict(out=True)
    else: raise SystemExit(f"unknown case {case}")
    A, Bm = mk(M, k), mk(n, k)
    bias = torch.randn(n, device=dev) if kw.get("bias") else None
    gamma = torch.rand(n, device=dev) if kw.get("gamma") else None
    x = torch.randn(M, n, device=dev) if kw.get("x") else None
    out = torch.empty(M, n, dtype=dt, device=dev) if kw.get("out") else None
    ep = K.epilogue(bias=bias, gamma=gamma, act=kw.get("act", 0), x=x, accumulate_x=bool(kw.get("x")), out=out, ld_out=n)
    for _ in range(2): K.gemm(a.precision, A, Bm, ep)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); K.gemm(a.precision, A, Bm, ep); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    lines.append(f"{case:10s} M={M} N={n} K={k}  {ms:8.3f} ms  {2.0 * M * n * k / ms / 1e9:8.1f} TFLOP/s")
    print(lines[-1], flush=True)
if a.out: open(a.out, "w").write("\n".join(lines) + "\n")
