#!/usr/bin/env python
"""Where a softmax warp's time goes inside the default attention kernel: clock64 stamps of every phase
(mde_k_attention_trace), averaged over the traced CTAs.  python tools/attn_trace.py [--batch 64] [--ntok 1370]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import kutil as K
from monocular_depth_estimation_trt_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64); ap.add_argument("--ntok", type=int, default=1370)
ap.add_argument("--heads", type=int, default=16); ap.add_argument("--precision", default="bf16")
a = ap.parse_args()
lib = _lib.load()
dt = K.TORCH_DT[a.precision]
qkv = torch.randn(a.batch * a.ntok, 3 * a.heads * 64, device="cuda").to(dt)
out = torch.empty(a.batch * a.ntok, a.heads * 64, dtype=dt, device="cuda")
trace = torch.zeros(2048, 4, 64, dtype=torch.int64, device="cuda")
for _ in range(2):
    trace.zero_()
    _lib.check(lib.mde_k_attention_trace(_lib.PRECISIONS[a.precision], K.ptr(qkv), K.ptr(out), a.batch, a.ntok, a.heads, K.ptr(trace), K.stream()), "trace")
torch.cuda.synchronize()
t = trace.cpu().numpy().astype(np.float64)
nkv = (a.ntok + 127) // 128
n_ctas = min(2048, ((a.ntok + 127) // 128) * a.heads * a.batch)
t = t[:n_ctas]
ok = t[:, :, 2 + 5 * nkv + 1] > 0
t = t[ok]                                                     # [traced warps, 64]
d = np.diff(t[:, : 2 + 5 * nkv + 2], axis=1)
names = ["prologue (entry -> CTA sync)", "wait first S"]
per_tile = ["wait S", "TMEM load + release S", "max + exponentials", "wait previous P V", "rescale / store P / announce"]
life = t[:, 2 + 5 * nkv + 1] - t[:, 0]
print(f"{t.shape[0]} softmax warps traced, {nkv} key tiles; CTA life (entry -> stored) {life.mean():.0f} clk (min {life.min():.0f}, max {life.max():.0f})")
print(f"  prologue (entry -> CTA sync)        {d[:, 0].mean():8.0f} clk  {100 * d[:, 0].mean() / life.mean():5.1f} %")
print(f"  wait first S                        {d[:, 1].mean():8.0f} clk  {100 * d[:, 1].mean() / life.mean():5.1f} %")
tile = d[:, 1: 1 + 5 * nkv].reshape(-1, nkv, 5)
for k, nm in enumerate(per_tile):
    first = 1 if k == 0 else 0                                # tile 0's "wait S" is reported above
    v = tile[:, first:, k]
    print(f"  {nm:34s}  {v.mean():8.0f} clk per tile, {100 * v.sum(axis=1).mean() / life.mean():5.1f} % of the life  (tile 1: {tile[:, 1, k].mean():.0f}, last: {tile[:, -1, k].mean():.0f})")
print(f"  wait last P V                       {d[:, 1 + 5 * nkv].mean():8.0f} clk  {100 * d[:, 1 + 5 * nkv].mean() / life.mean():5.1f} %")
print(f"  normalise + store                   {d[:, 2 + 5 * nkv].mean():8.0f} clk  {100 * d[:, 2 + 5 * nkv].mean() / life.mean():5.1f} %")
print(f"  one full tile, all phases           {tile[:, 1:-1].sum(axis=2).mean():8.0f} clk")
