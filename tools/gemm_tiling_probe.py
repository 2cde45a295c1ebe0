#!/usr/bin/env python
"""Time the encoder GEMMs of a SMALL batch with every tile width / pairing / split-K count forced (mde_k_gemm_tiled), next to
what the library picks: the data behind csrc/kernels.cu `pick_tiling`.  20 launches replayed from a CUDA graph, CUDA events around the replay.
    python tools/gemm_tiling_probe.py [--batch 1] [--precision fp16] [--tokens 1370] [--dim 1024]"""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kutil as K
from monocular_depth_estimation_trt_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1); ap.add_argument("--precision", default="fp16")
ap.add_argument("--tokens", type=int, default=1370); ap.add_argument("--dim", type=int, default=1024)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
lib = _lib.load()
dt = K.TORCH_DT[a.precision]; M, D = a.batch * a.tokens, a.dim
dev = "cuda"
def mk(r, c): return (torch.randn(r, c, device=dev) * 0.5).to(dt)
def timed(fn):
    """device time per launch: a.reps launches recorded into one CUDA graph (the host needs ~15 us per ctypes call, more than
    these kernels run at batch 1), replayed and timed with events"""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(a.reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps * 1000.0
for case in ("qkv", "proj", "fc1", "fc2"):
    if case == "qkv":   n, k, kw = 3 * D, D, dict(bias=True, out=True)
    elif case == "proj": n, k, kw = D, D, dict(bias=True, gamma=True, x=True)
    elif case == "fc1":  n, k, kw = 4 * D, D, dict(bias=True, act=1, out=True)
    else:                n, k, kw = D, 4 * D, dict(bias=True, gamma=True, x=True)
    A, Bm = mk(M, k), mk(n, k)
    bias = torch.randn(n, device=dev)
    gamma = torch.rand(n, device=dev) * 0.01 if kw.get("gamma") else None
    x = torch.randn(M, n, device=dev) if kw.get("x") else None
    out = torch.empty(M, n, dtype=dt, device=dev) if kw.get("out") else None
    ep = K.epilogue(bias=bias, gamma=gamma, act=kw.get("act", 0), x=x, accumulate_x=bool(kw.get("x")), out=out, ld_out=n)
    res = [("picked", timed(lambda: K.gemm(a.precision, A, Bm, ep)))]
    for bn in (256, 128, 64):
        for ctas in (2, 1):
            if ctas == 2 and bn < 128: continue
            for splits in ((1, 2, 3, 4) if kw.get("x") else (1,)):
                if splits > 1 and bn < 128: continue
                def f(): _lib.check(lib.mde_k_gemm_tiled(_lib.PRECISIONS[a.precision], K.ptr(A), M, k, A.stride(0), K.ptr(Bm), n, Bm.stride(0),
                                                          C.byref(ep), bn, ctas, splits, K.stream()), "mde_k_gemm_tiled")
                res.append((f"bn{bn} x{ctas} split{splits}", timed(f)))
    best = min(t for _, t in res)
    print(f"{case:5s} M={M} N={n} K={k}: " + "  ".join(f"{name} {t:.1f}us{'*' if t == best else ''}" for name, t in res), flush=True)
