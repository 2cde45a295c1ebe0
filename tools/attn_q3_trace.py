#!/usr/bin/env python
"""Phase stamps of the three-query-tile attention kernel (attention_q3.cuh, trace instantiation) -> gpurun_out/attn_q3_trace.npz
and a summary of the first items of a few CTAs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import kutil as K
from monocular_depth_estimation_trt_b200 import _lib
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
lib = _lib.load()
B, N, H = 64, 1370, 16
qkv = torch.randn(B * N, 3 * H * 64, device="cuda").to(K.TORCH_DT[prec])
out = torch.empty(B * N, H * 64, dtype=K.TORCH_DT[prec], device="cuda")
nsm = torch.cuda.get_device_properties(0).multi_processor_count
trace = torch.zeros(nsm, 16, 256, dtype=torch.int64, device="cuda")
for _ in range(2):
    trace.zero_()
    _lib.check(lib.mde_k_attention_trace(_lib.PRECISIONS[prec] + 16, K.ptr(qkv), K.ptr(out), B, N, H, K.ptr(trace), K.stream()), "trace")
torch.cuda.synchronize()
tr = trace.cpu().numpy()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "attn_q3_trace.npz"), trace=tr)
nkv = (N + 95) // 96
for cta in (0, 77):
    t0 = tr[cta, 15, 0]
    print(f"--- CTA {cta}: producer (fetched, Q requested, K/V issued) x items:", (tr[cta, 15, :9] - t0).tolist())
    for t in range(3):
        print(f"  MMA issue times of query tile {t} (S0 PV0 S1 PV1 ...):", (tr[cta, 12 + t, :12] - t0).tolist())
    for w in (0, 4, 8):
        r = tr[cta, w, :5 * 6] - t0
        print(f"  softmax warp {w} (per key tile: S available, S in regs, exps done, P announced):")
        for j in range(6):
            print("     ", r[4 * j:4 * j + 4].tolist())
    per_item = 4 * nkv + 4
    w0 = tr[cta, 0]
    n_items = int((w0 > 0).sum()) // per_item
    ends = [int(w0[(i + 1) * per_item - 1] - t0) for i in range(min(n_items, 4))]
    print("  item ends (warp 0):", ends)
