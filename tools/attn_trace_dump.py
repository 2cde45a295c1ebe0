#!/usr/bin/env python
"""Raw phase stamps of the traced attention kernel (mde_k_attention_trace) -> gpurun_out/attn_trace.npz, for the offline
analysis of how the two CTAs that share an SM interact (tools/attn_trace_pairs.py)."""
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
trace = torch.zeros(2048, 4, 64, dtype=torch.int64, device="cuda")
for _ in range(2):
    trace.zero_()
    _lib.check(lib.mde_k_attention_trace(_lib.PRECISIONS[prec], K.ptr(qkv), K.ptr(out), B, N, H, K.ptr(trace), K.stream()), "trace")
torch.cuda.synchronize()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "attn_trace.npz"), trace=trace.cpu().numpy())
print("saved", trace.shape)
