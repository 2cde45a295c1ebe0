#!/usr/bin/env python
"""A/B two builds of libmde_b200.so on the same box: the batch-64 ViT-L step timed with each, alternating, one process per run.
    python tools/ab_lib.py tools/_old_libmde_b200.so monocular_depth_estimation_trt_b200/libmde_b200.so [--rounds 3] [--precision fp16]"""
import argparse, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import ctypes
class _Tolerant(ctypes.CDLL):          # an older build may lack a per-kernel test entry point the current _lib declares
    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            if not name.startswith("mde_k_"):
                raise
            f = ctypes.CFUNCTYPE(ctypes.c_int)(lambda *a: -1)
            setattr(self, name, f)
            return f
ctypes.CDLL = _Tolerant
from monocular_depth_estimation_trt_b200 import _lib
_lib.LIB_PATH = LIB
from monocular_depth_estimation_trt_b200 import engine as E, weights as W
from oracle import dav2_torch as O
B = 64
meta = W.describe("vitl", 518, 518, 20.0)
eng = E.Engine(E.make_desc(meta, precision=PREC, batch=B, input_mode="u8_hwc", max_src_hw=(480, 640)), meta)
eng.load_state_dict(O.init_state_dict("vitl", 0)); eng.finalize()
ctx = eng.create_execution_context()
ctx.set_input_shape("input", (B, 480, 640, 3))
src = torch.randint(0, 256, (B, 480, 640, 3), dtype=torch.uint8, device="cuda")
out = torch.empty(B, 518, 518, device="cuda")
ctx.set_tensor_address("input", src.data_ptr()); ctx.set_tensor_address("output", out.data_ptr())
st = torch.cuda.Stream(); torch.cuda.set_stream(st); s = st.cuda_stream
for _ in range(5): ctx.execute_async_v3(s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(15): ctx.execute_async_v3(s)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 15
ops = ctx.execute_timed(s)
grp = {}
for label, t, fl, by in ops:
    k = label.split(" ")[0]
    grp[k] = grp.get(k, 0.0) + t
print(json.dumps({"lib": os.path.basename(LIB), "ms": ms, "img_s": B * 1000 / ms, "checksum": float(out[0].double().mean()), "groups": {k: round(v, 3) for k, v in grp.items()}}))
'''
ap = argparse.ArgumentParser()
ap.add_argument("libs", nargs="+"); ap.add_argument("--rounds", type=int, default=3); ap.add_argument("--precision", default="fp16")
a = ap.parse_args()
for r in range(a.rounds):
    for lib in a.libs:
        code = f"ROOT={ROOT!r}; LIB={os.path.abspath(lib)!r}; PREC={a.precision!r}\n" + CHILD
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
        print(p.stdout.strip() or p.stderr[-800:], flush=True)
