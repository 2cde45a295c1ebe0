#!/usr/bin/env python
"""Batch-1 latency of the ViT-L 518x518 engine, the reference's protocol (wall clock of do_inference incl. H2D + D2H, warm-up 20,
100 iterations, nearest-rank percentiles), plus the device time of one graph replay.
    python tools/b1_latency.py [--precision fp16] [--split-k] [--lib path/to/libmde_b200.so]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp16"); ap.add_argument("--split-k", action="store_true"); ap.add_argument("--lib", default="")
ap.add_argument("--encoder", default="vitl")
a = ap.parse_args()
import ctypes
if a.lib:
    class _Tolerant(ctypes.CDLL):
        def __getattr__(self, name):
            try:
                return super().__getattr__(name)
            except AttributeError:
                if not name.startswith("mde_k_"):
                    raise
                f = ctypes.CFUNCTYPE(ctypes.c_int)(lambda *x: -1)
                setattr(self, name, f)
                return f
    ctypes.CDLL = _Tolerant
import numpy as np, torch
from monocular_depth_estimation_trt_b200 import _lib
if a.lib:
    _lib.LIB_PATH = os.path.abspath(a.lib)
from monocular_depth_estimation_trt_b200 import common, engine as E, weights as W
from oracle import dav2_torch as O, harness_np as H
SRC_HW = (480, 640)
meta = W.describe(a.encoder, 518, 518, 20.0)
e1 = E.Engine(E.make_desc(meta, precision=a.precision, batch=1, input_mode="u8_hwc", max_src_hw=SRC_HW, split_k=a.split_k), meta)
e1.load_state_dict(O.init_state_dict(a.encoder, 0)); e1.finalize()
c1 = e1.create_execution_context()
c1.set_input_shape("input", (1, SRC_HW[0], SRC_HW[1], 3))
i1, o1, b1, s1 = common.allocate_buffers(e1, (1, 518, 518), profile_idx=0)
rng = np.random.default_rng(0)
i1[0].host = rng.integers(0, 256, size=(1, SRC_HW[0], SRC_HW[1], 3), dtype=np.uint8)
_, samples = H.measure(lambda: common.do_inference(c1, engine=e1, bindings=b1, inputs=i1, outputs=o1, stream=s1), warmup=20, iterations=100, sync=lambda: None)
st = H.stats(samples, warmup=20)
# device time of the replayed graph alone
src = torch.randint(0, 256, (1, SRC_HW[0], SRC_HW[1], 3), dtype=torch.uint8, device="cuda")
out = torch.empty(1, 518, 518, device="cuda")
c1.set_tensor_address("input", src.data_ptr()); c1.set_tensor_address("output", out.data_ptr())
s = torch.cuda.current_stream().cuda_stream
for _ in range(20): c1.execute_async_v3(s)
torch.cuda.synchronize()
e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): c1.execute_async_v3(s)
e1_.record(); torch.cuda.synchronize()
grp = {}
for label, t, fl, by in c1.execute_timed(s):
    k = " ".join(label.split(" ")[:2])
    grp[k] = grp.get(k, 0.0) + t
print(json.dumps({"lib": os.path.basename(_lib.LIB_PATH), "split_k": a.split_k, "p50_ms": st["p50_ms"], "p90_ms": st["p90_ms"], "min_ms": st["min_ms"],
                  "device_ms_back_to_back": e0.elapsed_time(e1_) / 100, "launches": c1.launches_per_enqueue,
                  "checksum": float(out.double().mean()),
                  "groups_ms": {k: round(v, 4) for k, v in sorted(grp.items(), key=lambda kv: -kv[1])[:14]}}))
