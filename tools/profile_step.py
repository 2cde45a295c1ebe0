#!/usr/bin/env python
"""Per-launch device times of one forward (CUDA events between launches), grouped by op label.
    python tools/profile_step.py [--encoder vitl] [--batch 64] [--precision bf16] [--h 518 --w 518]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from monocular_depth_estimation_trt_b200 import engine as E, weights as W

ap = argparse.ArgumentParser()
ap.add_argument("--encoder", default="vitl"); ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--precision", default="bf16"); ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--out", default=""); ap.add_argument("--h", type=int, default=518); ap.add_argument("--w", type=int, default=518)
a = ap.parse_args()
meta = W.describe(a.encoder, a.h, a.w, 20.0)
eng = E.Engine(E.make_desc(meta, precision=a.precision, batch=a.batch), meta)
from oracle import dav2_torch as O   # weights for a profiling run only (tooling, not a product path)
eng.load_state_dict(O.init_state_dict(a.encoder, 0)); eng.finalize()
ctx = eng.create_execution_context()
x = torch.randn(a.batch, 3, a.h, a.w, device="cuda"); out = torch.empty(a.batch, a.h, a.w, device="cuda")
ctx.set_tensor_address("input", x.data_ptr()); ctx.set_tensor_address("output", out.data_ptr())
s = torch.cuda.current_stream().cuda_stream
for _ in range(2): ctx.execute_async_v3(s)
torch.cuda.synchronize()
acc = {}
for _ in range(a.reps):
    for label, ms, fl, by in ctx.execute_timed(s):
        e = acc.setdefault(label, [0.0, 0.0, 0.0, 0])
        e[0] += ms / a.reps; e[1] += fl / a.reps; e[2] += by / a.reps; e[3] += 1
tot = sum(v[0] for v in acc.values())
lines = [f"{'op':62s} {'n':>4s} {'ms':>9s} {'%':>6s} {'TFLOP/s':>9s} {'GB/s':>8s}"]
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    lines.append(f"{k:62s} {v[3] // a.reps:4d} {v[0]:9.3f} {100 * v[0] / tot:6.2f} {v[1] / max(v[0], 1e-9) / 1e9:9.1f} {v[2] / max(v[0], 1e-9) / 1e6:8.1f}")
lines.append(f"total {tot:.3f} ms per step, {a.batch / tot * 1000:.1f} images/s (events between launches)")
txt = "\n".join(lines); print(txt)
if a.out: open(a.out, "w").write(txt + "\n")
