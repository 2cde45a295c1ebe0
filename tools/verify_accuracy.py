#!/usr/bin/env python
"""Engine-vs-reference accuracy report in the schema of the reference's reports/accuracy.json
(tools/verify_accuracy.py:150-194 there: one row per model with per-output rel_mean / abs_rel / rmse / max_abs / corr /
scale and a verdict).  The reference compares its TensorRT engine with the ONNX graph on reports/inputs/<model>.npy;
here the B200 engine is compared with the fp32 oracle forward on the same synthetic input and seeded weights.

    python tools/verify_accuracy.py [--encoders vits,vitl] [--precisions fp16,bf16] [--out profiles/r01_accuracy.json]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import refsetup as R            # seeded reference set-up + the reference's compare metrics (test tooling)
from monocular_depth_estimation_trt_b200 import engine as E, weights as W

WARN, FAIL = 0.01, 0.05         # the reference's thresholds on rel_mean (tools/verify_accuracy.py:54-55)


def verdict_for(o):
    structural = np.isfinite(o["corr"]) and o["corr"] < 0.99
    if o["rel_mean"] > FAIL:
        return "FAIL" if structural else "WARN"
    if o["rel_mean"] > WARN:
        return "WARN" if structural else "ok"
    return "ok"


def row(encoder, precision):
    sd, x, depth, _ = R.reference(encoder)
    meta = W.describe(encoder, 518, 518, 20.0)
    eng = E.Engine(E.make_desc(meta, precision=precision, batch=1), meta)
    eng.load_state_dict(sd); eng.finalize()
    out = torch.full((1, 518, 518), float("nan"), device="cuda")
    xd = x.cuda()
    with eng.create_execution_context() as ctx:
        ctx.set_tensor_address("input", xd.data_ptr()); ctx.set_tensor_address("output", out.data_ptr())
        ctx.execute_async_v3(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    eng.close()
    a, b = depth.numpy().astype(np.float64).ravel(), out.cpu().numpy().astype(np.float64).ravel()
    m = R.compare_depth(a, b)
    o = {"index": 0, "size": int(a.size), "ref_mag": float(np.abs(a).mean()), "rel_mean": m["rel_mean"], "abs_rel": m["abs_rel"],
         "rmse": float(np.sqrt(((a - b) ** 2).mean())), "max_abs": m["max_abs"], "max_rel": m["max_rel"], "corr": m["corr"],
         "scale": float((a * b).sum() / (b * b).sum())}
    o["verdict"] = verdict_for(o)
    return {"model": f"depth_anything_v2_{encoder}", "precision": precision, "reference": "oracle/dav2_torch.py fp32 forward (seeded calibrated init)",
            "input": "synthetic 480x640 uint8 seed 0 -> core/preprocess semantics -> [1,3,518,518]", "outputs": [o],
            "worst_rel": o["rel_mean"], "verdict": o["verdict"]}


ap = argparse.ArgumentParser()
ap.add_argument("--encoders", default="vits,vitl"); ap.add_argument("--precisions", default="fp16,bf16"); ap.add_argument("--out", default="")
a = ap.parse_args()
rows = [row(e, p) for e in a.encoders.split(",") for p in a.precisions.split(",")]
for r in rows:
    o = r["outputs"][0]
    print(f"{r['model']:24s} {r['precision']:5s} rel {o['rel_mean']:.2e}  abs_rel {o['abs_rel']:.2e}  max_rel {o['max_rel']:.2e}  corr {o['corr']:.6f}  {r['verdict']}")
if a.out:
    json.dump(rows, open(a.out, "w"), indent=1)
