#!/usr/bin/env python
"""Time the memory-bound kernels added for Depth Pro / VGGT alone (CUDA events) and print achieved GB/s over their
algorithmic bytes (operands read once, results written once).  python tools/elementwise_probe.py [--which all]"""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from monocular_depth_estimation_trt_b200 import depth_pro as DP, sharding as S, vggt as V
from monocular_depth_estimation_trt_b200.depth_pro import _Ops

ap = argparse.ArgumentParser(); ap.add_argument("--which", default="all"); ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
ops = _Ops("bf16"); ops.stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn, nbytes, label):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{label}: {ms * 1e3:.1f} us, {nbytes / ms / 1e6:.0f} GB/s over {nbytes / 1e6:.1f} MB")


if a.which in ("all", "rope"):
    frames, heads = 16, 16
    pos = torch.from_numpy(V.token_positions(37, 37)).repeat(frames, 1).cuda()
    rows, D = pos.shape[0], heads * 64
    qkv = torch.randn(rows, 3 * D, device="cuda").to(torch.bfloat16)
    w = [torch.ones(64, device="cuda"), torch.zeros(64, device="cuda"), torch.ones(64, device="cuda"), torch.zeros(64, device="cuda")]
    table = V.cos_sin_table(39).cuda()
    timed(lambda: ops.qknorm_rope(qkv, rows, heads, *w, 1e-5, pos, table, 39), rows * 2 * D * 2 * 2,
          f"qknorm_rope {rows} tokens x {heads} heads (q and k read + written)")
if a.which in ("all", "crops"):
    img = torch.randn(3, 1536, 1536, device="cuda")
    out = torch.empty(35, 3, 384, 384, device="cuda")
    plan = [(side, side, y0, x0) for _, side, y0, x0 in S.pyramid_plan(1536)]
    timed(lambda: ops.crops(img.data_ptr(), False, False, 1536, 1536, plan, out), img.numel() * 4 + out.numel() * 4, "resize_crops 35 crops of 1536x1536 (image read once + crops written)")
    src = torch.randint(0, 256, (2268, 3024, 3), dtype=torch.uint8, device="cuda")
    x = torch.empty(1, 3, 1536, 1536, device="cuda")
    timed(lambda: DP.preprocess_u8(src, 1536, x, stream_handle=torch.cuda.current_stream().cuda_stream), src.numel() + x.numel() * 4, "preprocess_u8 2268x3024 -> 1536x1536")
if a.which in ("all", "post"):
    inv = torch.rand(1536, 1536, device="cuda"); fov = torch.tensor([60.0], device="cuda")
    depth = torch.empty(2268, 3024, device="cuda")
    timed(lambda: DP.postprocess(inv.data_ptr(), fov.data_ptr(), 1536, 2268, 3024, depth, None, torch.cuda.current_stream().cuda_stream),
          inv.numel() * 4 + depth.numel() * 4, "depth_pro_post 1536x1536 -> 2268x3024")
