#!/usr/bin/env python
"""VGGT's aggregator (alternating frame / global attention blocks with qk-norm and 2-D RoPE) on N GPUs, one process per GPU,
frames sharded by rank.  Global blocks: the kernel that finishes K stores K|V into every rank's gathered buffer, a flag
hand-shake on the stream orders the ranks, attention reads local queries against the gathered keys / values ("fused");
--gather nccl is the all_gather_into_tensor baseline.  Rank 0 prints one JSON line: ms per forward (CUDA events, max over
ranks), and with --check the parity of every rank's taps against the UNSHARDED fp32 oracle.

    python tests/mgpu/vggt_aggregator.py --frames 16 --depth 24                       # one GPU, ViT-L width
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 \
        tests/mgpu/vggt_aggregator.py --check
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist

from monocular_depth_estimation_trt_b200 import vggt as P
from oracle import vggt_torch as V      # weights + checker (test tooling)

ap = argparse.ArgumentParser()
ap.add_argument("--dim", type=int, default=1024); ap.add_argument("--depth", type=int, default=24)
ap.add_argument("--frames", type=int, default=16); ap.add_argument("--grid", type=int, default=37)
ap.add_argument("--precision", default="bf16"); ap.add_argument("--gather", default="fused")
ap.add_argument("--check", action="store_true"); ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--graph", action="store_true", help="capture the forward into a CUDA graph and time replays")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if a.check:
    a.dim, a.depth, a.frames = 384, 2, 4
H, N = a.dim // 64, 5 + a.grid * a.grid
sd = V.init_aggregator(a.dim, a.depth, seed=4)
torch.manual_seed(5)
tok = torch.randn(a.frames, N, a.dim)
per = a.frames // world
taps = tuple(range(a.depth)) if a.check else (4, 11, 17, 23)
agg = P.Aggregator(sd, a.dim, a.depth, H, a.grid, a.grid, frames_total=a.frames, precision=a.precision, world=world, rank=rank,
                   gather=a.gather, taps=[t for t in taps if t < a.depth], device=local)
x = tok[rank * per:(rank + 1) * per].contiguous().cuda()
ap_stream = torch.cuda.Stream()
torch.cuda.set_stream(ap_stream)
stream = ap_stream.cuda_stream
for _ in range(2):
    agg.forward(x.data_ptr(), stream)
torch.cuda.synchronize()
graphed = a.graph and (world == 1 or a.gather == "fused")
if graphed:
    agg.capture(x.data_ptr(), stream)            # the whole sharded forward, hand-shakes included, as one graph launch
    run = lambda: agg.replay(stream)
else:
    run = lambda: agg.forward(x.data_ptr(), stream)
for _ in range(2):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(a.reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ts.append(float(t.item()))
flops = a.depth * 2 * (24 * a.frames * N * a.dim ** 2) + a.depth * 4 * a.dim * (a.frames * N * N + (a.frames * N) ** 2)
ms = sorted(ts)[len(ts) // 2]
result = {"model": "vggt aggregator", "dim": a.dim, "depth": a.depth, "frames": a.frames, "tokens_per_frame": N, "precision": a.precision,
          "world": world, "gather": a.gather, "cuda_graph": bool(graphed), "ms_per_forward": ms, "launches": agg.ops.launches,
          "algorithmic_tflop": flops / 1e12, "tflops_per_gpu": flops / 1e12 / (ms / 1e3) / world,
          "kv_bytes_gathered_per_global_layer": a.frames * N * 2 * a.dim * 2}
ok = True
if a.check:
    ref = V.aggregate(sd, tok, a.grid, a.grid, H, a.depth)
    worst = 0.0
    for t in range(a.depth):
        got = agg.tap_out[t].cpu().reshape(per, N, 2 * a.dim).double()
        r = ref[t][rank * per:(rank + 1) * per].double()
        worst = max(worst, float(((got - r) ** 2).mean().sqrt() / (r ** 2).mean().sqrt()) / (t + 1))
    gate = {"fp16": 1.2e-3, "bf16": 9e-3}[a.precision]
    w = torch.tensor([worst], device="cuda")
    if world > 1:
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
    result["worst_rms_rel_per_layer_over_ranks"] = float(w.item())
    ok = float(w.item()) < gate
    result["ok"] = ok
if rank == 0:
    print(json.dumps(result))
agg.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
