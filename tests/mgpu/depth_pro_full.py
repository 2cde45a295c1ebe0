#!/usr/bin/env python
"""The whole Depth Pro model (BASELINE.json configs[3]: 1536 x 1536, 35 crops through a shared ViT-L/16 trunk) on 1..8
GPUs, one process per GPU: crops sharded over the ranks, taps all-gathered by the kernel that produces them (or by NCCL,
--gather nccl), decoder on every rank.  Prints one JSON line: latency per image on the device (CUDA events, max over ranks)
and end to end through allocate_buffers / do_inference with pinned host buffers, the stage breakdown of rank 0, and with
--check (ViT-S trunks, 64 decoder features) the parity against the unsharded CPU oracle.

    python tests/mgpu/depth_pro_full.py [--encoder vitl] [--precision fp16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        tests/mgpu/depth_pro_full.py --check
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

from monocular_depth_estimation_trt_b200 import common, depth_pro as DPE
from oracle import depth_pro_torch as DP      # weights + checker (test tooling)

ap = argparse.ArgumentParser()
ap.add_argument("--encoder", default="vitl"); ap.add_argument("--features", type=int, default=256)
ap.add_argument("--precision", default="fp16"); ap.add_argument("--gather", default="fused")
ap.add_argument("--check", action="store_true"); ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

if a.check:
    import refsetup as R
    a.encoder, a.features, hooks = "vits", 64, (8, 5)
    sd, x, inv, fov, _ = R.depth_pro_reference()
else:
    hooks = (11, 5)
    sd = DP.init_full_state_dict(a.encoder, features=a.features, seed=21)
    x = DP.preprocess(np.random.default_rng(0).integers(0, 256, (480, 640, 3), dtype=np.uint8), 1536)

result = {"model": "depth_pro", "encoder": a.encoder, "features": a.features, "precision": a.precision, "world": world,
          "gather": a.gather, "input": [1, 3, 1536, 1536]}


def max_over_ranks(v):
    if world == 1:
        return v
    t = torch.tensor([v], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


with DPE.DepthProEngine(sd, encoder=a.encoder, features=a.features, precision=a.precision, hook_blocks=hooks, world=world, rank=rank,
                        gather=a.gather, device=local) as engine, engine.create_execution_context() as context:
    inputs, outputs, bindings, stream = common.allocate_buffers(engine)
    inputs[0].host = x.numpy()
    for _ in range(3):
        outs = common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)
    got_inv, got_fov = outs[0].reshape(1536, 1536).copy(), float(outs[1][0])
    result["launches"] = context.launches_per_enqueue
    # end to end: H2D of the float32 input + forward + D2H of both outputs, host clock around the blocking call
    import time
    e2e = []
    for _ in range(a.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)
        e2e.append(max_over_ranks((time.perf_counter() - t0) * 1e3))
    # device only: inputs resident, CUDA events on the inference stream
    from cuda.bindings import runtime as cudart
    dev = []
    ev0, ev1 = common.cuda_call(cudart.cudaEventCreate()), common.cuda_call(cudart.cudaEventCreate())
    for _ in range(a.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        common.cuda_call(cudart.cudaEventRecord(ev0, stream))
        context.execute_async_v3(stream_handle=stream)
        common.cuda_call(cudart.cudaEventRecord(ev1, stream))
        common.cuda_call(cudart.cudaStreamSynchronize(stream))
        dev.append(max_over_ranks(float(common.cuda_call(cudart.cudaEventElapsedTime(ev0, ev1)))))
    context.profile_stages = True
    context.execute_async_v3(stream_handle=stream)
    common.cuda_call(cudart.cudaStreamSynchronize(stream))
    stages = context.stage_times()
    context.profile_stages = False
    common.free_buffers(inputs, outputs, stream)

result.update(device_ms_p50=float(np.median(dev)), device_ms_min=float(np.min(dev)), e2e_ms_p50=float(np.median(e2e)),
              h2d_bytes=int(x.numel() * 4), d2h_bytes=1536 * 1536 * 4 + 4, stages_ms_rank0={k: round(v, 4) for k, v in stages})
if a.check:
    m = R.compare_depth(inv.numpy(), got_inv)
    result["parity"] = {"abs_rel": m["abs_rel"], "max_rel": m["max_rel"], "fov_err_deg": abs(got_fov - float(fov))}
    gate = {"fp16": (2e-3, 1e-2, 0.05), "bf16": (1.2e-2, 1.2e-1, 0.5)}[a.precision]
    ok = m["abs_rel"] <= gate[0] and m["max_rel"] <= gate[1] and abs(got_fov - float(fov)) <= gate[2]
    if world > 1:
        # every rank must hold the same outputs
        t = torch.from_numpy(got_inv).cuda()
        ref0 = t.clone(); dist.broadcast(ref0, 0)
        ok = ok and bool(torch.equal(t, ref0))
        flag = torch.tensor([int(ok)], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
    result["ok"] = ok
if rank == 0:
    print(json.dumps(result))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if result.get("ok", True) else 1)
