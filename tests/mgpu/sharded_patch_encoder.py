#!/usr/bin/env python
"""Depth Pro patch-encoder stage on N GPUs (one process per GPU, launch with torchrun): 35 crops of a synthetic
1536 x 1536 image sharded over the ranks, trunk-only ViT/16 engine per rank, taps all-gathered
(a) by the tap kernel itself over cudaIpc-mapped peer memory and (b) by NCCL.  Rank 0 checks both against the
unsharded CPU oracle and prints one JSON line with the timings (CUDA events, max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/mgpu/sharded_patch_encoder.py [--encoder vitl] [--precision bf16] [--check]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from monocular_depth_estimation_trt_b200 import engine as E, sharding as S, weights as W
from oracle import dav2_torch as O      # weights + checker (test tooling)

ap = argparse.ArgumentParser()
ap.add_argument("--encoder", default="vits"); ap.add_argument("--precision", default="fp16")
ap.add_argument("--check", action="store_true"); ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

torch.manual_seed(0)
image = torch.randn(3, 1536, 1536)
crops = S.make_crops(image)                                   # [35, 3, 384, 384], identical on every rank
per, bounds = S.shard_bounds(crops.shape[0], world)
first, count = bounds[rank]
mine = torch.zeros(per, 3, 384, 384)
mine[:count] = crops[first:first + count]
sd = O.init_state_dict(a.encoder, seed=5, patch=16, pos_grid=24)
meta = W.describe(a.encoder, 384, 384, max_depth=None, patch_size=16)
eng = E.Engine(E.make_desc(meta, precision=a.precision, batch=per, head="encoder_taps", tap_norm_mask=0x8, device=local), meta)
eng.load_state_dict({k: v for k, v in sd.items() if k.startswith("pretrained.")})
eng.finalize()
xd = mine.cuda()
stream = torch.cuda.current_stream().cuda_stream
result = {"world": world, "encoder": a.encoder, "precision": a.precision, "crops": int(crops.shape[0]), "per_rank": per}
outs = {}
for mode in ("fused", "nccl"):
    enc = S.ShardedPatchEncoder(eng, crops.shape[0], world, rank, mode)
    for _ in range(2):
        enc.enqueue(xd.data_ptr(), stream); enc.finish()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        e0.record(); enc.enqueue(xd.data_ptr(), stream); e1.record()
        enc.finish()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    result[f"{mode}_ms"] = sorted(ts)[len(ts) // 2]
    outs[mode] = enc.gathered().clone()
    enc.close()
same = torch.tensor([int(torch.equal(outs["fused"], outs["nccl"]))], device="cuda")
dist.all_reduce(same, op=dist.ReduceOp.MIN)
result["fused_equals_nccl_on_every_rank"] = bool(same.item())
if a.check and rank == 0:
    with torch.no_grad():
        ref = torch.stack(O.encoder_taps(sd, crops, O.MODEL_CONFIGS[a.encoder], norm_mask=0x8))
    got = outs["fused"].float().cpu()
    result["rms_rel_vs_oracle"] = [float(((got[i] - ref[i]) ** 2).mean().sqrt() / (ref[i] ** 2).mean().sqrt()) for i in range(4)]
if rank == 0:
    print(json.dumps(result))
dist.barrier()
eng.close()
dist.destroy_process_group()
