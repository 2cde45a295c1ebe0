"""The two PARTITIONED configurations of BASELINE.json (configs[3] Depth Pro, configs[4] VGGT) as legs of bench.py: strong
scaling on the GPUs of one box (fixed work -- 16 frames / 35 crops -- split over N ranks), device-timed with CUDA events,
max over ranks, and every rank's output checked against the UNSHARDED fp32 oracle on a small width (test tooling: the
oracle provides seeded weights in set-up and is the checker after the timed regions).

    vggt_aggregator   24 x (frame block, global block) at ViT-L width, 16 frames x 1374 tokens: frames sharded by rank, the
                      K|V all-gather of every global block fused into the qk-norm + RoPE kernel's stores (peer memory over
                      NVLink, flag hand-shake on the stream, whole forward replayed as one CUDA graph); "nccl" = the same
                      pipeline with all_gather_into_tensor
    vggt_model        the whole exported model around that aggregator (DINOv2-L-with-registers trunk, camera / register tokens,
                      depth head), 16 frames of 518 x 518
    depth_pro         the whole model at 1536 x 1536 (ViT-L/16 trunks): 35 crops sharded by rank, taps all-gathered by the
                      kernel that produces them (or NCCL), decoder on every rank

Called by bench.py on every rank (collectives inside); returns the dict that goes under "partitioned" (rank 0) or None.
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def _max_over_ranks(v: float, world: int) -> float:
    if world == 1:
        return v
    import torch
    import torch.distributed as dist
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(world: int) -> None:
    import torch
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def _time_vggt(agg, x, stream, world: int, graph: bool, reps: int) -> float:
    import torch
    for _ in range(2):
        agg.forward(x.data_ptr(), stream)
    torch.cuda.synchronize()
    if graph:
        agg.capture(x.data_ptr(), stream)
        run = lambda: agg.replay(stream)
    else:
        run = lambda: agg.forward(x.data_ptr(), stream)
    for _ in range(2):
        run()
    ts = []
    for _ in range(reps):
        _barrier(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ts.append(_max_over_ranks(e0.elapsed_time(e1), world))
    return sorted(ts)[len(ts) // 2]


def vggt_aggregator(world: int, rank: int, local: int, precision: str, reps: int = 5) -> dict:
    import torch
    from monocular_depth_estimation_trt_b200 import vggt as P
    from oracle import vggt_torch as V
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    out = {"precision": precision, "world": world, "scaling": "strong"}
    with torch.cuda.stream(side):
        stream = side.cuda_stream
        # ---- parity on a small width: every rank's taps against the unsharded oracle
        dim, depth, grid = 384, 2, 37
        frames = max(4, world)
        H, N = dim // 64, 5 + grid * grid
        sd = V.init_aggregator(dim, depth, seed=4)
        torch.manual_seed(5)
        tok = torch.randn(frames, N, dim)
        per = frames // world
        ref = V.aggregate(sd, tok, grid, grid, H, depth)
        parity = {}
        for mode in (("fused", "nccl") if world > 1 else ("fused",)):
            agg = P.Aggregator(sd, dim, depth, H, grid, grid, frames_total=frames, precision=precision, world=world, rank=rank,
                               gather=mode, taps=list(range(depth)), device=local)
            x = tok[rank * per:(rank + 1) * per].contiguous().cuda()
            agg.forward(x.data_ptr(), stream)
            torch.cuda.synchronize()
            worst = 0.0
            for t in range(depth):
                got = agg.tap_out[t].cpu().reshape(per, N, 2 * dim).double()
                r = ref[t][rank * per:(rank + 1) * per].double()
                worst = max(worst, float(((got - r) ** 2).mean().sqrt() / (r ** 2).mean().sqrt()) / (t + 1))
            parity[mode] = _max_over_ranks(worst, world)
            agg.close()
            _barrier(world)
        gate = {"fp16": 1.2e-3, "bf16": 9e-3}[precision]
        out["parity_rms"] = {"worst_rms_rel_per_layer_over_ranks": parity, "gate": gate,
                             "what": f"dim {dim}, {depth}+{depth} blocks, {frames} frames sharded over {world} rank(s) vs the unsharded fp32 oracle",
                             "ok": all(v < gate for v in parity.values())}
        # ---- timing at the model's size
        dim, depth, frames = 1024, 24, 16
        H, N = dim // 64, 5 + grid * grid
        sd = V.init_aggregator(dim, depth, seed=4)
        torch.manual_seed(5)
        per = frames // world
        x = torch.randn(per, N, dim).cuda()
        flops = depth * 2 * (24 * frames * N * dim ** 2) + depth * 4 * dim * (frames * N * N + (frames * N) ** 2)
        out.update(frames=frames, tokens_per_frame=N, dim=dim, blocks=f"{depth}+{depth}", algorithmic_tflop=flops / 1e12,
                   kv_bytes_gathered_per_global_layer=frames * N * 2 * dim * 2)
        for mode in (("fused", "nccl") if world > 1 else ("fused",)):
            agg = P.Aggregator(sd, dim, depth, H, grid, grid, frames_total=frames, precision=precision, world=world, rank=rank,
                               gather=mode, taps=[4, 11, 17, 23], device=local)
            ms = _time_vggt(agg, x, stream, world, graph=(mode == "fused"), reps=reps)
            key = "ms" if mode == "fused" else "ms_nccl_gather"
            out[key] = ms
            if mode == "fused":
                out["tflops_per_gpu"] = flops / 1e12 / (ms / 1e3) / world
                out["launches"] = agg.ops.launches
                out["cuda_graph"] = True
            agg.close()
            _barrier(world)
        del sd
    torch.cuda.empty_cache()
    return out


def vggt_model(world: int, rank: int, local: int, precision: str, reps: int = 5) -> dict:
    """The WHOLE exported VGGT model (trunk with registers -> camera / register tokens -> aggregator -> depth head), frames
    sharded by rank: parity of every rank's depth maps against the unsharded oracle at ViT-S widths, then device time of one
    16-frame scene at the model's size (ViT-L, 24 + 24 blocks)."""
    import numpy as np
    import torch
    import refsetup as R
    from cuda.bindings import runtime as cudart
    from monocular_depth_estimation_trt_b200 import common, vggt as P
    from oracle import preprocess_np as PP, vggt_torch as V
    out = {"precision": precision, "world": world, "scaling": "strong"}

    def run(sd, imgs_local, cfg, frames, timed):
        with P.VGGTEngine(sd, frames=frames, precision=precision, world=world, rank=rank, device=local, **cfg) as engine, \
                engine.create_execution_context() as context:
            inputs, outputs, bindings, stream = common.allocate_buffers(engine)
            inputs[0].host = imgs_local.numpy()
            for _ in range(2):
                res = common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)
            got = res[0].reshape(imgs_local.shape[0], 518, 518).copy()
            ms = None
            if timed:
                ev0, ev1 = common.cuda_call(cudart.cudaEventCreate()), common.cuda_call(cudart.cudaEventCreate())
                ts = []
                for _ in range(reps):
                    _barrier(world)
                    common.cuda_call(cudart.cudaEventRecord(ev0, stream))
                    context.execute_async_v3(stream_handle=stream)
                    common.cuda_call(cudart.cudaEventRecord(ev1, stream))
                    common.cuda_call(cudart.cudaStreamSynchronize(stream))
                    ts.append(_max_over_ranks(float(common.cuda_call(cudart.cudaEventElapsedTime(ev0, ev1))), world))
                ms = float(np.median(ts))
            launches = context.launches_per_enqueue
            common.free_buffers(inputs, outputs, stream)
        return got, ms, launches

    # ---- parity (ViT-S widths, 4 + 4 blocks): rank 0 runs the CPU oracle and hands weights, frames and reference to the others
    frames = {1: 3, 2: 4}.get(world, world)
    small = dict(encoder="vits", depth=4, features=64, out_channels=(48, 96, 192, 384), taps=(0, 1, 2, 3))

    def reference():
        sd = V.init_vggt("vits", depth=4, features=64, out_channels=(48, 96, 192, 384), seed=0)
        imgs = torch.cat([torch.from_numpy(PP.preprocess_square_pad_cubic(
            np.random.default_rng(i).integers(0, 256, (480, 640, 3), dtype=np.uint8), 518, 518))[0] for i in range(frames)])
        V.calibrate_vggt(sd, imgs, "vits", 4, small["taps"])
        return sd, imgs, V.vggt_depth(sd, imgs, "vits", 4, small["taps"])

    if world > 1:
        import torch.distributed as dist
        box = [reference() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        sd, imgs, ref = box[0]
    else:
        sd, imgs, ref = reference()
    per = frames // world
    got, _, _ = run(sd, imgs[rank * per:(rank + 1) * per].contiguous(), small, frames, False)
    worst_abs = worst_max = 0.0
    for s in range(per):
        m = R.compare_depth(ref[rank * per + s].numpy(), got[s])
        worst_abs, worst_max = max(worst_abs, m["abs_rel"]), max(worst_max, m["max_rel"])
    worst_abs, worst_max = _max_over_ranks(worst_abs, world), _max_over_ranks(worst_max, world)
    out["parity"] = {"abs_rel": worst_abs, "max_rel": worst_max, "gate": {"abs_rel": 2e-3, "max_rel": 1e-2},
                     "meets_gate": bool(worst_abs <= 2e-3 and worst_max <= 1e-2),
                     "what": f"ViT-S widths, 4 + 4 blocks, {frames} frames of 518 x 518 sharded over {world} rank(s): every rank's depth maps "
                             f"vs the unsharded fp32 oracle (oracle/vggt_torch.py, parity unpinned against upstream VGGT)"}
    _barrier(world)
    # ---- timing at the model's size
    frames = 16
    per = frames // world
    sd = V.init_vggt("vitl", depth=24, features=256, out_channels=(256, 512, 1024, 1024), seed=0)
    torch.manual_seed(1)
    imgs = torch.rand(per, 3, 518, 518)
    big = dict(encoder="vitl", depth=24, features=256, out_channels=(256, 512, 1024, 1024), taps=(4, 11, 17, 23))
    _, ms, launches = run(sd, imgs, big, frames, True)
    out.update(frames=frames, image=[518, 518], ms=ms, frames_per_s=frames / (ms / 1e3), launches_per_rank=launches,
               what="trunk (DINOv2-L with registers) + 24 x (frame, global) blocks + depth head, device time, CUDA events, max over ranks")
    del sd
    torch.cuda.empty_cache()
    return out


def depth_pro(world: int, rank: int, local: int, precision: str, reps: int = 10) -> dict:
    import numpy as np
    import torch
    import refsetup as R
    from cuda.bindings import runtime as cudart
    from monocular_depth_estimation_trt_b200 import common, depth_pro as DPE
    from oracle import depth_pro_torch as DP
    out = {"precision": precision, "world": world, "scaling": "strong", "input": [1, 3, 1536, 1536], "crops": 35}

    def run(sd, x, encoder, features, hooks, mode, timed):
        with DPE.DepthProEngine(sd, encoder=encoder, features=features, precision=precision, hook_blocks=hooks, world=world, rank=rank,
                                gather=mode, device=local) as engine, engine.create_execution_context() as context:
            inputs, outputs, bindings, stream = common.allocate_buffers(engine)
            inputs[0].host = x.numpy()
            for _ in range(3):
                outs = common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)
            got_inv, got_fov = outs[0].reshape(1536, 1536).copy(), float(outs[1][0])
            ms = None
            if timed:
                ev0, ev1 = common.cuda_call(cudart.cudaEventCreate()), common.cuda_call(cudart.cudaEventCreate())
                ts = []
                for _ in range(reps):
                    _barrier(world)
                    common.cuda_call(cudart.cudaEventRecord(ev0, stream))
                    context.execute_async_v3(stream_handle=stream)
                    common.cuda_call(cudart.cudaEventRecord(ev1, stream))
                    common.cuda_call(cudart.cudaStreamSynchronize(stream))
                    ts.append(_max_over_ranks(float(common.cuda_call(cudart.cudaEventElapsedTime(ev0, ev1))), world))
                ms = float(np.median(ts))
            launches = context.launches_per_enqueue
            common.free_buffers(inputs, outputs, stream)
        return got_inv, got_fov, ms, launches

    # ---- parity: ViT-S trunks, 64 decoder features, every rank against the unsharded oracle; identical on every rank.
    # Rank 0 runs the CPU oracle (25 s of host time) and hands the seeded, calibrated weights and the reference to the others.
    if world > 1:
        import torch.distributed as dist
        box = [R.depth_pro_reference()[:4] if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        sd, x, inv, fov = box[0]
    else:
        sd, x, inv, fov, _ = R.depth_pro_reference()
    parity = {}
    for mode in (("fused", "nccl") if world > 1 else ("fused",)):
        got_inv, got_fov, _, _ = run(sd, x, "vits", 64, (8, 5), mode, False)
        m = R.compare_depth(inv.numpy(), got_inv)
        same = True
        if world > 1:
            import torch.distributed as dist
            t = torch.from_numpy(got_inv).cuda()
            ref0 = t.clone(); dist.broadcast(ref0, 0)
            flag = torch.tensor([int(torch.equal(t, ref0))], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            same = bool(flag.item())
        parity[mode] = {"abs_rel": _max_over_ranks(m["abs_rel"], world), "max_rel": _max_over_ranks(m["max_rel"], world),
                        "fov_err_deg": _max_over_ranks(abs(got_fov - float(fov)), world), "identical_on_every_rank": same}
        _barrier(world)
    out["parity"] = {"modes": parity, "gate": {"abs_rel": 2e-3, "max_rel": 1e-2},
                     "what": "ViT-S trunks, 64 decoder features, whole model vs the unsharded fp32 oracle (pinned on transformers' DepthPro)",
                     "ok": all(v["abs_rel"] <= 2e-3 and v["max_rel"] <= 1e-2 and v["identical_on_every_rank"] for v in parity.values())}
    # ---- timing: ViT-L trunks, 256 decoder features
    sd = DP.init_full_state_dict("vitl", features=256, seed=21)
    x = DP.preprocess(np.random.default_rng(0).integers(0, 256, (480, 640, 3), dtype=np.uint8), 1536)
    for mode in (("fused", "nccl") if world > 1 else ("fused",)):
        _, _, ms, launches = run(sd, x, "vitl", 256, (11, 5), mode, True)
        out["ms" if mode == "fused" else "ms_nccl_gather"] = ms
        if mode == "fused":
            out["launches"] = launches
        _barrier(world)
    del sd
    torch.cuda.empty_cache()
    return out


def run_all(world: int, rank: int, local: int, precision: str) -> dict | None:
    import torch
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(1, world)))      # torchrun pins OMP_NUM_THREADS=1: the oracle legs need more
    res = {"vggt_aggregator": vggt_aggregator(world, rank, local, precision),
           "vggt_model": vggt_model(world, rank, local, precision),
           "depth_pro": depth_pro(world, rank, local, precision)}
    return res if rank == 0 else None
