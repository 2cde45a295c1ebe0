"""The N > 1 path of bench.py on the CPU: two gloo ranks shard the synthetic images by rank (weak scaling, no
data-path collective) and the only collective is the max-over-ranks of the timings."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bench


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = bench.synthetic_batch(3, rank)                       # this rank's shard: global images 3*rank .. 3*rank+2
    my_ms = 100.0 + 25.0 * rank                                   # pretend rank 1 was slower
    slowest = bench.max_over_ranks(my_ms, world, device="cpu")
    sums = torch.tensor([float(frames[i].astype(np.int64).sum()) for i in range(3)], dtype=torch.float64)
    gathered = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, sums)                               # test-only: collect what each rank generated
    if rank == 0:
        out.put((slowest, [g.tolist() for g in gathered], bench.aggregate_rate(world, 3, 10, slowest)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_images_and_report_the_slowest_rank():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    slowest, gathered, rate = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert slowest == 125.0                                      # max over ranks, not rank 0's own time
    assert abs(rate - 2 * 3 * 10 / 0.125) < 1e-6                  # all ranks' images / slowest time
    flat = gathered[0] + gathered[1]
    assert len(set(flat)) == 6                                   # six distinct images: no overlap between shards
    want = [float(np.random.default_rng(i).integers(0, 256, (480, 640, 3), dtype=np.uint8).astype(np.int64).sum()) for i in range(6)]
    assert flat == want                                          # seed == global image index


def test_single_rank_needs_no_process_group():
    assert bench.max_over_ranks(3.5, 1) == 3.5
    assert bench.aggregate_rate(1, 64, 10, 1000.0) == 640.0


# ------------------------------------------------------------------ Depth Pro patch sharding (host logic)
from monocular_depth_estimation_trt_b200 import sharding as S


def test_depth_pro_pyramid_is_35_overlapping_crops():
    plan = S.pyramid_plan(1536)
    assert len(plan) == 35
    assert [sum(1 for p in plan if p[0] == lvl) for lvl in range(3)] == [25, 9, 1]
    assert S.crop_origins(1536, 0.25) == [0, 288, 576, 864, 1152]          # stride 288, last crop ends at 1536
    assert S.crop_origins(768, 0.5) == [0, 192, 384]
    assert S.crop_origins(384, 0.0) == [0]
    img = torch.arange(3 * 1536 * 1536, dtype=torch.float32).reshape(3, 1536, 1536)
    crops = S.make_crops(img)
    assert crops.shape == (35, 3, 384, 384)
    assert torch.equal(crops[7], img[:, 288:672, 576:960])                 # level 0 is cut from the image itself (crop 7 = row 1, col 2)
    assert torch.equal(crops[34], torch.nn.functional.interpolate(img[None], size=(384, 384), mode="bilinear", align_corners=False)[0])


def test_shard_bounds_pad_the_tail():
    assert S.shard_bounds(35, 1) == (35, [(0, 35)])
    assert S.shard_bounds(35, 2) == (18, [(0, 18), (18, 17)])
    assert S.shard_bounds(35, 4) == (9, [(0, 9), (9, 9), (18, 9), (27, 8)])
    assert S.shard_bounds(35, 8) == (5, [(0, 5), (5, 5), (10, 5), (15, 5), (20, 5), (25, 5), (30, 5), (35, 0)])
    assert S.shard_bounds(3, 8)[0] == 1


def _patch_worker(rank, world, port, out):
    """Each rank runs the ORACLE trunk on its shard of crops and the ranks all-gather the taps in the layout the CUDA
    path uses ([4][world * per_rank][T][D]); rank 0 checks the result against the unsharded run."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import dav2_torch as O
    torch.manual_seed(0)
    torch.set_num_threads(2)
    n_items = 5
    crops = torch.randn(n_items, 3, 64, 64)                      # tiny crops: 4 x 4 tokens of 16 x 16 pixels
    sd = O.init_state_dict("vits", seed=3, patch=16, pos_grid=4)
    cfg = O.MODEL_CONFIGS["vits"]
    per, bounds = S.shard_bounds(n_items, world)
    first, count = bounds[rank]
    mine = torch.zeros(per, 3, 64, 64)
    mine[:count] = crops[first:first + count]
    with torch.no_grad():
        taps = torch.stack(O.encoder_taps(sd, mine, cfg, norm_mask=0x8))        # [4, per, T, D]
    gathered = torch.zeros(4, world * per, 16, 384)
    for i in range(4):
        dist.all_gather_into_tensor(gathered[i], taps[i].contiguous())
    if rank == 0:
        with torch.no_grad():
            ref = torch.stack(O.encoder_taps(sd, crops, cfg, norm_mask=0x8))
        out.put(float((gathered[:, :n_items] - ref).abs().max()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_crops_and_all_gather_the_taps():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_patch_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    err = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-5          # batch composition changes fp32 summation order at most


# ------------------------------------------------------------------ VGGT: frames sharded by rank, K|V all-gathered per global block
def _vggt_worker(rank, world, port, out):
    """Each rank runs the ORACLE blocks on its frames.  Frame blocks are local.  In a global block every rank projects its
    own tokens, finishes q and k (qk-norm + RoPE), all-gathers K|V into the layout the CUDA path uses ([frames_total * N, 2D],
    rank r's rows at r * rows_local) and attends with its queries over all keys / values; rank 0 compares its frames with
    the unsharded oracle."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch.nn.functional as F
    from oracle import vggt_torch as V
    from monocular_depth_estimation_trt_b200 import vggt as P
    torch.set_num_threads(2)
    D, H, S_total, depth, gh, gw = 128, 2, 4, 2, 3, 3
    N = 5 + gh * gw
    sd = V.init_aggregator(D, depth, seed=2)
    torch.manual_seed(1)
    tokens = torch.randn(S_total, N, D)
    per = S_total // world
    t = tokens[rank * per:(rank + 1) * per].clone()
    pos = torch.from_numpy(P.token_positions(gh, gw)).long()            # the product's table, one frame
    rows = per * N

    def attend(pre, x, kv_from_all):
        B, n, _ = x.shape
        y = F.layer_norm(x, (D,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], V.LN_EPS)
        qkv = F.linear(y, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"]).reshape(B, n, 3, H, 64).permute(2, 0, 3, 1, 4)
        p_ = pos.repeat(n // N, 1)
        q = V.rope_2d(F.layer_norm(qkv[0], (64,), sd[pre + "attn.q_norm.weight"], sd[pre + "attn.q_norm.bias"], V.QK_EPS), p_)
        k = V.rope_2d(F.layer_norm(qkv[1], (64,), sd[pre + "attn.k_norm.weight"], sd[pre + "attn.k_norm.bias"], V.QK_EPS), p_)
        v = qkv[2]
        if kv_from_all:
            local = torch.cat([k.permute(0, 2, 1, 3).reshape(n, D), v.permute(0, 2, 1, 3).reshape(n, D)], dim=1).contiguous()   # [rows, 2D]
            gathered = torch.zeros(world * rows, 2 * D)
            dist.all_gather_into_tensor(gathered, local)                 # rank r's rows land at r * rows: the CUDA path's layout
            k = gathered[:, :D].reshape(1, world * rows, H, 64).permute(0, 2, 1, 3)
            v = gathered[:, D:].reshape(1, world * rows, H, 64).permute(0, 2, 1, 3)
        a = torch.softmax((q * 64 ** -0.5) @ k.transpose(-2, -1), dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, n, D)
        x = x + sd[pre + "ls1.gamma"] * F.linear(a, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
        y = F.layer_norm(x, (D,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], V.LN_EPS)
        return x + sd[pre + "ls2.gamma"] * F.linear(F.gelu(F.linear(y, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])),
                                                     sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])

    layers = []
    with torch.no_grad():
        for i in range(depth):
            t = attend(f"aggregator.frame_blocks.{i}.", t, False)
            frame = t
            t = attend(f"aggregator.global_blocks.{i}.", t.reshape(1, rows, D), True).reshape(per, N, D)
            layers.append(torch.cat([frame, t], dim=-1))
    if rank == 0:
        ref = V.aggregate(sd, tokens, gh, gw, H, depth)
        out.put(max(float((layers[i] - ref[i][:per]).abs().max()) for i in range(depth)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_frames_and_all_gather_keys_and_values():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_vggt_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    err = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 2e-5
