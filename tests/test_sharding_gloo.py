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
