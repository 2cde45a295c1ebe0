"""Patch sharding of Depth Pro's patch-encoder stage over the GPUs of one box (SURVEY section 8 e, row 2).

Depth Pro (models/depth_pro/onnx_export.py:15-22: `dinov2l16_384` trunks, 1536 x 1536 input) cuts its input pyramid
into 35 overlapping 384 x 384 crops -- 25 at full resolution (stride 288), 9 at half (stride 192), 1 at quarter --
and pushes all of them through ONE shared ViT-L/16 trunk, keeping the final tokens and two hooked block outputs.
The crops are independent, so they shard over ranks (one process per GPU, weights replicated); the only exchange is
an all-gather of the trunk's outputs, after which every rank holds all 35 and can run the decoder.

Two ways to do that exchange are provided:

* ``mode="fused"``   the kernel that produces a tap (final LayerNorm / conversion of the residual stream) writes each
                     row straight into every rank's gather buffer -- local HBM for its own, plain stores over
                     NVLink into cudaIpc-mapped peer memory for the others.  No collective launch, no staging copy;
                     the transfer overlaps the rest of the trunk (three of the four taps are produced long before
                     the last block finishes).
* ``mode="nccl"``    the trunk writes its own output binding and ``torch.distributed.all_gather_into_tensor`` (NCCL
                     over NVLink / NVSwitch) assembles the result: the baseline the fused path is measured against.

The host logic (crop geometry, shard bounds, buffer layout) is pure Python and is tested on the CPU with gloo.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

PATCH = 384
LEVELS = ((1.0, 0.25), (0.5, 0.5), (0.25, 0.0))     # (image scale, overlap ratio) of the three pyramid levels


def crop_origins(size: int, overlap: float, patch: int = PATCH) -> List[int]:
    """Top-left coordinates of the sliding crops along one axis (upstream `split`: stride = int(patch * (1 - overlap)),
    steps = ceil((size - patch) / stride) + 1)."""
    if size < patch:
        raise ValueError(f"image side {size} is smaller than the crop {patch}")
    if size == patch:
        return [0]
    stride = int(patch * (1.0 - overlap))
    steps = int(math.ceil((size - patch) / stride)) + 1
    return [min(i * stride, size - patch) for i in range(steps)]


def pyramid_plan(image_size: int = 1536) -> List[Tuple[int, int, int, int]]:
    """[(level, level_size, y0, x0)] for every crop, in upstream's order (level by level, row-major): 25 + 9 + 1 = 35."""
    plan = []
    for lvl, (scale, overlap) in enumerate(LEVELS):
        side = int(round(image_size * scale))
        for y0 in crop_origins(side, overlap):
            for x0 in crop_origins(side, overlap):
                plan.append((lvl, side, y0, x0))
    return plan


def make_crops(image):
    """image: torch float tensor [3, S, S] (already normalised) -> [35, 3, 384, 384].  The lower pyramid levels are
    bilinear down-samplings (align_corners=False), as upstream builds them."""
    import torch
    import torch.nn.functional as F
    S = image.shape[-1]
    levels = {}
    out = []
    for lvl, side, y0, x0 in pyramid_plan(S):
        if lvl not in levels:
            levels[lvl] = image if side == S else F.interpolate(image[None], size=(side, side), mode="bilinear", align_corners=False)[0]
        out.append(levels[lvl][:, y0:y0 + PATCH, x0:x0 + PATCH])
    return torch.stack(out)


def torch_stream(stream_handle):
    """The torch stream object of a raw CUDA stream handle, for ordering NCCL collectives on the caller's stream.  Handle 0
    is the legacy default stream: torch.cuda.ExternalStream(0) would silently create a NEW pool stream instead."""
    import torch
    h = int(stream_handle)
    return torch.cuda.default_stream() if h == 0 else torch.cuda.ExternalStream(h)


def shard_bounds(n_items: int, world: int) -> Tuple[int, List[Tuple[int, int]]]:
    """Equal-size shards: every rank owns `per_rank = ceil(n / world)` slots so that one static engine serves all of
    them; the trailing slots of the last rank(s) are padding.  -> (per_rank, [(first_item, n_real_items)] per rank)."""
    if n_items < 1 or world < 1:
        raise ValueError("need at least one item and one rank")
    per = -(-n_items // world)
    return per, [(r * per, max(0, min(per, n_items - r * per))) for r in range(world)]


class _DevBuf:
    """A cudaMalloc'd (IPC-shareable) buffer exposed through __cuda_array_interface__ so torch can view it."""

    def __init__(self, ptr: int, shape: Sequence[int], typestr: str):
        self.ptr = int(ptr)
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": typestr,
                                         "data": (self.ptr, False), "version": 2}


class PeerBuffers:
    """One 16-bit buffer of `shape` per rank, cudaMalloc'd and mapped into every process of the job with cudaIpc
    (`ptrs[r]` is rank r's buffer as seen from this process; `ptrs[rank]` is our own)."""

    def __init__(self, world: int, rank: int, shape: Sequence[int], precision: str):
        import torch.distributed as dist
        from cuda.bindings import runtime as cudart
        from . import common
        self._cudart, self._common = cudart, common
        self.world, self.rank = world, rank
        self.shape = tuple(int(v) for v in shape)
        self.precision = precision
        self.nbytes = int(np.prod(self.shape)) * 2
        self.own = int(common.cuda_call(cudart.cudaMalloc(self.nbytes)))
        common.cuda_call(cudart.cudaMemset(self.own, 0, self.nbytes))
        self.ptrs = [0] * world
        self.ptrs[rank] = self.own
        self._opened = []
        if world > 1:
            handle = common.cuda_call(cudart.cudaIpcGetMemHandle(self.own))
            blobs = [None] * world
            dist.all_gather_object(blobs, bytes(handle.reserved))
            for r, blob in enumerate(blobs):
                if r == rank:
                    continue
                h = cudart.cudaIpcMemHandle_t()
                h.reserved = blob
                p = int(common.cuda_call(cudart.cudaIpcOpenMemHandle(h, cudart.cudaIpcMemLazyEnablePeerAccess)))
                self.ptrs[r] = p
                self._opened.append(p)

    def view(self):
        """torch view of this rank's own buffer in the 16-bit type."""
        import torch
        t = torch.as_tensor(_DevBuf(self.own, self.shape, "<i2"), device="cuda")
        return t.view(torch.bfloat16 if self.precision == "bf16" else torch.float16)

    def close(self):
        for p in self._opened:
            self._cudart.cudaIpcCloseMemHandle(p)
        self._opened = []
        if self.own:
            self._cudart.cudaFree(self.own)
            self.own = 0


class PeerSync:
    """Stream-ordered replacement for `torch.cuda.synchronize(); dist.barrier()` around the fused gather kernels:
    two flag arrays per rank in peer-mapped memory -- "ready" (my stores into your buffer have landed) and "ack" (I have
    finished reading what you sent me) -- driven by two one-warp kernels on the caller's stream.

        sync.wait_acks(stream)      # before overwriting the peers' buffers again
        ... producer kernels ...    # e.g. the trunk with the fused tap gather, or the QKV GEMM with the fused K|V gather
        sync.signal_ready(stream); sync.wait_ready(stream)
        ... consumer kernels ...    # read the gathered buffer
        sync.signal_acks(stream)
    """

    def __init__(self, world: int, rank: int):
        from . import _lib
        self._lib, self._check = _lib.load(), _lib.check
        self.world, self.rank = world, rank
        self.ready = PeerBuffers(world, rank, (64,), "fp16")      # 128 zeroed bytes: room for 8 uint32 flags
        self.acks = PeerBuffers(world, rank, (64,), "fp16")
        # the round number lives in device memory (one uint32 of a private buffer): launch arguments never change, so a
        # forward that uses the hand-shake can be captured into a CUDA graph and replayed
        self.counter = PeerBuffers(1, 0, (64,), "fp16")

    def _signal(self, bufs, advance, stream_handle):
        import ctypes as C
        arr = (C.c_void_p * self.world)(*[C.c_void_p(int(p)) for p in bufs.ptrs])
        self._check(self._lib.mde_k_peer_signal_counter(arr, self.world, self.rank, C.c_void_p(self.counter.own), int(advance),
                                                        C.c_void_p(int(stream_handle))), "mde_k_peer_signal_counter")

    def _wait(self, bufs, stream_handle):
        import ctypes as C
        self._check(self._lib.mde_k_peer_wait_counter(C.c_void_p(bufs.own), self.world, C.c_void_p(self.counter.own),
                                                      C.c_void_p(int(stream_handle))), "mde_k_peer_wait_counter")

    def wait_acks(self, stream_handle):
        """Every peer has finished reading the previous round (trivially true in the first round: the counter is 0)."""
        self._wait(self.acks, stream_handle)

    def signal_ready(self, stream_handle):
        self._signal(self.ready, 1, stream_handle)

    def wait_ready(self, stream_handle):
        self._wait(self.ready, stream_handle)

    def signal_acks(self, stream_handle):
        self._signal(self.acks, 0, stream_handle)

    def close(self):
        self.ready.close()
        self.acks.close()
        self.counter.close()


class GatherBuffers(PeerBuffers):
    """The patch encoder's gather buffers: [4][world * per_rank][T][D] per rank."""

    def __init__(self, world: int, rank: int, per_rank: int, tokens: int, dim: int, precision: str):
        super().__init__(world, rank, (4, world * per_rank, tokens, dim), precision)


class ShardedPatchEncoder:
    """Runs a trunk-only engine (head "encoder_taps", batch = per_rank) on this rank's crops and leaves ALL crops'
    taps in ``gathered()``.  ``engine`` must have been built with batch == shard_bounds(n_items, world)[0]."""

    def __init__(self, engine, n_items: int, world: int, rank: int, mode: str = "fused", external_sync: bool = False):
        """`external_sync`: the caller orders the ranks itself (DepthProContext wraps every round in a `PeerSync`
        wait_acks / signal_ready / wait_ready / signal_acks hand-shake on the stream).  Otherwise the class does: `finish()`
        ends a round with a barrier, and the next fused `enqueue()` first waits -- stream drained, barrier -- until every rank
        has released the previous round (`release()`), so that no rank's stores can land in a buffer a peer still reads."""
        import torch
        if mode not in ("fused", "nccl"):
            raise ValueError(f"[MDET] unknown gather mode {mode!r}")
        self.engine, self.world, self.rank, self.mode, self.n_items = engine, world, rank, mode, n_items
        self.external_sync, self._round_open = bool(external_sync), False
        self.per_rank, self.bounds = shard_bounds(n_items, world)
        shape = engine.get_tensor_shape("output")              # [4, per_rank, T, D]
        if shape[1] != self.per_rank:
            raise ValueError(f"[MDET] engine batch {shape[1]} != shard size {self.per_rank}")
        self.tokens, self.dim = shape[2], shape[3]
        self.precision = "bf16" if engine.get_tensor_dtype("output") == np.dtype(np.uint16) else "fp16"
        self.ctx = engine.create_execution_context()
        self.buffers = GatherBuffers(world, rank, self.per_rank, self.tokens, self.dim, self.precision)
        self.local = None
        if mode == "fused":
            self.ctx.set_gather(world, rank, self.buffers.ptrs)
        else:
            dt = torch.bfloat16 if self.precision == "bf16" else torch.float16
            self.local = torch.empty(shape, dtype=dt, device="cuda")
            self.ctx.set_tensor_address("output", self.local.data_ptr())

    def enqueue(self, input_ptr: int, stream_handle: int) -> None:
        """Launch the trunk on this rank's crops (float32 [per_rank,3,384,384] at `input_ptr`); in "nccl" mode also the
        collective.  Asynchronous: call `finish()` before reading `gathered()`."""
        import torch
        import torch.distributed as dist
        if self.mode == "fused" and self.world > 1 and not self.external_sync and self._round_open:
            self.release()
        self.ctx.set_tensor_address("input", input_ptr)
        self.ctx.execute_async_v3(stream_handle)
        self._round_open = True
        if self.mode == "nccl" and self.world > 1:
            g = self.buffers.view()
            # per tap: the ranks' [per_rank, T, D] slabs are contiguous in the gathered layout.  The collective is ordered
            # on the caller's stream (behind the trunk, ahead of whatever the caller enqueues next).
            with torch.cuda.stream(torch_stream(stream_handle)):
                for i in range(4):
                    dist.all_gather_into_tensor(g[i], self.local[i])
        elif self.mode == "nccl":
            self.buffers.view().copy_(self.local)

    def finish(self) -> None:
        """Every rank's stores have landed in every buffer: drain the stream, then a barrier across the ranks."""
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()

    def release(self) -> None:
        """This rank is done reading `gathered()` of the current round (its consumers have been enqueued): drain them and meet
        the other ranks, after which anyone may overwrite the buffers.  Called by the next `enqueue()` if the caller did not."""
        import torch
        import torch.distributed as dist
        if self._round_open and self.world > 1 and not self.external_sync:
            torch.cuda.synchronize()
            dist.barrier()
        self._round_open = False

    def gathered(self):
        """[4, n_items, T, D]: the padded slots of the last rank(s) are cut off."""
        return self.buffers.view()[:, :self.n_items]

    def close(self):
        self.ctx.close()
        self.buffers.close()


def merge_features(taps, precision: str, stream_handle: int, hook_taps: Sequence[int] = (1, 0), final_tap: int = 3):
    """The gathered taps of the 35 crops ([4, 35, 576, D] on the device, `ShardedPatchEncoder.gathered()`) -> Depth Pro's
    five merged NHWC feature maps [f24, f48, f96, hook_a 96, hook_b 96] (the order transformers' DepthProPatchEncoder
    returns them; hooks (11, 5) of the released ViT-L trunk are taps (1, 0)).  One small gather kernel per map."""
    import ctypes as C
    import torch
    from . import _lib
    lib = _lib.load()
    D = int(taps.shape[-1])
    plan = [(final_tap, 34, 1, 0, 24), (final_tap, 25, 3, 6, 48), (final_tap, 0, 5, 3, 96)]
    plan += [(t, 0, 5, 3, 96) for t in hook_taps]
    out = []
    for tap, first, per_side, pad, side in plan:
        o = torch.empty(side, side, D, dtype=taps.dtype, device=taps.device)
        src = taps[tap, first:first + per_side * per_side]
        _lib.check(lib.mde_k_merge_patches(_lib.PRECISIONS[precision], C.c_void_p(src.data_ptr()), per_side, 24, pad, D,
                                           C.c_void_p(o.data_ptr()), C.c_void_p(int(stream_handle))), "mde_k_merge_patches")
        out.append(o)
    return out
