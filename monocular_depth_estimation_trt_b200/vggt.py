"""VGGT's aggregator on the B200 kernels (BASELINE.json configs[4]; SURVEY section 8 e row 3): the alternating frame /
global attention blocks of the model models/vggt/onnx_export.py:38-52 exports, with the frames of a scene sharded over one
process per GPU.

  frame block    every frame attends to its own N = 5 + gh*gw tokens: local to the rank that owns the frame
  global block   every token attends to ALL frames' tokens: queries stay local, keys / values are all-gathered.  The
                 exchange is fused into the kernel that finishes K (per-head LayerNorm + 2-D rotary embedding,
                 `mde_k_qknorm_rope`): it stores the K and V rows straight into every rank's gathered buffer (peer memory
                 over NVLink), a flag hand-shake on the stream (`sharding.PeerSync`) orders the ranks, and the attention
                 kernel reads queries from the local rows and keys / values from the gathered buffer.  No collective launch,
                 no host synchronisation.  `gather="nccl"` is the baseline (all_gather_into_tensor on the same stream).

A block is eight launches: LayerNorm, QKV GEMM, qk-norm + RoPE, attention, projection GEMM (LayerScale + residual in the
epilogue), LayerNorm, FC1 GEMM (+GELU), FC2 GEMM (LayerScale + residual).  torch provides device memory only.
The DINOv2-with-registers trunk in front of the aggregator and the DPT head behind it are not built.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Sequence

import numpy as np

from . import _lib, sharding as S
from .depth_pro import _Ops, _t

LN_EPS = 1e-5          # aggregator blocks: nn.LayerNorm's default (the DINOv2 trunk in front uses 1e-6)
QK_EPS = 1e-5
ROPE_FREQUENCY = 100.0


def token_positions(gh: int, gw: int, n_special: int = 5) -> np.ndarray:
    """int32 [n_special + gh*gw, 2]: (0, 0) for the camera / register tokens, (y + 1, x + 1) for patch (y, x), row-major
    (core/export_compat.py:84-93 builds the patch grid; the aggregator shifts it by one and prepends the zeros)."""
    yy, xx = np.meshgrid(np.arange(gh), np.arange(gw), indexing="ij")
    grid = np.stack([yy.reshape(-1), xx.reshape(-1)], axis=1) + 1
    return np.concatenate([np.zeros((n_special, 2), dtype=np.int64), grid]).astype(np.int32)


def cos_sin_table(max_pos: int, frequency: float = ROPE_FREQUENCY):
    """float32 [max_pos, 32]: cos(p * f_j), j < 16, then sin(p * f_j), with f_j = frequency ** -(2j / 32) evaluated in fp32
    the way upstream's rope.py does (torch ops, so the table is bit-identical to the one the oracle indexes)."""
    import torch
    inv_freq = 1.0 / (frequency ** (torch.arange(0, 32, 2).float() / 32))
    ang = torch.einsum("i,j->ij", torch.arange(max_pos).float(), inv_freq)
    return torch.cat([ang.cos(), ang.sin()], dim=1).contiguous()


class _BlockWeights:
    def __init__(self, sd: Mapping, prefix: str, dtype, device):
        g = lambda k: _t(sd[prefix + k])
        w16 = lambda k: g(k).to(dtype).contiguous().to(device)
        f32 = lambda k: g(k).contiguous().to(device)
        self.n1w, self.n1b, self.n2w, self.n2b = f32("norm1.weight"), f32("norm1.bias"), f32("norm2.weight"), f32("norm2.bias")
        self.qkv, self.qkv_b = w16("attn.qkv.weight"), f32("attn.qkv.bias")
        self.qw, self.qb = f32("attn.q_norm.weight"), f32("attn.q_norm.bias")
        self.kw, self.kb = f32("attn.k_norm.weight"), f32("attn.k_norm.bias")
        self.proj, self.proj_b, self.ls1 = w16("attn.proj.weight"), f32("attn.proj.bias"), f32("ls1.gamma")
        self.fc1, self.fc1_b = w16("mlp.fc1.weight"), f32("mlp.fc1.bias")
        self.fc2, self.fc2_b, self.ls2 = w16("mlp.fc2.weight"), f32("mlp.fc2.bias"), f32("ls2.gamma")


class Aggregator:
    """`depth` x (frame block, global block) over this rank's `frames_local` frames of one scene."""

    def __init__(self, state_dict: Mapping, dim: int, depth: int, num_heads: int, gh: int, gw: int, frames_total: int,
                 precision: str = "bf16", n_special: int = 5, world: int = 1, rank: int = 0, gather: str = "fused",
                 taps: Sequence[int] = (), device: int = 0):
        import torch
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"[MDET] precision {precision!r} is not supported; use one of {sorted(_lib.PRECISIONS)}")
        if dim != num_heads * 64:
            raise ValueError("[MDET] the attention kernels need a head dimension of 64")
        if frames_total % world:
            raise ValueError(f"[MDET] {frames_total} frames do not divide over {world} ranks")
        if gather not in ("fused", "nccl"):
            raise ValueError(f"[MDET] unknown gather mode {gather!r}")
        self.D, self.depth, self.heads, self.world, self.rank, self.gather = dim, depth, num_heads, world, rank, gather
        self.N = n_special + gh * gw
        self.S_total, self.S_local = frames_total, frames_total // world
        self.rows, self.rows_total = self.S_local * self.N, frames_total * self.N
        self.precision, self.taps = precision, [int(t) for t in taps]
        self.dtype = torch.bfloat16 if precision == "bf16" else torch.float16
        dev = self.device = torch.device("cuda", device)
        self.blocks = [(_BlockWeights(state_dict, f"aggregator.frame_blocks.{i}.", self.dtype, dev),
                        _BlockWeights(state_dict, f"aggregator.global_blocks.{i}.", self.dtype, dev)) for i in range(depth)]
        self.ops = _Ops(precision)
        pos = token_positions(gh, gw, n_special)
        self.pos = torch.from_numpy(np.tile(pos, (self.S_local, 1))).to(dev)
        self.max_pos = int(pos.max()) + 1
        self.cos_sin = cos_sin_table(self.max_pos).to(dev)
        z = lambda *shape, dt=self.dtype: torch.zeros(*shape, dtype=dt, device=dev)
        D = dim
        self.x = z(self.rows, D, dt=torch.float32)
        self.ln, self.qkv, self.att, self.hid = z(self.rows, D), z(self.rows, 3 * D), z(self.rows, D), z(self.rows, 4 * D)
        self.tap_out = {t: z(self.rows, 2 * D, dt=torch.float32) for t in self.taps}      # [frame | global] residual streams
        self.kv = self.sync = None
        if world > 1:
            self.kv = S.PeerBuffers(world, rank, (self.rows_total, 2 * D), precision)
            self.mine = [p + rank * self.rows * 2 * D * 2 for p in self.kv.ptrs]              # our row range in every rank's buffer
            if gather == "fused":
                self.sync = S.PeerSync(world, rank)

    def _block(self, w: _BlockWeights, global_block: bool, sh: int) -> None:
        o, D, rows = self.ops, self.D, self.rows
        o.layernorm(self.x, w.n1w, w.n1b, self.ln, rows, D, LN_EPS)
        o.gemm(self.ln, rows, D, D, w.qkv, 3 * D, o.ep(bias=w.qkv_b, out=self.qkv, ld_out=3 * D))
        sharded = global_block and self.world > 1
        if sharded and self.sync is not None:
            self.sync.wait_acks(sh)                                # every peer has finished reading the previous layer's K|V
        o.qknorm_rope(self.qkv, rows, self.heads, w.qw, w.qb, w.kw, w.kb, QK_EPS, self.pos, self.cos_sin, self.max_pos,
                      gather=self.mine if (sharded and self.sync is not None) else (), gather_ld=2 * D)
        if not global_block:
            o.attention(self.qkv, self.att, self.S_local, self.N, self.heads)
        elif not sharded:
            o.attention(self.qkv, self.att, 1, rows, self.heads)
        else:
            if self.sync is not None:
                self.sync.signal_ready(sh)
                self.sync.wait_ready(sh)
            else:
                import torch
                import torch.distributed as dist
                with torch.cuda.stream(S.torch_stream(sh)):
                    dist.all_gather_into_tensor(self.kv.view(), self.qkv[:, D:].contiguous())
            o.attention_kv(self.qkv, 3 * D, self.kv.own, 2 * D, 0, D, self.att, 1, rows, self.rows_total, self.heads)
            if self.sync is not None:
                self.sync.signal_acks(sh)
        o.gemm(self.att, rows, D, D, w.proj, D, o.ep(bias=w.proj_b, gamma=w.ls1, x=self.x, accumulate_x=True, ld_out=D))
        o.layernorm(self.x, w.n2w, w.n2b, self.ln, rows, D, LN_EPS)
        o.gemm(self.ln, rows, D, D, w.fc1, 4 * D, o.ep(bias=w.fc1_b, act=1, out=self.hid, ld_out=4 * D))
        o.gemm(self.hid, rows, 4 * D, 4 * D, w.fc2, D, o.ep(bias=w.fc2_b, gamma=w.ls2, x=self.x, accumulate_x=True, ld_out=D))

    def _tap(self, layer: int, half: int, sh: int) -> None:
        from .common_runtime import cuda_call, cudart
        dst = self.tap_out[layer]
        cuda_call(cudart.cudaMemcpy2DAsync(dst.data_ptr() + half * self.D * 4, 2 * self.D * 4, self.x.data_ptr(), self.D * 4,
                                           self.D * 4, self.rows, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice, sh))

    def forward(self, tokens_ptr: int, stream_handle) -> None:
        """tokens: float32 [frames_local, N, D] on the device (special tokens first).  Asynchronous on `stream_handle`;
        afterwards `x` holds the last global block's output and `tap_out[layer]` the [frame | global] pair of each tap."""
        from .common_runtime import cuda_call, cudart
        sh = int(stream_handle)
        self.ops.stream = C.c_void_p(sh)
        self.ops.launches = 0
        cuda_call(cudart.cudaMemcpyAsync(self.x.data_ptr(), int(tokens_ptr), self.rows * self.D * 4,
                                         cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice, sh))
        for i, (fw, gw_) in enumerate(self.blocks):
            self._block(fw, False, sh)
            if i in self.tap_out:
                self._tap(i, 0, sh)
            self._block(gw_, True, sh)
            if i in self.tap_out:
                self._tap(i, 1, sh)

    def capture(self, tokens_ptr: int, stream_handle) -> None:
        """Record one forward (every launch, copy and flag hand-shake; ~430 nodes at 24 + 24 blocks) into a CUDA graph on
        `stream_handle`; `replay` then costs one launch on the host.  The hand-shake keeps its round number in device memory, so
        the captured arguments stay valid for every replay.  Not for gather="nccl" on more than one rank."""
        from .common_runtime import cuda_call, cudart
        if self.world > 1 and self.sync is None:
            raise RuntimeError("[MDET] the NCCL baseline is not captured; use gather='fused'")
        sh = int(stream_handle)
        if sh == 0:
            raise ValueError("[MDET] capture needs a non-default stream")
        self.release_graph()
        cuda_call(cudart.cudaStreamBeginCapture(sh, cudart.cudaStreamCaptureMode.cudaStreamCaptureModeThreadLocal))
        try:
            self.forward(tokens_ptr, sh)
        finally:
            graph = cuda_call(cudart.cudaStreamEndCapture(sh))
        self._graph_exec = cuda_call(cudart.cudaGraphInstantiate(graph, 0))
        cuda_call(cudart.cudaGraphDestroy(graph))

    def replay(self, stream_handle) -> None:
        from .common_runtime import cuda_call, cudart
        if getattr(self, "_graph_exec", None) is None:
            raise RuntimeError("[MDET] replay before capture")
        cuda_call(cudart.cudaGraphLaunch(self._graph_exec, int(stream_handle)))

    def release_graph(self) -> None:
        from .common_runtime import cudart
        if getattr(self, "_graph_exec", None) is not None:
            cudart.cudaGraphExecDestroy(self._graph_exec)
        self._graph_exec = None

    def close(self) -> None:
        self.release_graph()
        if self.sync is not None:
            self.sync.close()
            self.sync = None
        if self.kv is not None:
            self.kv.close()
            self.kv = None
