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

`VGGTEngine` is the whole exported model (models/vggt/onnx_export.py:38-52 `VGGTDepthOnlyWrapper`; spec.json: images float32
[1, S, 3, 518, 518] scaled by 1/255, output `depth`) behind the engine / context surface `allocate_buffers` / `do_inference`
drive: the DINOv2-with-registers trunk (a trunk-only C-ABI engine with four register tokens and the graph's ImageNet normalisation
in its input stage), camera / register tokens (`mde_k_assemble_tokens`), the aggregator above, and the depth head --
LayerNorm over the 2D-wide [frame | global] taps, DPT reassemble with the head's sin / cos position embedding added after each
projection and again after the final up-sampling, RefineNets, `exp` -- composed from the GEMM / conv kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import Mapping, Sequence

import numpy as np

from . import _lib, engine as E, sharding as S, weights as W
from .depth_pro import _Ops, _t, pack_conv3x3, pack_conv3x3_s2

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
                 taps: Sequence[int] = (), device: int = 0, causal: bool = False, cache_frames: int = 0):
        """causal: StreamVGGT's temporal causal attention -- in a global block the tokens of frame i see the tokens of frames
        <= i (oracle/vggt_torch.py `aggregate`).  cache_frames > 0 (with frames_total == 1): the streaming form of the same
        model -- `forward(tokens, stream, frame_index=t)` takes ONE frame, appends its keys / values to a per-block cache
        ([cache_frames * N, 2D] 16-bit, written by the qk-norm + RoPE kernel) and attends to the t + 1 frames held there; frame t
        of a stream then equals frame t of the causal forward over the whole sequence."""
        import torch
        if causal and world > 1:
            raise ValueError("[MDET] the causal (StreamVGGT) aggregator is single-GPU: a frame only needs its predecessors' keys")
        if cache_frames and (frames_total != 1 or not causal):
            raise ValueError("[MDET] the key / value cache belongs to the causal aggregator fed one frame at a time")
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
        self.causal, self.cache_frames, self.frame_index = bool(causal), int(cache_frames), 0
        self.cache = [z(self.cache_frames * self.N, 2 * D) for _ in range(depth)] if cache_frames else None
        self.kv = self.sync = None
        if world > 1:
            self.kv = S.PeerBuffers(world, rank, (self.rows_total, 2 * D), precision)
            self.mine = [p + rank * self.rows * 2 * D * 2 for p in self.kv.ptrs]              # our row range in every rank's buffer
            if gather == "fused":
                self.sync = S.PeerSync(world, rank)

    def _block(self, w: _BlockWeights, global_block: bool, sh: int, layer: int = 0) -> None:
        o, D, rows = self.ops, self.D, self.rows
        streaming = global_block and self.cache is not None
        t, N = self.frame_index, self.N
        o.layernorm(self.x, w.n1w, w.n1b, self.ln, rows, D, LN_EPS)
        o.gemm(self.ln, rows, D, D, w.qkv, 3 * D, o.ep(bias=w.qkv_b, out=self.qkv, ld_out=3 * D))
        sharded = global_block and self.world > 1
        if sharded and self.sync is not None:
            self.sync.wait_acks(sh)                                # every peer has finished reading the previous layer's K|V
        if streaming:       # this frame's finished K rows and its V rows go to rows [t * N, (t + 1) * N) of the block's cache
            gather = (self.cache[layer].data_ptr() + t * N * 2 * D * 2,)
        else:
            gather = self.mine if (sharded and self.sync is not None) else ()
        o.qknorm_rope(self.qkv, rows, self.heads, w.qw, w.qb, w.kw, w.kb, QK_EPS, self.pos, self.cos_sin, self.max_pos,
                      gather=gather, gather_ld=2 * D)
        if not global_block:
            o.attention(self.qkv, self.att, self.S_local, self.N, self.heads)
        elif streaming:
            o.attention_kv(self.qkv, 3 * D, self.cache[layer], 2 * D, 0, D, self.att, 1, N, (t + 1) * N, self.heads)
        elif self.causal and self.S_local > 1:
            # frame f's queries against the keys / values of frames 0..f, read in place from the packed q|k|v rows
            for f in range(self.S_local):
                o.attention_kv(self.qkv[f * N:], 3 * D, self.qkv[:, D:], 3 * D, 0, D, self.att[f * N:], 1, N, (f + 1) * N, self.heads)
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

    def forward(self, tokens_ptr: int, stream_handle, frame_index: int = None) -> None:
        """tokens: float32 [frames_local, N, D] on the device (special tokens first).  Asynchronous on `stream_handle`;
        afterwards `x` holds the last global block's output and `tap_out[layer]` the [frame | global] pair of each tap.
        frame_index (streaming aggregators only): the position of this frame in its stream, 0 starts a new one."""
        from .common_runtime import cuda_call, cudart
        sh = int(stream_handle)
        if self.cache is not None:
            if frame_index is None or not 0 <= int(frame_index) < self.cache_frames:
                raise ValueError(f"[MDET] streaming forward needs frame_index in [0, {self.cache_frames}), got {frame_index}")
            self.frame_index = int(frame_index)
        self.ops.stream = C.c_void_p(sh)
        self.ops.launches = 0
        cuda_call(cudart.cudaMemcpyAsync(self.x.data_ptr(), int(tokens_ptr), self.rows * self.D * 4,
                                         cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice, sh))
        for i, (fw, gw_) in enumerate(self.blocks):
            self._block(fw, False, sh, i)
            if i in self.tap_out:
                self._tap(i, 0, sh)
            self._block(gw_, True, sh, i)
            if i in self.tap_out:
                self._tap(i, 1, sh)

    def capture(self, tokens_ptr: int, stream_handle) -> None:
        """Record one forward (every launch, copy and flag hand-shake; ~430 nodes at 24 + 24 blocks) into a CUDA graph on
        `stream_handle`; `replay` then costs one launch on the host.  The hand-shake keeps its round number in device memory, so
        the captured arguments stay valid for every replay.  Not for gather="nccl" on more than one rank."""
        from .common_runtime import cuda_call, cudart
        if self.world > 1 and self.sync is None:
            raise RuntimeError("[MDET] the NCCL baseline is not captured; use gather='fused'")
        if self.cache is not None:
            raise RuntimeError("[MDET] a streaming step changes its cache offsets from frame to frame: not captured")
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


# ================================================================================================ the whole model
TRUNK = "aggregator.patch_embed."
RESNET_MEAN, RESNET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)      # reports/profile/vggt.json layer 2 (`Sub`, `Div`)
HEAD_LN_EPS = 1e-5
N_SPECIAL = 5


def head_pos_embed(channels: int, h: int, w: int, image_w: int, image_h: int):
    """float32 [h, w, channels]: what `DPTHead._apply_pos_embed` adds -- 0.1 * [sincos(u) | sincos(v)] of the h x w grid whose
    (u, v) span the image with its diagonal normalised to 1; sin / cos table as the reference exports it
    (core/export_compat.py:145-152: float32 frequencies 1 / 100 ** (i / (channels / 4)), sines first)."""
    import torch
    aspect = image_w / image_h
    diag = (aspect ** 2 + 1.0) ** 0.5
    sx, sy = aspect / diag, 1.0 / diag
    xs = torch.linspace(-sx * (w - 1) / w, sx * (w - 1) / w, steps=w, dtype=torch.float32)
    ys = torch.linspace(-sy * (h - 1) / h, sy * (h - 1) / h, steps=h, dtype=torch.float32)

    def sincos(pos, dim):
        omega = torch.arange(dim // 2, dtype=torch.float32)
        omega /= dim / 2.0
        omega = 1.0 / 100.0 ** omega
        out = torch.einsum("m,d->md", pos.to(torch.float32), omega)
        return torch.cat([torch.sin(out), torch.cos(out)], dim=1)

    eu, ev = sincos(xs, channels // 2), sincos(ys, channels // 2)                  # [w, C/2], [h, C/2]
    return (torch.cat([eu[None].expand(h, -1, -1), ev[:, None].expand(-1, w, -1)], dim=-1) * 0.1).contiguous()


def pack_deconv(w, s: int, dtype):
    """ConvTranspose2d(kernel == stride == s) [cin, cout, s, s] -> GEMM B [s*s*cout, cin], row (ky*s + kx)*cout + o."""
    cin, cout = w.shape[:2]
    return w.permute(2, 3, 1, 0).reshape(s * s * cout, cin).to(dtype).contiguous()


def required_tensors(encoder: str = "vitl", depth: int = 24, features: int = 256, out_channels: Sequence[int] = (256, 512, 1024, 1024),
                     trunk_grid: int = 37) -> dict:
    """Name -> shape of every tensor the engine reads from a VGGT / StreamVGGT state dict, under upstream's module tree (the model
    models/vggt/onnx_export.py:38-52 and models/streamvggt/onnx_export.py:35-53 load their checkpoints into): the DINOv2 trunk with
    four registers under `aggregator.patch_embed.`, the two-variant camera / register tokens, `depth` frame and global blocks with
    q / k LayerNorms, and the depth head (LayerNorm over 2D, projections, resize layers, layer_rn, RefineNets -- refinenet4
    without its first residual unit -- and the output convolutions with two channels: log depth and confidence).
    `export_checkpoint` checks a checkpoint against it before anything is written."""
    c = W.ENCODERS[encoder]
    D, L, F = c["embed_dim"], c["depth"], int(features)
    oc = [int(x) for x in out_channels]
    s = {}
    t = TRUNK
    s[t + "cls_token"] = (1, 1, D); s[t + "register_tokens"] = (1, 4, D); s[t + "pos_embed"] = (1, 1 + trunk_grid * trunk_grid, D)
    s[t + "patch_embed.proj.weight"] = (D, 3, 14, 14); s[t + "patch_embed.proj.bias"] = (D,)
    s[t + "norm.weight"] = (D,); s[t + "norm.bias"] = (D,)

    def block(b, qk_norm):
        for n in ("norm1", "norm2"):
            s[b + n + ".weight"] = (D,); s[b + n + ".bias"] = (D,)
        s[b + "attn.qkv.weight"] = (3 * D, D); s[b + "attn.qkv.bias"] = (3 * D,)
        s[b + "attn.proj.weight"] = (D, D); s[b + "attn.proj.bias"] = (D,)
        if qk_norm:
            for n in ("q_norm", "k_norm"):
                s[b + f"attn.{n}.weight"] = (64,); s[b + f"attn.{n}.bias"] = (64,)
        s[b + "ls1.gamma"] = (D,); s[b + "ls2.gamma"] = (D,)
        s[b + "mlp.fc1.weight"] = (4 * D, D); s[b + "mlp.fc1.bias"] = (4 * D,)
        s[b + "mlp.fc2.weight"] = (D, 4 * D); s[b + "mlp.fc2.bias"] = (D,)

    for i in range(L):
        block(f"{t}blocks.{i}.", False)
    s["aggregator.camera_token"] = (1, 2, 1, D); s["aggregator.register_token"] = (1, 2, 4, D)
    for i in range(int(depth)):
        block(f"aggregator.frame_blocks.{i}.", True)
        block(f"aggregator.global_blocks.{i}.", True)
    h = "depth_head."
    s[h + "norm.weight"] = (2 * D,); s[h + "norm.bias"] = (2 * D,)
    for i in range(4):
        s[h + f"projects.{i}.weight"] = (oc[i], 2 * D, 1, 1); s[h + f"projects.{i}.bias"] = (oc[i],)
        s[h + f"scratch.layer{i + 1}_rn.weight"] = (F, oc[i], 3, 3)
        r = h + f"scratch.refinenet{i + 1}."
        s[r + "out_conv.weight"] = (F, F, 1, 1); s[r + "out_conv.bias"] = (F,)
        for u in (("resConfUnit2",) if i == 3 else ("resConfUnit1", "resConfUnit2")):
            for cv in ("conv1", "conv2"):
                s[r + f"{u}.{cv}.weight"] = (F, F, 3, 3); s[r + f"{u}.{cv}.bias"] = (F,)
    s[h + "resize_layers.0.weight"] = (oc[0], oc[0], 4, 4); s[h + "resize_layers.0.bias"] = (oc[0],)
    s[h + "resize_layers.1.weight"] = (oc[1], oc[1], 2, 2); s[h + "resize_layers.1.bias"] = (oc[1],)
    s[h + "resize_layers.3.weight"] = (oc[3], oc[3], 3, 3); s[h + "resize_layers.3.bias"] = (oc[3],)
    s[h + "scratch.output_conv1.weight"] = (F // 2, F, 3, 3); s[h + "scratch.output_conv1.bias"] = (F // 2,)
    s[h + "scratch.output_conv2.0.weight"] = (32, F // 2, 3, 3); s[h + "scratch.output_conv2.0.bias"] = (32,)
    s[h + "scratch.output_conv2.2.weight"] = (2, 32, 1, 1); s[h + "scratch.output_conv2.2.bias"] = (2,)
    return s


def export_checkpoint(checkpoint_path: str, out_path: str, encoder: str = "vitl", depth: int = 24, features: int = 256,
                      out_channels: Sequence[int] = (256, 512, 1024, 1024), taps: Sequence[int] = (4, 11, 17, 23), frames: int = 1,
                      image_hw=(518, 518), family: str = "vggt", stream_frames: int = 0) -> dict:
    """Stage `export` for an upstream VGGT (`family="vggt"`, models/vggt/onnx_export.py:60-75) or StreamVGGT (`"streamvggt"`,
    models/streamvggt/onnx_export.py:70-78: `ckpt = torch.load(...); model.load_state_dict(ckpt)`) checkpoint, .pt / .pth or
    .safetensors: read the state dict, check it against `required_tensors` (the camera / point / track heads the reference's
    wrappers never run are dropped), write the .mdew file `common.get_engine` builds the engine from."""
    import torch
    if checkpoint_path.endswith(".safetensors"):
        from safetensors.torch import load_file
        sd = load_file(checkpoint_path, device="cpu")
    else:
        sd = torch.load(checkpoint_path, map_location="cpu", weights_only=True)
    if isinstance(sd, dict) and "state_dict" in sd and "aggregator.camera_token" not in sd:
        sd = sd["state_dict"]
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    grid = 37
    if TRUNK + "pos_embed" in sd:
        n = int(sd[TRUNK + "pos_embed"].shape[1]) - 1
        grid = int(round(n ** 0.5))
    need = required_tensors(encoder, depth, features, out_channels, trunk_grid=grid)
    missing = [k for k in need if k not in sd]
    wrong = [f"{k}: {tuple(sd[k].shape)} != {shape}" for k, shape in need.items() if k in sd and tuple(sd[k].shape) != shape]
    if missing or wrong:
        raise ValueError(f"[MDET] {checkpoint_path} is not a {family} ({encoder}, {depth} + {depth} blocks, {features} head features) "
                         f"checkpoint: {len(missing)} tensors missing (first: {missing[:3]}), {len(wrong)} with another shape (first: {wrong[:3]})")
    meta = W.describe_vggt(encoder, depth, features, out_channels, taps, frames, image_hw, family, stream_frames)
    meta["source_checkpoint_sha256"] = W.file_sha256(checkpoint_path)
    W.save(out_path, {k: sd[k] for k in need}, meta)
    return meta


class _HeadWeights:
    def __init__(self, sd: Mapping, D: int, F: int, oc, gh: int, gw: int, H: int, Wd: int, frames: int, dtype, device):
        import torch
        g = lambda k: _t(sd["depth_head." + k])
        dev = lambda t: t.contiguous().to(device)
        self.norm_w, self.norm_b = dev(g("norm.weight")), dev(g("norm.bias"))
        self.proj, self.proj_b, self.pe, self.rn = [], [], [], []
        for i in range(4):
            self.proj.append(dev(g(f"projects.{i}.weight").flatten(1).to(dtype)))
            self.proj_b.append(dev(g(f"projects.{i}.bias")))
            pe = head_pos_embed(oc[i], gh, gw, Wd, H).reshape(gh * gw, oc[i])
            self.pe.append(dev(pe.repeat(frames, 1).to(dtype)))             # added as a 16-bit residual in the projection's epilogue
            self.rn.append(dev(pack_conv3x3(g(f"scratch.layer{i + 1}_rn.weight"), dtype)))
        self.ct0, self.ct0_b = dev(pack_deconv(g("resize_layers.0.weight"), 4, dtype)), dev(g("resize_layers.0.bias"))
        self.ct1, self.ct1_b = dev(pack_deconv(g("resize_layers.1.weight"), 2, dtype)), dev(g("resize_layers.1.bias"))
        self.rs3, self.rs3_b = dev(pack_conv3x3_s2(g("resize_layers.3.weight"), dtype)), dev(g("resize_layers.3.bias"))
        self.ref = []
        for i in range(4):
            r = f"scratch.refinenet{i + 1}."
            d = {"out.w": dev(g(r + "out_conv.weight").flatten(1).to(dtype)), "out.b": dev(g(r + "out_conv.bias"))}
            for u in (("resConfUnit2",) if i == 3 else ("resConfUnit1", "resConfUnit2")):
                for cv in ("conv1", "conv2"):
                    d[f"{u}.{cv}.w"] = dev(pack_conv3x3(g(r + f"{u}.{cv}.weight"), dtype))
                    d[f"{u}.{cv}.b"] = dev(g(r + f"{u}.{cv}.bias"))
            self.ref.append(d)
        self.oc1, self.oc1_b = dev(pack_conv3x3(g("scratch.output_conv1.weight"), dtype)), dev(g("scratch.output_conv1.bias"))
        self.oc2, self.oc2_b = dev(pack_conv3x3(g("scratch.output_conv2.0.weight"), dtype)), dev(g("scratch.output_conv2.0.bias"))
        last_w, last_b = g("scratch.output_conv2.2.weight"), g("scratch.output_conv2.2.bias")
        self.head_w, self.head_b = dev(last_w[0].flatten()), float(last_b[0])    # channel 0 = log depth (channel 1, the confidence, is not exported)
        self.pe_out = dev(head_pos_embed(F // 2, H, Wd, Wd, H).reshape(H * Wd, F // 2))   # fp32, added by the up-sampling kernel


class VGGTEngine:
    """Stands in for the tensorrt.ICudaEngine models/vggt/onnx2trt.py builds: bindings `images` float32 [1, S, 3, H, W] (values
    0..1) and `depth` float32 [1, S, H, W, 1].  With `world` > 1 (one process per GPU) the S frames of the scene are sharded by
    rank: the bindings then hold this rank's `S / world` frames."""

    TensorIOMode = E.TensorIOMode
    IO = (("images", True), ("depth", False))

    def __init__(self, state_dict: Mapping, encoder: str = "vitl", depth: int = 24, features: int = 256,
                 out_channels: Sequence[int] = (256, 512, 1024, 1024), taps: Sequence[int] = (4, 11, 17, 23), frames: int = 16,
                 image_hw=(518, 518), precision: str = "fp16", world: int = 1, rank: int = 0, gather: str = "fused", device: int = 0,
                 causal: bool = False, stream_frames: int = 0):
        """causal=True: StreamVGGT (models/streamvggt/onnx_export.py:35-53: the same aggregator -> depth head pair with temporal
        causal attention in the global blocks; identical to VGGT at the one frame the reference exports).  stream_frames > 0
        (needs frames == 1): the streaming form -- every `execute_async_v3` takes the NEXT frame of a stream of up to
        `stream_frames` frames and attends to the cached keys / values of its predecessors; `context.reset_stream()` starts a
        new stream.  Frame t of a stream equals frame t of the causal forward over the sequence (tests/test_vggt_gpu.py)."""
        import torch
        if stream_frames and (int(frames) != 1 or world != 1):
            raise ValueError("[MDET] a streaming engine takes one frame per call on one GPU (frames=1, world=1)")
        causal = bool(causal or stream_frames)
        self.causal, self.stream_frames = causal, int(stream_frames)
        cfg = W.ENCODERS[encoder]
        H, Wd = int(image_hw[0]), int(image_hw[1])
        if H % 14 or Wd % 14:
            raise ValueError(f"[MDET] image size {H}x{Wd} is not a multiple of the patch size 14")
        if frames % world:
            raise ValueError(f"[MDET] {frames} frames do not divide over {world} ranks")
        if features % 64 or any(c % 8 for c in out_channels):
            raise ValueError("[MDET] decoder widths must be multiples of 64 (features) / 8 (out_channels)")
        self.encoder, self.depth, self.F, self.oc, self.taps = encoder, int(depth), int(features), [int(c) for c in out_channels], [int(t) for t in taps]
        self.D, self.heads, self.H, self.W, self.gh, self.gw = cfg["embed_dim"], cfg["num_heads"], H, Wd, H // 14, Wd // 14
        self.S_total, self.S, self.world, self.rank, self.precision = int(frames), int(frames) // int(world), int(world), int(rank), precision
        self.dtype = torch.bfloat16 if precision == "bf16" else torch.float16
        self.device = torch.device("cuda", device)
        self.trunk = None
        self.agg = None
        try:
            # ---- trunk: DINOv2 with four registers; the graph's (x - mean) / std (profile layer 2) applied where the patch rows are formed
            sd = {"pretrained." + k[len(TRUNK):]: _t(v) for k, v in state_dict.items() if k.startswith(TRUNK)}
            if not sd:
                raise ValueError(f"[MDET] the state dict holds no '{TRUNK}*' tensors")
            L = cfg["depth"]
            meta = W.describe(encoder, H, Wd, None)
            meta["taps"] = [L - 4, L - 3, L - 2, L - 1]
            meta["registers"] = int(sd["pretrained.register_tokens"].shape[1])
            self.trunk = E.Engine(E.make_desc(meta, precision=precision, batch=self.S, head="encoder_taps", tap_norm_mask=0x8, device=device,
                                                  normalise_f32=True, mean=RESNET_MEAN, std=RESNET_STD), meta)
            self.trunk.load_state_dict(sd)
            self.trunk.finalize()
            # ---- aggregator + head
            self.agg = Aggregator(state_dict, self.D, self.depth, self.heads, self.gh, self.gw, frames_total=self.S_total, precision=precision,
                                  world=world, rank=rank, gather=gather, taps=self.taps, device=device, causal=causal,
                                  cache_frames=self.stream_frames)
            self.special = torch.cat([_t(state_dict["aggregator.camera_token"])[0], _t(state_dict["aggregator.register_token"])[0]],
                                     dim=1).contiguous().to(self.device)                      # [2, 5, D]
            self.head = _HeadWeights(state_dict, self.D, self.F, self.oc, self.gh, self.gw, H, Wd, self.S, self.dtype, self.device)
        except Exception:
            self.close()
            raise

    @property
    def num_io_tensors(self) -> int:
        return len(self.IO)

    def get_tensor_name(self, i: int) -> str:
        return self.IO[i][0]

    def get_tensor_shape(self, name: str):
        return {"images": (1, self.S, 3, self.H, self.W), "depth": (1, self.S, self.H, self.W, 1)}[name]

    def get_tensor_profile_shape(self, name: str, profile_idx: int):
        s = self.get_tensor_shape(name)
        return (s, s, s)

    def get_tensor_dtype(self, name: str) -> np.dtype:
        self.get_tensor_shape(name)
        return np.dtype(np.float32)

    def get_tensor_mode(self, name: str):
        return E.TensorIOMode.INPUT if dict(self.IO)[name] else E.TensorIOMode.OUTPUT

    def create_execution_context(self) -> "VGGTContext":
        return VGGTContext(self)

    def close(self) -> None:
        if getattr(self, "agg", None) is not None:
            self.agg.close()
            self.agg = None
        if getattr(self, "trunk", None) is not None:
            self.trunk.close()
            self.trunk = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class VGGTContext:
    """`set_tensor_address` for the two bindings, `execute_async_v3(stream)`: trunk engine (its own CUDA graph), token assembly,
    aggregator (captured into a CUDA graph at the first call), depth head.  Every buffer is allocated here."""

    def __init__(self, engine: VGGTEngine):
        import torch
        e = self.e = engine
        self.ops = e.agg.ops
        self.addr = {}
        dev, dt = e.device, e.dtype
        S_, T, D, F, oc, gh, gw = e.S, e.gh * e.gw, e.D, e.F, e.oc, e.gh, e.gw
        z16 = lambda *shape: torch.zeros(*shape, dtype=dt, device=dev)
        self.trunk_ctx = e.trunk.create_execution_context()
        self.trunk_out = z16(4, S_, T, D)
        self.tokens = torch.zeros(S_, N_SPECIAL + T, D, dtype=torch.float32, device=dev)
        self.lvl = [(4 * gh, 4 * gw), (2 * gh, 2 * gw), (gh, gw), ((gh - 1) // 2 + 1, (gw - 1) // 2 + 1)]
        b = {"ln": z16(S_ * T, 2 * D)}
        for i in range(4):
            b[f"pr{i}"] = z16(S_ * T, oc[i])
            h, w = self.lvl[i]
            b[f"l{i}"] = b[f"pr{i}"] if i == 2 else z16(S_ * h * w, oc[i])
            for name in ("r", "r.relu", "a", "s", "s.relu", "a2", "u", "q"):
                b[f"{name}{i}"] = z16(S_ * h * w, F)
            ho, wo = (self.lvl[i - 1] if i > 0 else (2 * h, 2 * w))
            b[f"path{i}"] = z16(S_ * ho * wo, F)
        b["cols3"] = z16(S_ * self.lvl[3][0] * self.lvl[3][1], 9 * oc[3])
        h1, w1 = 2 * self.lvl[0][0], 2 * self.lvl[0][1]
        b["o1"] = z16(S_ * h1 * w1, F // 2)
        b["up"] = z16(S_ * e.H * e.W, F // 2)
        self.b = b
        self.launches_per_enqueue = 0
        self._captured = False
        self.stream_index = 0              # streaming engines: position of the next frame in its stream

    def reset_stream(self) -> None:
        """Streaming engines: the next `execute_async_v3` starts a new stream (its frame sees only itself)."""
        self.stream_index = 0

    def set_tensor_address(self, name: str, ptr: int) -> bool:
        self.e.get_tensor_shape(name)
        self.addr[name] = int(ptr)
        return True

    def set_input_shape(self, name: str, shape) -> bool:
        if tuple(int(s) for s in shape) != self.e.get_tensor_shape(name):
            raise ValueError(f"[MDET] {name}: engines are static, shape {tuple(shape)} != {self.e.get_tensor_shape(name)}")
        return True

    def get_buffer(self, name: str):
        return self.b[name]

    def execute_async_v3(self, stream_handle) -> bool:
        missing = [n for n, _ in self.e.IO if not self.addr.get(n)]
        if missing:
            raise RuntimeError(f"[MDET] execute before set_tensor_address for {missing}")
        e, o, b, hw = self.e, self.ops, self.b, self.e.head
        sh = int(stream_handle)
        S_, T, D, F, oc, gh, gw = e.S, e.gh * e.gw, e.D, e.F, e.oc, e.gh, e.gw
        # 1. trunk: normalised patch tokens of every frame (slice 3 of the trunk-only engine's output)
        self.trunk_ctx.set_tensor_address("input", self.addr["images"])
        self.trunk_ctx.set_tensor_address("output", self.trunk_out.data_ptr())
        self.trunk_ctx.execute_async_v3(sh)
        o.stream = C.c_void_p(sh)
        o.launches = 0
        # 2. camera + register tokens in front
        streaming = e.stream_frames > 0
        if streaming and self.stream_index >= e.stream_frames:
            raise RuntimeError(f"[MDET] the stream is {e.stream_frames} frames long (the size of the key / value cache): reset_stream() first")
        o.assemble_tokens(self.trunk_out[3], e.special, S_, T, N_SPECIAL, D, self.stream_index if streaming else e.rank * S_, self.tokens)
        # 3. aggregator (graph replay after the first call on a capturable stream)
        if streaming:
            e.agg.forward(self.tokens.data_ptr(), sh, frame_index=self.stream_index)
            self._agg_launches = o.launches
            self.stream_index += 1
        elif sh != 0 and (e.world == 1 or e.agg.sync is not None):
            if not self._captured:
                e.agg.forward(self.tokens.data_ptr(), sh)              # warm run: every kernel's attributes are set outside the capture
                self._agg_launches = o.launches
                e.agg.capture(self.tokens.data_ptr(), sh)
                self._captured = True
            else:
                e.agg.replay(sh)
        else:
            e.agg.forward(self.tokens.data_ptr(), sh)
            self._agg_launches = o.launches
        agg_launches = self._agg_launches
        o.stream = C.c_void_p(sh)
        o.launches = 0
        # 4. depth head
        N = N_SPECIAL + T
        for i, layer in enumerate(e.taps):
            o.layernorm(e.agg.tap_out[layer], hw.norm_w, hw.norm_b, b["ln"], S_ * N, 2 * D, HEAD_LN_EPS, drop=N_SPECIAL, ntok=N)
            o.gemm(b["ln"], S_ * T, 2 * D, 2 * D, hw.proj[i], oc[i], o.ep(bias=hw.proj_b[i], res1=hw.pe[i], out=b[f"pr{i}"], ld_out=oc[i]))
            if i == 0:
                o.gemm(b["pr0"], S_ * T, oc[0], oc[0], hw.ct0, 16 * oc[0], o.ep(bias=hw.ct0_b, out=b["l0"], ld_out=oc[0], shuffle=(4, oc[0], gh, gw)))
            elif i == 1:
                o.gemm(b["pr1"], S_ * T, oc[1], oc[1], hw.ct1, 4 * oc[1], o.ep(bias=hw.ct1_b, out=b["l1"], ld_out=oc[1], shuffle=(2, oc[1], gh, gw)))
            elif i == 3:
                o.im2col_s2(b["pr3"], gh, gw, oc[3], b["cols3"], batch=S_)
                h4, w4 = self.lvl[3]
                o.gemm(b["cols3"], S_ * h4 * w4, 9 * oc[3], 9 * oc[3], hw.rs3, oc[3], o.ep(bias=hw.rs3_b, out=b["l3"], ld_out=oc[3]))
        for i in range(4):
            h, w = self.lvl[i]
            o.conv3x3(b[f"l{i}"], h, w, oc[i], hw.rn[i], F, o.ep(out=b[f"r{i}"], out_relu=b[f"r.relu{i}"], ld_out=F), batch=S_)
        path = None
        for i in (3, 2, 1, 0):
            h, w = self.lvl[i]
            rf = hw.ref[i]
            s_relu, s_raw = b[f"r.relu{i}"], b[f"r{i}"]
            if i != 3:      # s = path + RCU1(r_i)
                o.conv3x3(b[f"r.relu{i}"], h, w, F, rf["resConfUnit1.conv1.w"], F, o.ep(bias=rf["resConfUnit1.conv1.b"], act=2, out=b[f"a{i}"], ld_out=F), batch=S_)
                o.conv3x3(b[f"a{i}"], h, w, F, rf["resConfUnit1.conv2.w"], F,
                          o.ep(bias=rf["resConfUnit1.conv2.b"], res1=b[f"r{i}"], res2=path, out=b[f"s{i}"], out_relu=b[f"s.relu{i}"], ld_out=F), batch=S_)
                s_relu, s_raw = b[f"s.relu{i}"], b[f"s{i}"]
            o.conv3x3(s_relu, h, w, F, rf["resConfUnit2.conv1.w"], F, o.ep(bias=rf["resConfUnit2.conv1.b"], act=2, out=b[f"a2{i}"], ld_out=F), batch=S_)
            o.conv3x3(b[f"a2{i}"], h, w, F, rf["resConfUnit2.conv2.w"], F, o.ep(bias=rf["resConfUnit2.conv2.b"], res1=s_raw, out=b[f"u{i}"], ld_out=F), batch=S_)
            # 1x1 out_conv before the bilinear up-sampling (both linear, interpolation weights sum to 1: they commute)
            o.gemm(b[f"u{i}"], S_ * h * w, F, F, rf["out.w"], F, o.ep(bias=rf["out.b"], out=b[f"q{i}"], ld_out=F))
            ho, wo = (self.lvl[i - 1] if i > 0 else (2 * h, 2 * w))
            o.bilinear(b[f"q{i}"], b[f"path{i}"], S_, h, w, ho, wo, F)
            path = b[f"path{i}"]
        h1, w1 = 2 * self.lvl[0][0], 2 * self.lvl[0][1]
        o.conv3x3(path, h1, w1, F, hw.oc1, F // 2, o.ep(bias=hw.oc1_b, out=b["o1"], ld_out=F // 2), batch=S_)
        o.bilinear(b["o1"], b["up"], S_, h1, w1, e.H, e.W, F // 2, addend=hw.pe_out)
        o.conv3x3(b["up"], e.H, e.W, F // 2, hw.oc2, 32,
                  o.ep(bias=hw.oc2_b, ld_out=32, head_w=hw.head_w, head_b=hw.head_b, head_out=self.addr["depth"], head_act=1), batch=S_)
        self.launches_per_enqueue = self.trunk_ctx.launches_per_enqueue + 1 + agg_launches + o.launches
        return True

    def close(self) -> None:
        if getattr(self, "trunk_ctx", None) is not None:
            self.trunk_ctx.close()
            self.trunk_ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
