"""ORACLE (test infrastructure, never shipped, never on the product path).

VGGT's aggregator blocks, restated: the reference exports `VGGT` from the un-vendored, un-pinned
github.com/facebookresearch/vggt package (models/vggt/onnx_export.py:21,38-52: `aggregator(images)` -> `depth_head`).
Its aggregator alternates "frame" attention (every frame's tokens on their own) and "global" attention (all frames'
tokens of a scene as one sequence) over the same kind of block: a DINOv2-style pre-norm block whose attention
normalises q and k per head (`qk_norm`: LayerNorm over the 64 head features) and rotates them with a 2-D rotary
position embedding (frequency 100; the first half of the head features turns with the token's row, the second half
with its column; special tokens sit at position 0, patch (y, x) at (y + 1, x + 1)).

PARITY UNPINNED: no implementation of VGGT is importable in this container (transformers 5.5 has none) and the
reference holds no golden vector for it.  What anchors this restatement inside the reference: the position table
the reference itself re-implements for export (core/export_compat.py:84-93 `patched_call`: rows of (y, x) over the
patch grid), the alternating frame/global structure and the 1374 = 1 + 4 + 1369 tokens per frame visible in
reports/profile/vggt.json, and the wrapper at models/vggt/onnx_export.py:38-52.  The attention arithmetic itself is
cross-checked against torch's scaled_dot_product_attention, the rotary embedding against a complex-number form
(tests/test_oracle_vggt.py).
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F

LN_EPS = 1e-6          # block norms (DINOv2-style blocks: partial(nn.LayerNorm, eps=1e-6))
QK_EPS = 1e-5          # q_norm / k_norm: nn.LayerNorm(head_dim) with its default eps
ROPE_FREQUENCY = 100.0
N_SPECIAL = 5          # camera token + 4 register tokens ahead of the patch tokens


def positions(gh: int, gw: int, n_special: int = N_SPECIAL) -> torch.Tensor:
    """int64 [n_special + gh*gw, 2]: (0, 0) for the special tokens, (y + 1, x + 1) for patch (y, x) in row-major order
    (core/export_compat.py:84-93 builds the patch part; the aggregator shifts it by one and prepends zeros)."""
    yy = torch.arange(gh).unsqueeze(1).expand(-1, gw)
    xx = torch.arange(gw).unsqueeze(0).expand(gh, -1)
    grid = torch.stack([yy.reshape(-1), xx.reshape(-1)], dim=1) + 1
    return torch.cat([torch.zeros(n_special, 2, dtype=grid.dtype), grid], dim=0)


def rope_tables(max_pos: int, feature_dim: int = 32, frequency: float = ROPE_FREQUENCY):
    """cos / sin [max_pos, feature_dim]: angle(p, i) = p / frequency ** (2 * (i % (feature_dim/2)) / feature_dim)."""
    exponents = torch.arange(0, feature_dim, 2).float() / feature_dim
    inv_freq = 1.0 / (frequency ** exponents)
    angles = torch.einsum("i,j->ij", torch.arange(max_pos).float(), inv_freq)
    angles = torch.cat((angles, angles), dim=-1)
    return angles.cos(), angles.sin()


def _rotate(x: torch.Tensor) -> torch.Tensor:
    d = x.shape[-1]
    x1, x2 = x[..., : d // 2], x[..., d // 2:]
    return torch.cat((-x2, x1), dim=-1)


def rope_2d(t: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
    """t [B, heads, N, 64], pos [N, 2] (or [B, N, 2]) -> rotated: features [0, 32) with pos[:, 0], [32, 64) with pos[:, 1]."""
    if pos.dim() == 2:
        pos = pos[None].expand(t.shape[0], -1, -1)
    half = t.shape[-1] // 2
    cos, sin = rope_tables(int(pos.max()) + 1, half)
    out = []
    for part, p in ((t[..., :half], pos[..., 0]), (t[..., half:], pos[..., 1])):
        c, s = F.embedding(p, cos)[:, None], F.embedding(p, sin)[:, None]
        out.append(part * c + _rotate(part) * s)
    return torch.cat(out, dim=-1)


def block_param_shapes(dim: int, prefix: str) -> Dict[str, tuple]:
    hd = 64
    s = {}
    for n in ("norm1", "norm2"):
        s[prefix + n + ".weight"] = (dim,); s[prefix + n + ".bias"] = (dim,)
    s[prefix + "attn.qkv.weight"] = (3 * dim, dim); s[prefix + "attn.qkv.bias"] = (3 * dim,)
    for n in ("q_norm", "k_norm"):
        s[prefix + f"attn.{n}.weight"] = (hd,); s[prefix + f"attn.{n}.bias"] = (hd,)
    s[prefix + "attn.proj.weight"] = (dim, dim); s[prefix + "attn.proj.bias"] = (dim,)
    s[prefix + "ls1.gamma"] = (dim,); s[prefix + "ls2.gamma"] = (dim,)
    s[prefix + "mlp.fc1.weight"] = (4 * dim, dim); s[prefix + "mlp.fc1.bias"] = (4 * dim,)
    s[prefix + "mlp.fc2.weight"] = (dim, 4 * dim); s[prefix + "mlp.fc2.bias"] = (dim,)
    return s


def init_aggregator(dim: int, depth: int, seed: int = 0) -> Dict[str, torch.Tensor]:
    """`depth` frame blocks and `depth` global blocks under upstream's names (aggregator.frame_blocks.i / global_blocks.i),
    seeded non-degenerate init (the recipe of oracle/dav2_torch.py; LayerScale 0.5 so that every block matters)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for kind in ("frame_blocks", "global_blocks"):
        for i in range(depth):
            for k, shp in block_param_shapes(dim, f"aggregator.{kind}.{i}.").items():
                if k.endswith("gamma"):
                    v = torch.full(shp, 0.5)
                elif "norm" in k and k.endswith("weight"):
                    v = 1.0 + 0.1 * torch.randn(shp, generator=g)
                elif k.endswith("bias"):
                    v = 0.1 * torch.randn(shp, generator=g)
                else:
                    v = torch.randn(shp, generator=g) / math.sqrt(shp[-1])
                sd[k] = v.float().contiguous()
    return sd


def block(sd, prefix: str, t: torch.Tensor, pos: torch.Tensor, num_heads: int) -> torch.Tensor:
    """t [B, N, D], pos [N, 2] -> the block's output (norm1 -> attention with qk-norm + RoPE -> LayerScale -> residual;
    norm2 -> MLP (exact-erf GELU) -> LayerScale -> residual)."""
    B, N, D = t.shape
    hd = D // num_heads
    y = F.layer_norm(t, (D,), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], LN_EPS)
    qkv = F.linear(y, sd[prefix + "attn.qkv.weight"], sd[prefix + "attn.qkv.bias"]).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = F.layer_norm(q, (hd,), sd[prefix + "attn.q_norm.weight"], sd[prefix + "attn.q_norm.bias"], QK_EPS)
    k = F.layer_norm(k, (hd,), sd[prefix + "attn.k_norm.weight"], sd[prefix + "attn.k_norm.bias"], QK_EPS)
    q, k = rope_2d(q, pos), rope_2d(k, pos)
    a = torch.softmax((q * hd ** -0.5) @ k.transpose(-2, -1), dim=-1) @ v
    a = a.transpose(1, 2).reshape(B, N, D)
    t = t + sd[prefix + "ls1.gamma"] * F.linear(a, sd[prefix + "attn.proj.weight"], sd[prefix + "attn.proj.bias"])
    y = F.layer_norm(t, (D,), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], LN_EPS)
    y = F.gelu(F.linear(y, sd[prefix + "mlp.fc1.weight"], sd[prefix + "mlp.fc1.bias"]))
    return t + sd[prefix + "ls2.gamma"] * F.linear(y, sd[prefix + "mlp.fc2.weight"], sd[prefix + "mlp.fc2.bias"])


@torch.no_grad()
def aggregate(sd, tokens: torch.Tensor, gh: int, gw: int, num_heads: int, depth: int) -> List[torch.Tensor]:
    """tokens [S, N, D] (one scene: S frames of N = 5 + gh*gw tokens, special tokens first) -> per layer the concatenation
    [S, N, 2D] of the frame block's and the global block's output, as the aggregator hands them to the heads."""
    S, N, D = tokens.shape
    pos = positions(gh, gw, N - gh * gw)
    out = []
    t = tokens
    for i in range(depth):
        t = block(sd, f"aggregator.frame_blocks.{i}.", t, pos, num_heads)                          # S sequences of N tokens
        frame = t
        t = block(sd, f"aggregator.global_blocks.{i}.", t.reshape(1, S * N, D), pos.repeat(S, 1), num_heads).reshape(S, N, D)
        out.append(torch.cat([frame, t], dim=-1))
    return out
