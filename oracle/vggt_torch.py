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

LN_EPS = 1e-5          # aggregator block norms: upstream's Aggregator builds its Blocks with the default norm_layer = nn.LayerNorm
                       # (eps 1e-5); only the DINOv2 trunk in front (vit_large: partial(nn.LayerNorm, eps=1e-6)) uses 1e-6
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


def block(sd, prefix: str, t: torch.Tensor, pos: torch.Tensor, num_heads: int, mask: torch.Tensor = None) -> torch.Tensor:
    """t [B, N, D], pos [N, 2] -> the block's output (norm1 -> attention with qk-norm + RoPE -> LayerScale -> residual;
    norm2 -> MLP (exact-erf GELU) -> LayerScale -> residual).  mask: bool [N, N], True where a query may see a key (None: all)."""
    B, N, D = t.shape
    hd = D // num_heads
    y = F.layer_norm(t, (D,), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"], LN_EPS)
    qkv = F.linear(y, sd[prefix + "attn.qkv.weight"], sd[prefix + "attn.qkv.bias"]).reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = F.layer_norm(q, (hd,), sd[prefix + "attn.q_norm.weight"], sd[prefix + "attn.q_norm.bias"], QK_EPS)
    k = F.layer_norm(k, (hd,), sd[prefix + "attn.k_norm.weight"], sd[prefix + "attn.k_norm.bias"], QK_EPS)
    q, k = rope_2d(q, pos), rope_2d(k, pos)
    scores = (q * hd ** -0.5) @ k.transpose(-2, -1)
    if mask is not None:
        scores = scores.masked_fill(~mask, float("-inf"))
    a = torch.softmax(scores, dim=-1) @ v
    a = a.transpose(1, 2).reshape(B, N, D)
    t = t + sd[prefix + "ls1.gamma"] * F.linear(a, sd[prefix + "attn.proj.weight"], sd[prefix + "attn.proj.bias"])
    y = F.layer_norm(t, (D,), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"], LN_EPS)
    y = F.gelu(F.linear(y, sd[prefix + "mlp.fc1.weight"], sd[prefix + "mlp.fc1.bias"]))
    return t + sd[prefix + "ls2.gamma"] * F.linear(y, sd[prefix + "mlp.fc2.weight"], sd[prefix + "mlp.fc2.bias"])


@torch.no_grad()
def aggregate(sd, tokens: torch.Tensor, gh: int, gw: int, num_heads: int, depth: int, causal: bool = False) -> List[torch.Tensor]:
    """tokens [S, N, D] (one scene: S frames of N = 5 + gh*gw tokens, special tokens first) -> per layer the concatenation
    [S, N, 2D] of the frame block's and the global block's output, as the aggregator hands them to the heads.
    causal: StreamVGGT's temporal causal attention (models/streamvggt/onnx_export.py:35-53 exports the same aggregator / depth
    head pair as VGGT's; upstream streamvggt/models/aggregator.py masks the global blocks so that the tokens of frame i see the
    tokens of frames <= i, which is what lets a stream be processed frame by frame against cached keys / values).  With one
    frame -- all the reference ever exports -- the two coincide."""
    S, N, D = tokens.shape
    pos = positions(gh, gw, N - gh * gw)
    mask = None
    if causal and S > 1:
        frame_of = torch.arange(S).repeat_interleave(N)
        mask = frame_of[:, None] >= frame_of[None, :]
    out = []
    t = tokens
    for i in range(depth):
        t = block(sd, f"aggregator.frame_blocks.{i}.", t, pos, num_heads)                          # S sequences of N tokens
        frame = t
        t = block(sd, f"aggregator.global_blocks.{i}.", t.reshape(1, S * N, D), pos.repeat(S, 1), num_heads, mask).reshape(S, N, D)
        out.append(torch.cat([frame, t], dim=-1))
    return out


# ================================================================================================ the whole model
# VGGT as the reference exports it (models/vggt/onnx_export.py:38-52 `VGGTDepthOnlyWrapper`: aggregator -> depth_head, the
# depth map only).  Restated from the published facebookresearch/vggt sources (vggt/models/aggregator.py, vggt/heads/dpt_head.py,
# vggt/heads/utils.py, vggt/heads/head_act.py) and anchored inside the reference on
#   * reports/profile/vggt.json `engine_layers`: the input normalisation (layer 2 `Sub` / `Div`), the DINOv2 trunk under
#     `/aggregator/patch_embed/` (layers 4-...), 1374 tokens per frame, the head's order -- `projects.i/Conv + Add_*` (the
#     position embedding added right after each projection, layers 668-671), the four resize layers, `layer*_rn`, the
#     refinenets (refinenet4 without its first residual unit), `output_conv1`, `Resize`, `Add_15` (position embedding again),
#     `output_conv2.0 + Relu`, `output_conv2.2`, `Exp` (layers 672-711);
#   * core/export_compat.py:145-152: the sin / cos embedding of the head's position grid, as the engine computes it (float32);
#   * models/vggt/spec.json: images float32 [1, S, 3, 518, 518] scaled by 1/255 only; output `depth`.
# The DINOv2-with-registers trunk is pinned against transformers' Dinov2WithRegistersModel (tests/test_oracle_vggt.py).
# PARITY UNPINNED for the model as a whole (no VGGT implementation is importable here).
TRUNK = "aggregator.patch_embed."
RESNET_MEAN, RESNET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
HEAD_LN_EPS = 1e-5     # depth_head.norm = nn.LayerNorm(2 * embed_dim), default eps
POS_RATIO = 0.1


def make_sincos_pos_embed(embed_dim: int, pos: torch.Tensor, omega_0: float = 100.0) -> torch.Tensor:
    """core/export_compat.py:145-152 (the float32 form the reference exports with; upstream builds omega in float64)."""
    omega = torch.arange(embed_dim // 2, dtype=torch.float32)
    omega /= embed_dim / 2.0
    omega = 1.0 / omega_0 ** omega
    out = torch.einsum("m,d->md", pos.reshape(-1).to(torch.float32), omega)
    return torch.cat([torch.sin(out), torch.cos(out)], dim=1).float()


def create_uv_grid(width: int, height: int, aspect_ratio: float) -> torch.Tensor:
    """vggt/heads/utils.py `create_uv_grid`: [height, width, 2] of (u, v) spanning the image with the diagonal normalised to 1."""
    diag = (aspect_ratio ** 2 + 1.0) ** 0.5
    span_x, span_y = aspect_ratio / diag, 1.0 / diag
    xs = torch.linspace(-span_x * (width - 1) / width, span_x * (width - 1) / width, steps=width, dtype=torch.float32)
    ys = torch.linspace(-span_y * (height - 1) / height, span_y * (height - 1) / height, steps=height, dtype=torch.float32)
    uu, vv = torch.meshgrid(xs, ys, indexing="xy")
    return torch.stack((uu, vv), dim=-1)


def head_pos_embed(channels: int, h: int, w: int, image_w: int, image_h: int) -> torch.Tensor:
    """`DPTHead._apply_pos_embed`'s addend: [channels, h, w] = 0.1 * [sincos(u) | sincos(v)] of the h x w grid."""
    grid = create_uv_grid(w, h, image_w / image_h).reshape(-1, 2)
    emb = torch.cat([make_sincos_pos_embed(channels // 2, grid[:, 0]), make_sincos_pos_embed(channels // 2, grid[:, 1])], dim=-1)
    return (emb.reshape(h, w, channels) * POS_RATIO).permute(2, 0, 1).contiguous()


def init_vggt(encoder: str = "vitl", depth: int = 24, features: int = 256, out_channels=(256, 512, 1024, 1024), seed: int = 0):
    """The tensors the exported graph needs, under upstream's names: the DINOv2-with-registers trunk (aggregator.patch_embed.*),
    the camera / register tokens (two variants each: first frame, other frames), `depth` frame + global blocks, and the depth
    head (LayerNorm over the 2 * D concatenated taps, DPT on them, two output channels: depth and confidence logits)."""
    from oracle import dav2_torch as O
    D = O.MODEL_CONFIGS[encoder]["embed_dim"]
    g = torch.Generator().manual_seed(seed + 1000)
    sd = {}
    trunk = O.init_state_dict(encoder, seed=seed + 1, registers=4)
    for k, v in trunk.items():
        if k.startswith("pretrained."):
            sd[TRUNK + k[len("pretrained."):]] = v
    sd["aggregator.camera_token"] = 0.5 * torch.randn(1, 2, 1, D, generator=g)
    sd["aggregator.register_token"] = 0.5 * torch.randn(1, 2, 4, D, generator=g)
    sd.update(init_aggregator(D, depth, seed=seed + 2))
    h = "depth_head."
    oc = list(out_channels)
    shapes = {h + "norm.weight": (2 * D,), h + "norm.bias": (2 * D,)}
    for i in range(4):
        shapes[h + f"projects.{i}.weight"] = (oc[i], 2 * D, 1, 1); shapes[h + f"projects.{i}.bias"] = (oc[i],)
        shapes[h + f"scratch.layer{i + 1}_rn.weight"] = (features, oc[i], 3, 3)
        r = h + f"scratch.refinenet{i + 1}."
        shapes[r + "out_conv.weight"] = (features, features, 1, 1); shapes[r + "out_conv.bias"] = (features,)
        for u in (("resConfUnit2",) if i == 3 else ("resConfUnit1", "resConfUnit2")):      # refinenet4: has_residual=False
            for cv in ("conv1", "conv2"):
                shapes[r + f"{u}.{cv}.weight"] = (features, features, 3, 3); shapes[r + f"{u}.{cv}.bias"] = (features,)
    shapes[h + "resize_layers.0.weight"] = (oc[0], oc[0], 4, 4); shapes[h + "resize_layers.0.bias"] = (oc[0],)
    shapes[h + "resize_layers.1.weight"] = (oc[1], oc[1], 2, 2); shapes[h + "resize_layers.1.bias"] = (oc[1],)
    shapes[h + "resize_layers.3.weight"] = (oc[3], oc[3], 3, 3); shapes[h + "resize_layers.3.bias"] = (oc[3],)
    shapes[h + "scratch.output_conv1.weight"] = (features // 2, features, 3, 3); shapes[h + "scratch.output_conv1.bias"] = (features // 2,)
    shapes[h + "scratch.output_conv2.0.weight"] = (32, features // 2, 3, 3); shapes[h + "scratch.output_conv2.0.bias"] = (32,)
    shapes[h + "scratch.output_conv2.2.weight"] = (2, 32, 1, 1); shapes[h + "scratch.output_conv2.2.bias"] = (2,)
    for k, shp in shapes.items():
        if "norm" in k and k.endswith("weight"):
            v = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            v = 0.1 * torch.randn(shp, generator=g)
        elif k.startswith(h + "resize_layers.") and k[len(h + "resize_layers.")] in "01":
            v = torch.randn(shp, generator=g) / math.sqrt(shp[0])
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            v = torch.randn(shp, generator=g) / math.sqrt(fan_in)
        sd[k] = v.float().contiguous()
    return sd


def trunk_state_dict(sd) -> Dict[str, torch.Tensor]:
    """The DINOv2 trunk's tensors under oracle/dav2_torch.py's key names."""
    return {"pretrained." + k[len(TRUNK):]: v for k, v in sd.items() if k.startswith(TRUNK)}


def special_tokens(sd, frames: int) -> torch.Tensor:
    """[frames, 5, D]: camera token then the four register tokens; frame 0 takes variant 0, every other frame variant 1
    (vggt/models/aggregator.py `slice_expand_and_flatten`)."""
    cam, reg = sd["aggregator.camera_token"][0], sd["aggregator.register_token"][0]          # [2, 1, D], [2, 4, D]
    per = torch.cat([cam, reg], dim=1)                                                        # [2, 5, D]
    return torch.cat([per[:1], per[1:].expand(frames - 1, -1, -1)], dim=0) if frames > 1 else per[:1]


@torch.no_grad()
def frame_tokens(sd, images: torch.Tensor, encoder: str = "vitl") -> torch.Tensor:
    """images float32 [S, 3, H, W] in 0..1 -> the aggregator's input [S, 5 + gh*gw, D]: ImageNet normalisation, the DINOv2 trunk's
    normalised patch tokens (`x_norm_patchtokens`), camera + register tokens in front."""
    from oracle import dav2_torch as O
    cfg = dict(O.MODEL_CONFIGS[encoder])
    cfg["taps"] = [cfg["depth"] - 1]
    mean = torch.tensor(RESNET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(RESNET_STD).view(1, 3, 1, 1)
    patch = O.encoder_taps(trunk_state_dict(sd), (images - mean) / std, cfg, norm_mask=0x1)[0]
    return torch.cat([special_tokens(sd, images.shape[0]), patch], dim=1)


def _rcu(sd, pre: str, x: torch.Tensor) -> torch.Tensor:
    y = F.conv2d(F.relu(x), sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
    y = F.conv2d(F.relu(y), sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1)
    return y + x


def _fusion(sd, i: int, x0, x1, size):
    r = f"depth_head.scratch.refinenet{i}."
    out = x0
    if x1 is not None:
        out = out + _rcu(sd, r + "resConfUnit1.", x1)
    out = _rcu(sd, r + "resConfUnit2.", out)
    out = F.interpolate(out, **({"scale_factor": 2} if size is None else {"size": size}), mode="bilinear", align_corners=True)
    return F.conv2d(out, sd[r + "out_conv.weight"], sd[r + "out_conv.bias"])


@torch.no_grad()
def depth_head(sd, taps: List[torch.Tensor], gh: int, gw: int, image_h: int, image_w: int, trace=None) -> torch.Tensor:
    """taps: the four [S, N, 2D] aggregator outputs (layers 4 / 11 / 17 / 23 of 24) -> pre-activation [S, 2, H, W]
    (channel 0: log depth, channel 1: confidence logit)."""
    h = "depth_head."
    l = []
    for i, t in enumerate(taps):
        x = t[:, N_SPECIAL:]
        S, _, C = x.shape
        x = F.layer_norm(x, (C,), sd[h + "norm.weight"], sd[h + "norm.bias"], HEAD_LN_EPS)
        f = x.permute(0, 2, 1).reshape(S, C, gh, gw)
        f = F.conv2d(f, sd[h + f"projects.{i}.weight"], sd[h + f"projects.{i}.bias"])
        f = f + head_pos_embed(f.shape[1], gh, gw, image_w, image_h)
        if i == 0:
            f = F.conv_transpose2d(f, sd[h + "resize_layers.0.weight"], sd[h + "resize_layers.0.bias"], stride=4)
        elif i == 1:
            f = F.conv_transpose2d(f, sd[h + "resize_layers.1.weight"], sd[h + "resize_layers.1.bias"], stride=2)
        elif i == 3:
            f = F.conv2d(f, sd[h + "resize_layers.3.weight"], sd[h + "resize_layers.3.bias"], stride=2, padding=1)
        l.append(f)
    r = [F.conv2d(l[i], sd[h + f"scratch.layer{i + 1}_rn.weight"], None, padding=1) for i in range(4)]
    if trace is not None:
        for i in range(4):
            trace[f"layer{i + 1}_rn"] = r[i]
    p = _fusion(sd, 4, r[3], None, r[2].shape[2:])
    p = _fusion(sd, 3, p, r[2], r[1].shape[2:])
    p = _fusion(sd, 2, p, r[1], r[0].shape[2:])
    p = _fusion(sd, 1, p, r[0], None)
    if trace is not None:
        trace["path_1"] = p
    out = F.conv2d(p, sd[h + "scratch.output_conv1.weight"], sd[h + "scratch.output_conv1.bias"], padding=1)
    out = F.interpolate(out, (gh * 14, gw * 14), mode="bilinear", align_corners=True)
    out = out + head_pos_embed(out.shape[1], gh * 14, gw * 14, image_w, image_h)
    out = F.relu(F.conv2d(out, sd[h + "scratch.output_conv2.0.weight"], sd[h + "scratch.output_conv2.0.bias"], padding=1))
    return F.conv2d(out, sd[h + "scratch.output_conv2.2.weight"], sd[h + "scratch.output_conv2.2.bias"])


@torch.no_grad()
def vggt_depth(sd, images: torch.Tensor, encoder: str = "vitl", depth: int = 24, taps=(4, 11, 17, 23), trace=None,
               causal: bool = False) -> torch.Tensor:
    """images float32 [S, 3, H, W] in 0..1 (one scene) -> depth [S, H, W] = exp(channel 0 of the head)  (`activate_head`,
    activation "exp"; the confidence channel is computed by the graph but not returned by the reference's wrapper).
    causal=True: StreamVGGT (see `aggregate`)."""
    from oracle import dav2_torch as O
    S, _, H, W = images.shape
    gh, gw = H // 14, W // 14
    tok = frame_tokens(sd, images, encoder)
    if trace is not None:
        trace["tokens"] = tok
    layers = aggregate(sd, tok, gh, gw, O.MODEL_CONFIGS[encoder]["num_heads"], depth, causal)
    if trace is not None:
        trace["aggregated"] = [layers[t] for t in taps]
    logits = depth_head(sd, [layers[t] for t in taps], gh, gw, H, W, trace)
    if trace is not None:
        trace["logits"] = logits
    return torch.exp(logits[:, 0])


@torch.no_grad()
def calibrate_vggt(sd, images: torch.Tensor, encoder: str, depth: int, taps) -> None:
    """Rescale / shift the last 1x1 conv so that the log-depth channel on `images` is ~N(0, 0.5^2): depths spread over
    roughly 0.2 .. 5, positive by construction (fixed once, never tuned to a kernel)."""
    from oracle import dav2_torch as O
    S, _, H, W = images.shape
    gh, gw = H // 14, W // 14
    tok = frame_tokens(sd, images, encoder)
    layers = aggregate(sd, tok, gh, gw, O.MODEL_CONFIGS[encoder]["num_heads"], depth)
    z = depth_head(sd, [layers[t] for t in taps], gh, gw, H, W)[:, 0]
    m, s_ = float(z.mean()), float(z.std())
    w, b = "depth_head.scratch.output_conv2.2.weight", "depth_head.scratch.output_conv2.2.bias"
    sd[w] = sd[w].clone(); sd[b] = sd[b].clone()
    sd[w][0] = sd[w][0] * (0.5 / s_)
    sd[b][0] = (sd[b][0] - m) * (0.5 / s_)
