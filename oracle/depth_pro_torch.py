"""ORACLE (test infrastructure, never shipped, never on the product path).

Depth Pro's patch-encoder stage, restated: the reference runs `depth_pro.create_model_and_transforms` from the
un-vendored github.com/apple/ml-depth-pro package (models/depth_pro/onnx_export.py:2,15-29: `dinov2l16_384` patch /
image encoders, 1536 x 1536 input).  Its encoder (`depth_pro/network/encoder.py`: `split`, `merge`, the two hooked
blocks) is public; transformers carries an independent implementation (`DepthProPatchEncoder`) that loads the released
checkpoint, and tests/test_oracle_depth_pro.py pins this file against it on copied weights.

  image [3, 1536, 1536] -> pyramid 1, 1/2, 1/4 (bilinear, align_corners=False) -> 25 + 9 + 1 crops of 384 x 384
  -> shared ViT/16 trunk -> per crop: normalised final tokens and the raw outputs of two hooked blocks, 24 x 24 x D
  -> merged maps: 96 x 96 (final, full-resolution crops, 3 tokens trimmed at every inner edge), 48 x 48 (half
     resolution, 6 trimmed), 24 x 24 (quarter resolution), and 96 x 96 for each hook (full-resolution crops only).
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from oracle import dav2_torch as O

GRID = 24           # tokens per crop side (384 / 16)
PATCH_PX = 384      # crop side in pixels
MERGE_PADDING = 3   # tokens trimmed per inner edge at full resolution; x2, x4 at the lower levels (capped at GRID // 4)


def merge_crops(tokens: torch.Tensor, per_side: int, padding: int) -> torch.Tensor:
    """tokens [per_side**2, GRID*GRID, D] (row-major crops) -> [S, S, D] with S = per_side*GRID - 2*padding*(per_side-1).
    Inner edges of every crop lose `padding` tokens (upstream `merge`, transformers `merge_patches`)."""
    n, t, d = tokens.shape
    assert n == per_side * per_side and t == GRID * GRID
    padding = min(GRID // 4, padding) if per_side > 1 else 0
    grid = tokens.reshape(per_side, per_side, GRID, GRID, d)
    rows = []
    for h in range(per_side):
        top = padding if h != 0 else 0
        bottom = GRID - (padding if h != per_side - 1 else 0)
        cols = []
        for w in range(per_side):
            left = padding if w != 0 else 0
            right = GRID - (padding if w != per_side - 1 else 0)
            cols.append(grid[h, w, top:bottom, left:right])
        rows.append(torch.cat(cols, dim=1))
    return torch.cat(rows, dim=0)


def merged_features(taps: Sequence[torch.Tensor], hook_taps: Sequence[int] = (1, 0), final_tap: int = 3) -> List[torch.Tensor]:
    """taps: the trunk's four outputs [35, 576, D] each (crop order: 25 full-resolution, 9 half, 1 quarter).
    -> [f24, f48, f96, hook_a 96, hook_b 96] as [S, S, D], the order transformers' DepthProPatchEncoder returns them
    (low resolution first, then the hooks in the configured order -- (11, 5) for the released model = taps (1, 0))."""
    fin = taps[final_tap]
    out = [merge_crops(fin[34:35], 1, 0), merge_crops(fin[25:34], 3, 2 * MERGE_PADDING), merge_crops(fin[0:25], 5, MERGE_PADDING)]
    for t in hook_taps:
        out.append(merge_crops(taps[t][0:25], 5, MERGE_PADDING))
    return out


@torch.no_grad()
def patch_encoder_features(sd, image: torch.Tensor, encoder: str, hook_taps: Sequence[int] = (1, 0)) -> List[torch.Tensor]:
    """image [3, S, S] float32 (normalised) -> the five merged maps, through the oracle trunk."""
    from monocular_depth_estimation_trt_b200 import sharding as S      # crop geometry only (pure host code)
    crops = S.make_crops(image)
    taps = O.encoder_taps(sd, crops, O.MODEL_CONFIGS[encoder], norm_mask=0x8)
    return merged_features(taps, hook_taps)


# ------------------------------------------------------------------------------------------------ the whole model
# Key names follow the module tree of the un-vendored apple/ml-depth-pro package (`DepthPro`: encoder / decoder / head /
# fov, the model models/depth_pro/onnx_export.py:15-29 builds and exports with outputs "canonical_inverse_depth" and
# "fov_deg", :58-59) as publicly documented by its checkpoint layout; the arithmetic is pinned against transformers'
# independent `DepthProForDepthEstimation` on copied weights (tests/test_oracle_depth_pro.py, tests/hf_bridge.py).

import math

import torch.nn.functional as F

TRUNKS = ("encoder.patch_encoder.", "encoder.image_encoder.", "fov.encoder.0.")


def decoder_dims(encoder: str, features: int):
    """Channel widths of the five decoder inputs, high resolution first (upstream: [256, 256, 512, 1024, 1024])."""
    D = O.MODEL_CONFIGS[encoder]["embed_dim"]
    return [features, features, D // 2, D, D]


def full_param_shapes(encoder: str, features: int = 256):
    D = O.MODEL_CONFIGS[encoder]["embed_dim"]
    Fd = features
    s = {}
    trunk = {k[len("pretrained."):]: v for k, v in O.param_shapes(encoder, patch=16, pos_grid=GRID).items() if k.startswith("pretrained.")}
    for pre in TRUNKS:
        for k, v in trunk.items():
            s[pre + k] = v
    e = "encoder."
    # 1x1 projection (no bias) followed by n ConvTranspose2d(k=2, s=2, no bias): [in, out, 2, 2]
    for name, mid, n_up in (("upsample_latent0", Fd, 3), ("upsample_latent1", Fd, 2), ("upsample0", D // 2, 1),
                            ("upsample1", D, 1), ("upsample2", D, 1)):
        s[e + f"{name}.0.weight"] = (mid, D, 1, 1)
        for j in range(n_up):
            s[e + f"{name}.{j + 1}.weight"] = (mid, mid, 2, 2)
    s[e + "upsample_lowres.weight"] = (D, D, 2, 2)
    s[e + "upsample_lowres.bias"] = (D,)
    s[e + "fuse_lowres.weight"] = (D, 2 * D, 1, 1)
    s[e + "fuse_lowres.bias"] = (D,)
    dims = decoder_dims(encoder, Fd)
    for i in range(1, 5):                                   # convs.0 is the identity (dims[0] == features)
        s[f"decoder.convs.{i}.weight"] = (Fd, dims[i], 3, 3)
    for i in range(5):
        f = f"decoder.fusions.{i}."
        for r in (("resnet1",) if i < 4 else ()) + ("resnet2",):
            for j in (1, 3):
                s[f + f"{r}.residual.{j}.weight"] = (Fd, Fd, 3, 3)
                s[f + f"{r}.residual.{j}.bias"] = (Fd,)
        if i > 0:
            s[f + "deconv.weight"] = (Fd, Fd, 2, 2)
        s[f + "out_conv.weight"] = (Fd, Fd, 1, 1)
        s[f + "out_conv.bias"] = (Fd,)
    s["head.0.weight"] = (Fd // 2, Fd, 3, 3); s["head.0.bias"] = (Fd // 2,)
    s["head.1.weight"] = (Fd // 2, Fd // 2, 2, 2); s["head.1.bias"] = (Fd // 2,)
    s["head.2.weight"] = (32, Fd // 2, 3, 3); s["head.2.bias"] = (32,)
    s["head.4.weight"] = (1, 32, 1, 1); s["head.4.bias"] = (1,)
    s["fov.encoder.1.weight"] = (Fd // 2, D); s["fov.encoder.1.bias"] = (Fd // 2,)
    s["fov.downsample.0.weight"] = (Fd // 2, Fd, 3, 3); s["fov.downsample.0.bias"] = (Fd // 2,)
    s["fov.head.0.weight"] = (Fd // 4, Fd // 2, 3, 3); s["fov.head.0.bias"] = (Fd // 4,)
    s["fov.head.2.weight"] = (Fd // 8, Fd // 4, 3, 3); s["fov.head.2.bias"] = (Fd // 8,)
    s["fov.head.4.weight"] = (1, Fd // 8, 6, 6); s["fov.head.4.bias"] = (1,)
    return s


def init_full_state_dict(encoder: str, features: int = 256, seed: int = 0):
    """Seeded non-degenerate init, the same recipe as oracle/dav2_torch.py `init_state_dict` (fan-in normal weights,
    N(0, 0.1) biases, LayerScale 0.5); transposed convolutions with kernel == stride see `in` taps per output pixel."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in full_param_shapes(encoder, features).items():
        transposed = len(shp) == 4 and shp[2:] == (2, 2)
        if k.endswith("gamma"):
            v = torch.full(shp, 0.5)
        elif "norm" in k and k.endswith("weight"):
            v = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            v = 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(("cls_token", "pos_embed", "mask_token")):
            v = 0.5 * torch.randn(shp, generator=g)
        elif transposed:
            v = torch.randn(shp, generator=g) / math.sqrt(shp[0])
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            v = torch.randn(shp, generator=g) / math.sqrt(fan_in)
        sd[k] = v.float().contiguous()
    return sd


def trunk_state_dict(sd, prefix: str):
    """One of the three ViT trunks under the key names oracle/dav2_torch.py (and the engine) use."""
    return {"pretrained." + k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def _nchw(m: torch.Tensor) -> torch.Tensor:
    return m.permute(2, 0, 1)[None]


def _residual(sd, pre: str, x: torch.Tensor) -> torch.Tensor:
    y = F.conv2d(F.relu(x), sd[pre + "residual.1.weight"], sd[pre + "residual.1.bias"], padding=1)
    y = F.conv2d(F.relu(y), sd[pre + "residual.3.weight"], sd[pre + "residual.3.bias"], padding=1)
    return y + x


def _upsample(sd, name: str, x: torch.Tensor) -> torch.Tensor:
    x = F.conv2d(x, sd[f"encoder.{name}.0.weight"])
    j = 1
    while f"encoder.{name}.{j}.weight" in sd:
        x = F.conv_transpose2d(x, sd[f"encoder.{name}.{j}.weight"], stride=2)
        j += 1
    return x


@torch.no_grad()
def full_forward(sd, x: torch.Tensor, encoder: str, hook_taps: Sequence[int] = (1, 0), trace: dict | None = None):
    """x float32 [1, 3, S, S] (normalised; S = 1536 for the exported model) -> (canonical_inverse_depth [1, 1, S, S],
    fov_deg [1]): the two outputs of the reference's engine (models/depth_pro/onnx_export.py:58-59, spec.json)."""
    cfg = O.MODEL_CONFIGS[encoder]
    image = x[0]
    f24, f48, f96, hook_a, hook_b = [_nchw(m) for m in
                                     patch_encoder_features(trunk_state_dict(sd, TRUNKS[0]), image, encoder, hook_taps)]
    low = F.interpolate(x, size=(PATCH_PX, PATCH_PX), mode="bilinear", align_corners=False)
    img_tok = O.encoder_taps(trunk_state_dict(sd, TRUNKS[1]), low, cfg, norm_mask=0x8)[3]          # [1, 576, D]
    img = img_tok.reshape(1, GRID, GRID, -1).permute(0, 3, 1, 2)
    # encoder: bring every map to its decoder resolution
    lat0 = _upsample(sd, "upsample_latent0", hook_b)          # 96 -> 768
    lat1 = _upsample(sd, "upsample_latent1", hook_a)          # 96 -> 384
    x0 = _upsample(sd, "upsample0", f96)                      # 96 -> 192
    x1 = _upsample(sd, "upsample1", f48)                      # 48 -> 96
    x2 = _upsample(sd, "upsample2", f24)                      # 24 -> 48
    xg = F.conv_transpose2d(img, sd["encoder.upsample_lowres.weight"], sd["encoder.upsample_lowres.bias"], stride=2)
    xg = F.conv2d(torch.cat((x2, xg), dim=1), sd["encoder.fuse_lowres.weight"], sd["encoder.fuse_lowres.bias"])
    enc = [lat0, lat1, x0, x1, xg]
    # decoder: project to `features` channels, fuse from the lowest resolution up
    proj = [enc[0]] + [F.conv2d(enc[i], sd[f"decoder.convs.{i}.weight"], padding=1) for i in range(1, 5)]
    lowres = proj[4]
    feat = None
    for i in (4, 3, 2, 1, 0):
        f = f"decoder.fusions.{i}."
        if feat is None:
            feat = proj[i]
        else:
            feat = feat + _residual(sd, f + "resnet1.", proj[i])
        feat = _residual(sd, f + "resnet2.", feat)
        if i > 0:
            feat = F.conv_transpose2d(feat, sd[f + "deconv.weight"], stride=2)
        feat = F.conv2d(feat, sd[f + "out_conv.weight"], sd[f + "out_conv.bias"])
        if trace is not None:
            trace[f"fusion{i}"] = feat
    h = F.conv2d(feat, sd["head.0.weight"], sd["head.0.bias"], padding=1)
    h = F.conv_transpose2d(h, sd["head.1.weight"], sd["head.1.bias"], stride=2)
    h = F.relu(F.conv2d(h, sd["head.2.weight"], sd["head.2.bias"], padding=1))
    inv = F.relu(F.conv2d(h, sd["head.4.weight"], sd["head.4.bias"]))
    # field of view: a third trunk on the quarter-resolution image + the decoder's low-resolution features
    tok = O.encoder_taps(trunk_state_dict(sd, TRUNKS[2]), low, cfg, norm_mask=0x8)[3]
    tok = F.linear(tok, sd["fov.encoder.1.weight"], sd["fov.encoder.1.bias"])                        # [1, 576, F/2]
    fv = tok.reshape(1, GRID, GRID, -1).permute(0, 3, 1, 2)
    fv = fv + F.relu(F.conv2d(lowres, sd["fov.downsample.0.weight"], sd["fov.downsample.0.bias"], stride=2, padding=1))
    fv = F.relu(F.conv2d(fv, sd["fov.head.0.weight"], sd["fov.head.0.bias"], stride=2, padding=1))
    fv = F.relu(F.conv2d(fv, sd["fov.head.2.weight"], sd["fov.head.2.bias"], stride=2, padding=1))
    fov = F.conv2d(fv, sd["fov.head.4.weight"], sd["fov.head.4.bias"]).flatten()
    if trace is not None:
        trace.update(lowres=lowres, features=feat, enc=enc)
    return inv, fov


def postprocess(inv: torch.Tensor, fov_deg: torch.Tensor, src_h: int, src_w: int):
    """models/depth_pro/onnx2trt.py:118-134: f_px from the predicted field of view, inverse depth scaled by W / f_px,
    resized back to the source size (bilinear, align_corners=False), depth = 1 / clamp(inverse, 1e-4, 1e4)."""
    f_px = 0.5 * src_w / torch.tan(0.5 * torch.deg2rad(fov_deg.float()))
    inverse = inv * (src_w / f_px)
    if inv.shape[-2:] != (src_h, src_w):
        inverse = F.interpolate(inverse, size=(src_h, src_w), mode="bilinear", align_corners=False)
    return 1.0 / torch.clamp(inverse, min=1e-4, max=1e4), f_px.squeeze()


def preprocess(image_rgb_u8, size: int = 1536) -> torch.Tensor:
    """models/depth_pro/onnx2trt.py:56-74: Compose([ToTensor(), Normalize([0.5]*3, [0.5]*3)]) on the RGB uint8 image, then
    F.interpolate(bilinear, align_corners=False) to size x size when the source is not already that size.
    (ToTensor = HWC uint8 -> CHW float32 / 255; Normalize = (x - mean) / std, both in fp32.)"""
    t = torch.from_numpy(image_rgb_u8).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    t = (t - 0.5) / 0.5
    t = t[None]
    if t.shape[-2:] != (size, size):
        t = F.interpolate(t, size=(size, size), mode="bilinear", align_corners=False)
    return t


@torch.no_grad()
def calibrate_full(sd, x: torch.Tensor, encoder: str, hook_taps: Sequence[int] = (1, 0)):
    """Random-init weights leave the final ReLU half dead and the field of view near zero; rescale the last 1x1 convolution
    so the canonical inverse depth on `x` is ~N(3, 0.5^2) (positive everywhere, like the trained model's output) and shift
    the last field-of-view bias so the prediction on `x` is 60 degrees.  Fixed once, before any kernel existed."""
    trace = {}
    full_forward(sd, x, encoder, hook_taps, trace)
    h = trace["features"]
    h = F.conv2d(h, sd["head.0.weight"], sd["head.0.bias"], padding=1)
    h = F.conv_transpose2d(h, sd["head.1.weight"], sd["head.1.bias"], stride=2)
    h = F.relu(F.conv2d(h, sd["head.2.weight"], sd["head.2.bias"], padding=1))
    z = F.conv2d(h, sd["head.4.weight"], sd["head.4.bias"])
    m, s = float(z.mean()), float(z.std())
    sd["head.4.weight"] = (sd["head.4.weight"] * (0.5 / s)).contiguous()
    sd["head.4.bias"] = ((sd["head.4.bias"] - m) * (0.5 / s) + 3.0).contiguous()
    _, fov = full_forward(sd, x, encoder, hook_taps)
    sd["fov.head.4.bias"] = (sd["fov.head.4.bias"] + (60.0 - float(fov))).contiguous()
    return m, s
