"""Weights container for the B200 runtime: the artefact of the "export" stage.

In the reference `run.py export <model>` runs models/<name>/onnx_export.py in the model's own
environment and leaves an ONNX file that `run.py build` hands to TensorRT
(run.py:40-44, models/depth_anything_v2/onnx_export.py:24-65).  Here export leaves an `.mdew`
file instead: the upstream state dict (key names untouched, fp32) plus a JSON description of the
architecture and the input size.  Build (`core.common.get_engine` replacement) needs nothing
else -- in particular not the upstream package, keeping the reference's environment separation
(run.py:9-15).

Layout:  b"MDEW0001" | u32 meta_len | meta JSON (utf-8) | u32 count |
         count x { u32 name_len | name | u32 ndim | i64 dims[ndim] | f32 data[] }
"""
from __future__ import annotations

import hashlib
import json
import math
import struct
from typing import Dict, Mapping, Tuple

import numpy as np

MAGIC = b"MDEW0001"

# models/depth_anything_v2/infer.py:55-60 -- the three released encoder sizes, plus the DINOv2
# widths/depths/head counts and tap indices that go with them (reports/profile/*.json layer lists).
ENCODERS = {
    "vits": dict(embed_dim=384, depth=12, num_heads=6, features=64,
                 out_channels=[48, 96, 192, 384], taps=[2, 5, 8, 11]),
    "vitb": dict(embed_dim=768, depth=12, num_heads=12, features=128,
                 out_channels=[96, 192, 384, 768], taps=[2, 5, 8, 11]),
    "vitl": dict(embed_dim=1024, depth=24, num_heads=16, features=256,
                 out_channels=[256, 512, 1024, 1024], taps=[4, 11, 17, 23]),
}


def describe(encoder: str, input_h: int = 518, input_w: int = 518, max_depth: float | None = 20.0,
             patch_size: int = 14) -> dict:
    """The architecture description stored next to the tensors."""
    if encoder not in ENCODERS:
        raise KeyError(f"unknown encoder {encoder!r}; available: {sorted(ENCODERS)}")
    if input_h % patch_size or input_w % patch_size:
        raise ValueError(f"input size {input_h}x{input_w} is not a multiple of the patch size {patch_size}")
    meta = dict(ENCODERS[encoder])
    meta.update(family="depth_anything_v2", encoder=encoder, patch_size=patch_size,
                input_h=int(input_h), input_w=int(input_w),
                max_depth=None if max_depth is None else float(max_depth))
    return meta


def describe_depth_pro(encoder: str = "vitl", features: int = 256, hook_blocks=(11, 5), image_size: int = 1536) -> dict:
    """Description stored next to a Depth Pro state dict (models/depth_pro/onnx_export.py:15-22: three `dinov2l16_384`
    trunks, decoder_features 256, field-of-view head; 1536 x 1536 is upstream's fixed size, spec.json "caveats")."""
    if encoder not in ENCODERS:
        raise KeyError(f"unknown encoder {encoder!r}; available: {sorted(ENCODERS)}")
    c = ENCODERS[encoder]
    return dict(family="depth_pro", encoder=encoder, embed_dim=c["embed_dim"], depth=c["depth"], num_heads=c["num_heads"],
                patch_size=16, features=int(features), hook_blocks=[int(b) for b in hook_blocks],
                input_h=int(image_size), input_w=int(image_size))


def describe_vggt(encoder: str = "vitl", depth: int = 24, features: int = 256, out_channels=(256, 512, 1024, 1024),
                  taps=(4, 11, 17, 23), frames: int = 1, image_hw=(518, 518), family: str = "vggt", stream_frames: int = 0) -> dict:
    """Description stored next to a VGGT (`family="vggt"`, models/vggt/onnx_export.py:38-52) or StreamVGGT (`"streamvggt"`,
    models/streamvggt/onnx_export.py:35-53: temporal causal attention in the global blocks) state dict: trunk, aggregator depth,
    DPT widths, tapped layers, frames per execute (the reference exports S = 1: spec.json rank-5 [1, 1, 3, 518, 518]).
    stream_frames > 0 (StreamVGGT only): the streaming engine -- one frame per execute against a key / value cache that long."""
    if encoder not in ENCODERS:
        raise KeyError(f"unknown encoder {encoder!r}; available: {sorted(ENCODERS)}")
    if family not in ("vggt", "streamvggt"):
        raise ValueError(f"unknown family {family!r}")
    if stream_frames and family != "streamvggt":
        raise ValueError("only StreamVGGT has a streaming form")
    c = ENCODERS[encoder]
    return dict(family=family, encoder=encoder, embed_dim=c["embed_dim"], num_heads=c["num_heads"], patch_size=14,
                aggregator_depth=int(depth), features=int(features), out_channels=[int(x) for x in out_channels],
                taps=[int(t) for t in taps], frames=int(frames), input_h=int(image_hw[0]), input_w=int(image_hw[1]),
                stream_frames=int(stream_frames))


def keep_ratio_size(src_h: int, src_w: int, target: int = 518, multiple: int = 14, rounding: str = "ceil") -> Tuple[int, int]:
    """The engine size for a source frame under the keep-ratio rule of the Depth Anything family (core/preprocess.py:157-171
    `resize_keep_ratio`, bound "lower", with :112-137 `_round_to_multiple`): short side to `target`, both sides snapped to
    a multiple of the patch size -- "ceil" for depth_anything_ac (4:3 -> 518 x 700), "constrain" for depth_anything_v2 (nearest,
    never below the target: 4:3 -> 518 x 686)."""
    scale = target / min(src_h, src_w)

    def snap(x: float) -> int:
        if rounding == "constrain":
            y = int(round(x / multiple)) * multiple if abs(x / multiple - round(x / multiple)) != 0.5 else int(np.round(x / multiple) * multiple)
            if y < target:
                y = int(math.ceil(x / multiple) * multiple)
            return max(y, multiple)
        if rounding == "ceil":
            return max(int(math.ceil(x / multiple)) * multiple, multiple)
        if rounding == "floor":
            return max(int(math.floor(x / multiple)) * multiple, multiple)
        raise ValueError(f"[MDET] unknown rounding {rounding!r}")

    return snap(src_h * scale), snap(src_w * scale)


def resize_pos_embed(pos_embed: np.ndarray, gh: int, gw: int, mode: str = "dinov2") -> np.ndarray:
    """DINOv2's position-embedding rule for a grid other than the trained square one: bicubic
    resize of the patch part with scale (g + 0.1) / m, cls part untouched.  At the trained grid
    (37 x 37 for 518 / 14) the table is used as is.  mode="bilinear": what the reference's Metric3D V2 export does instead
    (reports/profile/metric3d_v2.json layer 6 `/depth_model/encoder/Resize`: LINEAR, half-pixel, to the 44 x 76 grid by size)."""
    n = pos_embed.shape[1] - 1
    m = int(round(math.sqrt(n)))
    if gh * gw == n and gh == gw:
        return np.ascontiguousarray(pos_embed, dtype=np.float32)
    import torch
    import torch.nn.functional as F
    pe = torch.from_numpy(np.asarray(pos_embed, dtype=np.float32))
    d = pe.shape[-1]
    if mode == "bilinear":
        patch = F.interpolate(pe[:, 1:].reshape(1, m, m, d).permute(0, 3, 1, 2), size=(gh, gw), mode="bilinear", align_corners=False)
    elif mode == "dinov2":
        patch = F.interpolate(pe[:, 1:].reshape(1, m, m, d).permute(0, 3, 1, 2),
                              scale_factor=((gh + 0.1) / m, (gw + 0.1) / m), mode="bicubic", antialias=False)
    else:
        raise ValueError(f"[MDET] unknown position-embedding resize mode {mode!r}")
    if tuple(patch.shape[-2:]) != (gh, gw):
        raise ValueError(f"position embedding resize produced {tuple(patch.shape[-2:])}, wanted {(gh, gw)}")
    patch = patch.permute(0, 2, 3, 1).reshape(1, gh * gw, d)
    return torch.cat([pe[:, :1], patch], dim=1).contiguous().numpy()


def _as_numpy(v) -> np.ndarray:
    if hasattr(v, "detach"):
        v = v.detach().cpu().float().numpy()
    return np.ascontiguousarray(v, dtype=np.float32)


def save(path: str, state_dict: Mapping[str, object], meta: dict) -> str:
    """Write an .mdew file; returns the sha256 of its bytes."""
    h = hashlib.sha256()
    with open(path, "wb") as f:
        def w(b: bytes):
            f.write(b)
            h.update(b)
        mj = json.dumps(meta, sort_keys=True).encode("utf-8")
        w(MAGIC)
        w(struct.pack("<I", len(mj)))
        w(mj)
        items = [(k, _as_numpy(v)) for k, v in state_dict.items()]
        w(struct.pack("<I", len(items)))
        for name, arr in items:
            nb = name.encode("utf-8")
            w(struct.pack("<I", len(nb)))
            w(nb)
            w(struct.pack("<I", arr.ndim))
            w(struct.pack(f"<{arr.ndim}q", *arr.shape))
            w(arr.tobytes())
    return h.hexdigest()


def read_meta(path: str) -> dict:
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"[MDET] {path} is not an MDEW0001 weights file")
        (n,) = struct.unpack("<I", f.read(4))
        return json.loads(f.read(n).decode("utf-8"))


def load(path: str) -> Tuple[Dict[str, np.ndarray], dict]:
    """Read an .mdew file back (tests and tools; the engine reads the file natively)."""
    out: Dict[str, np.ndarray] = {}
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"[MDET] {path} is not an MDEW0001 weights file")
        (n,) = struct.unpack("<I", f.read(4))
        meta = json.loads(f.read(n).decode("utf-8"))
        (count,) = struct.unpack("<I", f.read(4))
        for _ in range(count):
            (ln,) = struct.unpack("<I", f.read(4))
            name = f.read(ln).decode("utf-8")
            (nd,) = struct.unpack("<I", f.read(4))
            dims = struct.unpack(f"<{nd}q", f.read(8 * nd))
            cnt = int(np.prod(dims)) if nd else 1
            out[name] = np.frombuffer(f.read(4 * cnt), dtype="<f4").reshape(dims).copy()
    return out, meta


def load_tensor(path: str, wanted: str) -> np.ndarray:
    """Read ONE tensor of an .mdew file, seeking past the data of the others."""
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"[MDET] {path} is not an MDEW0001 weights file")
        (n,) = struct.unpack("<I", f.read(4))
        f.seek(n, 1)
        (count,) = struct.unpack("<I", f.read(4))
        for _ in range(count):
            (ln,) = struct.unpack("<I", f.read(4))
            name = f.read(ln).decode("utf-8")
            (nd,) = struct.unpack("<I", f.read(4))
            dims = struct.unpack(f"<{nd}q", f.read(8 * nd))
            cnt = int(np.prod(dims)) if nd else 1
            if name == wanted:
                return np.frombuffer(f.read(4 * cnt), dtype="<f4").reshape(dims).copy()
            f.seek(4 * cnt, 1)
    raise KeyError(f"[MDET] {path} holds no tensor named {wanted!r}")


def file_sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


# ------------------------------------------------------------------------------------------ upstream checkpoints
def encoder_of(state_dict: Mapping[str, object]) -> str:
    """Which of the released encoder sizes a state dict is, from the width of the trunk and its depth
    (models/depth_anything_v2/infer.py:55-60 picks the same configs by name)."""
    try:
        dim = int(_shape(state_dict["pretrained.cls_token"])[-1])
    except KeyError:
        raise ValueError("[MDET] not a Depth Anything V2 state dict: 'pretrained.cls_token' is missing")
    depth = 1 + max(int(k.split(".")[2]) for k in state_dict if k.startswith("pretrained.blocks."))
    for name, cfg in ENCODERS.items():
        if cfg["embed_dim"] == dim and cfg["depth"] == depth:
            return name
    raise ValueError(f"[MDET] no released encoder has width {dim} and depth {depth}")


def _shape(v):
    return tuple(v.shape)


def from_hf_depth_anything(hf_sd: Mapping[str, object]) -> Dict[str, object]:
    """transformers' `DepthAnythingForDepthEstimation` key set (the `*-hf` checkpoints on the hub, e.g.
    depth-anything/Depth-Anything-V2-Small-hf) -> the upstream key names models/depth_anything_v2/infer.py:62-63 loads and
    the engine takes: separate query / key / value projections are stacked into `attn.qkv`, `neck.fusion_stage.layers.k` is
    `refinenet{4-k}`, `neck.reassemble_stage.layers.i` the `projects` / `resize_layers` pair, `head.conv1..3` the output convs.
    Tensors are passed through untouched (torch or numpy); tests/test_weights.py pins the mapping on transformers' own forward."""
    import numpy as np

    def cat(parts):
        if hasattr(parts[0], "detach"):
            import torch
            return torch.cat(list(parts), dim=0)
        return np.concatenate([np.asarray(p_) for p_ in parts], axis=0)

    g = hf_sd.__getitem__
    o = {}
    e = "backbone.embeddings."
    o["pretrained.cls_token"], o["pretrained.pos_embed"] = g(e + "cls_token"), g(e + "position_embeddings")
    if e + "mask_token" in hf_sd:
        o["pretrained.mask_token"] = g(e + "mask_token")
    o["pretrained.patch_embed.proj.weight"], o["pretrained.patch_embed.proj.bias"] = g(e + "patch_embeddings.projection.weight"), g(e + "patch_embeddings.projection.bias")
    layers = sorted({int(k.split(".")[3]) for k in hf_sd if k.startswith("backbone.encoder.layer.")})
    if not layers:
        raise ValueError("[MDET] not a transformers DepthAnything state dict: no 'backbone.encoder.layer.*' tensors")
    for i in layers:
        p_, q = f"pretrained.blocks.{i}.", f"backbone.encoder.layer.{i}."
        for n in ("norm1", "norm2"):
            o[p_ + n + ".weight"], o[p_ + n + ".bias"] = g(q + n + ".weight"), g(q + n + ".bias")
        o[p_ + "attn.qkv.weight"] = cat([g(q + f"attention.attention.{n}.weight") for n in ("query", "key", "value")])
        o[p_ + "attn.qkv.bias"] = cat([g(q + f"attention.attention.{n}.bias") for n in ("query", "key", "value")])
        o[p_ + "attn.proj.weight"], o[p_ + "attn.proj.bias"] = g(q + "attention.output.dense.weight"), g(q + "attention.output.dense.bias")
        o[p_ + "ls1.gamma"], o[p_ + "ls2.gamma"] = g(q + "layer_scale1.lambda1"), g(q + "layer_scale2.lambda1")
        for n in ("fc1", "fc2"):
            o[p_ + f"mlp.{n}.weight"], o[p_ + f"mlp.{n}.bias"] = g(q + f"mlp.{n}.weight"), g(q + f"mlp.{n}.bias")
    o["pretrained.norm.weight"], o["pretrained.norm.bias"] = g("backbone.layernorm.weight"), g("backbone.layernorm.bias")
    h = "depth_head."
    for i in range(4):
        o[h + f"projects.{i}.weight"] = g(f"neck.reassemble_stage.layers.{i}.projection.weight")
        o[h + f"projects.{i}.bias"] = g(f"neck.reassemble_stage.layers.{i}.projection.bias")
        if i != 2:
            o[h + f"resize_layers.{i}.weight"] = g(f"neck.reassemble_stage.layers.{i}.resize.weight")
            o[h + f"resize_layers.{i}.bias"] = g(f"neck.reassemble_stage.layers.{i}.resize.bias")
        o[h + f"scratch.layer{i + 1}_rn.weight"] = g(f"neck.convs.{i}.weight")
        r, f = h + f"scratch.refinenet{i + 1}.", f"neck.fusion_stage.layers.{3 - i}."
        o[r + "out_conv.weight"], o[r + "out_conv.bias"] = g(f + "projection.weight"), g(f + "projection.bias")
        for u, hu in (("resConfUnit1", "residual_layer1"), ("resConfUnit2", "residual_layer2")):
            for cv, hc in (("conv1", "convolution1"), ("conv2", "convolution2")):
                o[r + f"{u}.{cv}.weight"], o[r + f"{u}.{cv}.bias"] = g(f + f"{hu}.{hc}.weight"), g(f + f"{hu}.{hc}.bias")
    o[h + "scratch.output_conv1.weight"], o[h + "scratch.output_conv1.bias"] = g("head.conv1.weight"), g("head.conv1.bias")
    o[h + "scratch.output_conv2.0.weight"], o[h + "scratch.output_conv2.0.bias"] = g("head.conv2.weight"), g("head.conv2.bias")
    o[h + "scratch.output_conv2.2.weight"], o[h + "scratch.output_conv2.2.bias"] = g("head.conv3.weight"), g("head.conv3.bias")
    return o


def export_checkpoint(checkpoint_path: str, out_path: str, input_h: int = 518, input_w: int = 518,
                      max_depth: float | None = None) -> dict:
    """Stage `export` for an upstream checkpoint (models/depth_anything_v2/infer.py:62-63:
    `model.load_state_dict(torch.load(f"checkpoints/depth_anything_v2_{encoder}.pth"))`) or its transformers twin
    (`model.safetensors` / `pytorch_model.bin` of a `*-hf` repository, key names mapped by `from_hf_depth_anything`): read the
    file, keep the tensors the engine needs under their upstream key names, write the .mdew file.  `max_depth` None = the relative
    model, 20 / 80 = the metric Hypersim / VKITTI heads (infer_metric.py:61-65).  Returns the stored description."""
    import torch
    if checkpoint_path.endswith(".safetensors"):          # the hub's `*-hf` repositories ship model.safetensors
        from safetensors.torch import load_file
        sd = load_file(checkpoint_path, device="cpu")
    else:
        sd = torch.load(checkpoint_path, map_location="cpu", weights_only=True)
    if isinstance(sd, dict) and "state_dict" in sd and "pretrained.cls_token" not in sd:
        sd = sd["state_dict"]
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    if "backbone.embeddings.cls_token" in sd:             # transformers' key names
        sd = from_hf_depth_anything(sd)
    enc = encoder_of(sd)
    keep = {k: v for k, v in sd.items() if k.startswith(("pretrained.", "depth_head."))}
    meta = describe(enc, input_h, input_w, max_depth)
    meta["source_checkpoint_sha256"] = file_sha256(checkpoint_path)
    save(out_path, keep, meta)
    return meta
