"""`get_engine` for the B200 runtime -- drop-in for the reference's core/common.py:141-312.

The reference parses an ONNX file and lets TensorRT search tactics for minutes, then caches the
serialized plan next to a fingerprint.  Here "build" is: read the exported weights (.mdew, see
weights.py), describe the engine (encoder, input size, precision, batch), pack the weights into the
kernels' layouts and upload them.  That takes seconds, so nothing heavy is cached; what *is* kept is
the reference's bookkeeping, because its tooling reads it:

  <engine>.engine        a small JSON stub describing what was built (stands in for the plan file
                         that core/build_conditions.stamp() sizes and dates)
  <engine>.fingerprint   sha256(weights) + runtime ABI + options + GPU name (core/common.py:92-117)
  engine_staleness()     same decision table as core/common.py:120-138

Everything core/common_runtime.py exports is re-exported, as core/common.py:43 does.
"""
from __future__ import annotations

import json
import os
import time
from typing import Optional, Sequence, Tuple

from . import _lib, weights as W
from .common_runtime import *          # noqa: F401,F403  (mirrors `from core.common_runtime import *`)
from .common_runtime import allocate_buffers, do_inference, free_buffers  # noqa: F401  explicit for linters
from .engine import Engine, ExecutionContext, TensorIOMode, make_desc   # noqa: F401


def GiB(val):
    return val * 1 << 30


def _gpu_name() -> Optional[str]:
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.get_device_name(0)
    except Exception:
        pass
    return None


def _engine_fingerprint(model_file_path, precision, workspace_gib, opt_level,
                        obey_precision_constraints, dynamic_input_shapes, batch=1, input_mode="f32_nchw", extra=None):
    """What an engine was built from, so a stale stub is not trusted.  `extra`: every further option that changes what the
    engine computes or which kernels it launches (output, max_src_hw, swap_rb, world, gather, split_k, pdl, graph, attn_poly)."""
    parts = [
        W.file_sha256(model_file_path),
        f"mde_b200_abi={_lib.load().mde_abi_version()}",
        f"precision={precision}",
        f"workspace={workspace_gib}",
        f"opt_level={opt_level}",
        f"obey_precision={obey_precision_constraints}",
        f"dynamic={dynamic_input_shapes}",
        f"batch={batch}",
        f"input_mode={input_mode}",
    ]
    for k in sorted(extra or {}):
        parts.append(f"{k}={extra[k]}")
    gpu = _gpu_name()
    if gpu:
        parts.append(f"gpu={gpu}")
    return "\n".join(parts)


def engine_staleness(engine_file_path, fingerprint_path, fingerprint, model_exists):
    """Why the recorded engine cannot be trusted, or None if it can (core/common.py:120-138)."""
    if not os.path.exists(engine_file_path):
        return "no engine file"
    if not model_exists or fingerprint is None:
        return None
    if not os.path.exists(fingerprint_path):
        return "no fingerprint recorded"
    with open(fingerprint_path, encoding="utf-8") as f:
        return None if f.read() == fingerprint else "weights or build options changed"


def get_engine(
    onnx_file_path,
    engine_file_path="",
    precision="fp16",
    dynamic_input_shapes=None,
    workspace_gib=2,
    opt_level=None,
    obey_precision_constraints=False,
    check_fingerprint=True,
    *,
    batch: int = 1,
    input_mode: str = "f32_nchw",
    max_src_hw: Tuple[int, int] = (0, 0),
    swap_rb: bool = True,
    device: int = 0,
    output: str = "model_grid",
    world: int = 1,
    rank: int = 0,
    gather: str = "fused",
    split_k: bool = False,
    pdl: bool = True,
    graph: bool = True,
    attn_poly: int = -1,
):
    """Build the engine for the exported model at `onnx_file_path` (an .mdew file here).

    Positional arguments keep the reference's order and meaning.  `precision` is "fp16" (the
    reference's build target) or "bf16"; "fp32" is refused -- the B200 path is a 16-bit
    tensor-core path with fp32 accumulation and will not silently run something else.
    `dynamic_input_shapes` must be None: engines are static, as every engine the reference ships
    (models/depth_anything_v2/onnx2trt.py:67 "dynamic = False  # fail...(False only)").
    `workspace_gib`, `opt_level` and `obey_precision_constraints` only enter the fingerprint.
    Keyword-only extras: `batch` (images per execute), `input_mode` ("f32_nchw" = the reference's
    float32 contract fed by core/preprocess.py; "u8_hwc" = raw source frames, preprocessing fused
    on the GPU), `max_src_hw` for the latter; `output` ("model_grid" = spec.json's [B, H, W] map; "source_grid" =
    the scripts' post-processing fused in: resized back to `max_src_hw` / the bound source size and clamped,
    models/depth_anything_v2/onnx2trt.py:111-117); `split_k` / `pdl` / `graph` / `attn_poly` = the engine's tuning surface
    (mde_engine_desc.flags, attn_poly).  Every keyword enters the fingerprint.
    """
    extra = dict(output=output, max_src_hw=tuple(max_src_hw), swap_rb=bool(swap_rb), world=world, gather=gather,
                 split_k=bool(split_k), pdl=bool(pdl), graph=bool(graph), attn_poly=int(attn_poly))
    model_path = os.fspath(onnx_file_path)
    if not os.path.exists(model_path):
        raise FileNotFoundError(f"[MDET] model file {model_path} not found.")
    if dynamic_input_shapes is not None:
        raise ValueError("[MDET] dynamic input shapes are not supported: engines are static")

    begin = time.time()
    meta = W.read_meta(model_path)
    if meta.get("family") == "depth_pro":
        # models/depth_pro/onnx2trt.py:99: the same call, an engine with the same three bindings.  `world` / `rank` shard
        # the 35 crops over one process per GPU (gather = "fused" | "nccl"); batch and input mode are fixed by the model.
        from .depth_pro import DepthProEngine
        if batch != 1 or input_mode != "f32_nchw" or output != "model_grid":
            raise ValueError("[MDET] the Depth Pro engine is batch 1 with the float32 NCHW input of its spec.json")
        print(f"[MDET] Build engine ({engine_file_path or model_path})")
        sd, _ = W.load(model_path)
        engine = DepthProEngine(sd, encoder=meta["encoder"], features=meta["features"], precision=precision,
                                hook_blocks=tuple(meta["hook_blocks"]), image_size=meta["input_h"], world=world, rank=rank,
                                gather=gather, device=device)
        for i in range(engine.num_io_tensors):
            name = engine.get_tensor_name(i)
            kind = "input" if engine.get_tensor_mode(name) == TensorIOMode.INPUT else "output"
            print(f"[MDET] {kind}({i}) name: {name}, shape= {engine.get_tensor_shape(name)}")
        if engine_file_path:
            os.makedirs(os.path.dirname(engine_file_path) or ".", exist_ok=True)
            with open(engine_file_path, "w", encoding="utf-8") as f:
                json.dump({"backend": "mde_b200", "abi": _lib.load().mde_abi_version(), "meta": meta, "precision": precision,
                           "batch": 1, "input_mode": input_mode, "world": world}, f, indent=2)
            if check_fingerprint:
                with open(os.path.splitext(engine_file_path)[0] + ".fingerprint", "w", encoding="utf-8") as f:
                    f.write(_engine_fingerprint(model_path, precision, workspace_gib, opt_level, obey_precision_constraints,
                                                dynamic_input_shapes, 1, input_mode, extra))
        print(f"[MDET] Engine build done! ({time.time() - begin:.2f} [sec])")
        return engine
    if meta.get("family") in ("vggt", "streamvggt"):
        # models/vggt/onnx2trt.py / models/streamvggt/onnx2trt.py: bindings `images` [1, S, 3, H, W] -> `depth` [1, S, H, W, 1].
        # `world` / `rank` shard the frames of a scene (VGGT only); a StreamVGGT file written with stream_frames > 0 builds the
        # streaming engine (one frame per execute against cached keys / values, `context.reset_stream()` between streams).
        from .vggt import VGGTEngine
        if batch != 1 or input_mode != "f32_nchw" or output != "model_grid":
            raise ValueError("[MDET] the VGGT engines take one scene of float32 frames, as their spec.json says")
        print(f"[MDET] Build engine ({engine_file_path or model_path})")
        sd, _ = W.load(model_path)
        engine = VGGTEngine(sd, encoder=meta["encoder"], depth=meta["aggregator_depth"], features=meta["features"],
                            out_channels=tuple(meta["out_channels"]), taps=tuple(meta["taps"]), frames=meta["frames"],
                            image_hw=(meta["input_h"], meta["input_w"]), precision=precision, world=world, rank=rank, gather=gather,
                            device=device, causal=meta["family"] == "streamvggt", stream_frames=meta.get("stream_frames", 0))
        for i in range(engine.num_io_tensors):
            name = engine.get_tensor_name(i)
            kind = "input" if engine.get_tensor_mode(name) == TensorIOMode.INPUT else "output"
            print(f"[MDET] {kind}({i}) name: {name}, shape= {engine.get_tensor_shape(name)}")
        if engine_file_path:
            os.makedirs(os.path.dirname(engine_file_path) or ".", exist_ok=True)
            with open(engine_file_path, "w", encoding="utf-8") as f:
                json.dump({"backend": "mde_b200", "abi": _lib.load().mde_abi_version(), "meta": meta, "precision": precision,
                           "batch": 1, "input_mode": input_mode, "world": world}, f, indent=2)
            if check_fingerprint:
                with open(os.path.splitext(engine_file_path)[0] + ".fingerprint", "w", encoding="utf-8") as f:
                    f.write(_engine_fingerprint(model_path, precision, workspace_gib, opt_level, obey_precision_constraints,
                                                dynamic_input_shapes, 1, input_mode, extra))
        print(f"[MDET] Engine build done! ({time.time() - begin:.2f} [sec])")
        return engine
    desc = make_desc(meta, precision=precision, batch=batch, input_mode=input_mode,
                     max_src_hw=max_src_hw, swap_rb=swap_rb, device=device, output=output,
                     split_k=split_k, pdl=pdl, graph=graph, attn_poly=attn_poly)

    fingerprint = None
    fingerprint_path = os.path.splitext(engine_file_path)[0] + ".fingerprint" if engine_file_path else ""
    if check_fingerprint:
        fingerprint = _engine_fingerprint(model_path, precision, workspace_gib, opt_level,
                                          obey_precision_constraints, dynamic_input_shapes, batch, input_mode, extra)
    if engine_file_path:
        stale = engine_staleness(engine_file_path, fingerprint_path, fingerprint, True)
        if stale is None:
            print(f"[MDET] Engine record is current ({engine_file_path})")
        elif os.path.exists(engine_file_path):
            print(f"[MDET] Rebuilding engine - {stale}")
    print(f"[MDET] Build engine ({engine_file_path or model_path})")

    engine = Engine(desc, meta)
    try:
        engine.load_weights_file(model_path)
        engine.finalize()
    except Exception:
        engine.close()
        raise
    for i in range(engine.num_io_tensors):
        name = engine.get_tensor_name(i)
        kind = "input" if engine.get_tensor_mode(name) == TensorIOMode.INPUT else "output"
        print(f"[MDET] {kind}({i}) name: {name}, shape= {engine.get_tensor_shape(name)}")

    if engine_file_path:
        os.makedirs(os.path.dirname(engine_file_path) or ".", exist_ok=True)
        with open(engine_file_path, "w", encoding="utf-8") as f:
            json.dump({"backend": "mde_b200", "abi": _lib.load().mde_abi_version(), "meta": meta,
                       "precision": precision, "batch": batch, "input_mode": input_mode,
                       "workspace_bytes": engine.workspace_bytes}, f, indent=2)
        if fingerprint is not None:
            with open(fingerprint_path, "w", encoding="utf-8") as f:
                f.write(fingerprint)
    dur = time.time() - begin
    print(f"[MDET] Engine build done! ({dur:.2f} [sec])")
    return engine
