"""Engine and ExecutionContext: the objects `get_engine` returns in place of tensorrt.ICudaEngine /
IExecutionContext.

They expose exactly the surface the reference's callers touch (SURVEY section 8 b):

  engine   num_io_tensors, get_tensor_name(i), get_tensor_shape(name),
           get_tensor_profile_shape(name, idx), get_tensor_dtype(name), get_tensor_mode(name),
           create_execution_context(), `with engine:`          core/common_runtime.py:136-171
  context  set_tensor_address(name, ptr), execute_async_v3(stream_handle=...),
           set_input_shape(name, shape), `with context:`       core/common_runtime.py:268-275,
                                                               models/depth_anything_v2/onnx2trt.py:93-100

All compute goes through libmde_b200.so; nothing here touches tensors.
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Mapping, Optional, Sequence, Tuple

import numpy as np

from . import _lib, weights as W


class TensorIOMode(enum.Enum):
    NONE = 0
    INPUT = 1
    OUTPUT = 2


# bf16 has no numpy dtype: such a binding is described as uint16 (the bit pattern), as TensorRT's numpy helpers do
_NP_DTYPES = {_lib.MDE_DT_F32: np.dtype(np.float32), _lib.MDE_DT_U8: np.dtype(np.uint8),
              _lib.MDE_DT_F16: np.dtype(np.float16), _lib.MDE_DT_BF16: np.dtype(np.uint16)}


def make_desc(meta: Mapping, precision: str = "fp16", batch: int = 1, input_mode: str = "f32_nchw",
              max_src_hw: Tuple[int, int] = (0, 0), swap_rb: bool = True,
              mean: Sequence[float] = (0.485, 0.456, 0.406), std: Sequence[float] = (0.229, 0.224, 0.225),
              device: int = 0, head: str = "dpt", tap_norm_mask: int = 0xF, output: str = "model_grid",
              split_k: bool = False, pdl: bool = True, graph: bool = True, attn_poly: int = -1,
              registers: int = 0, scale_dtype: str = "float64", normalise_f32: bool = False) -> _lib.EngineDesc:
    """`split_k`, `pdl`, `graph` and `attn_poly` are the engine's tuning surface (mde_engine_desc.flags / attn_poly): they are
    part of the description -- and of the fingerprint `get_engine` records -- not environment variables."""
    if precision not in _lib.PRECISIONS:
        # The reference also builds "fp32" engines (core/common.py:141-150).  The B200 path is a
        # 16-bit tensor-core path with fp32 accumulation; refuse instead of silently downgrading.
        raise ValueError(f"[MDET] precision {precision!r} is not supported; use one of {sorted(_lib.PRECISIONS)}")
    if input_mode not in ("f32_nchw", "u8_hwc"):
        raise ValueError(f"[MDET] unknown input_mode {input_mode!r}")
    d = _lib.EngineDesc()
    d.struct_size = C.sizeof(_lib.EngineDesc)
    d.embed_dim, d.depth, d.num_heads = meta["embed_dim"], meta["depth"], meta["num_heads"]
    d.patch_size, d.features = meta["patch_size"], meta["features"]
    for i in range(4):
        d.out_channels[i] = meta["out_channels"][i]
        d.taps[i] = meta["taps"][i]
    d.input_h, d.input_w, d.batch = meta["input_h"], meta["input_w"], int(batch)
    d.precision = _lib.PRECISIONS[precision]
    d.input_mode = _lib.MDE_INPUT_U8_HWC if input_mode == "u8_hwc" else _lib.MDE_INPUT_F32_NCHW
    d.max_src_h, d.max_src_w = int(max_src_hw[0]), int(max_src_hw[1])
    d.swap_rb = 1 if swap_rb else 0
    for i in range(3):
        d.norm_mean[i], d.norm_std[i] = float(mean[i]), float(std[i])
    d.max_depth = float(meta["max_depth"]) if meta.get("max_depth") else 0.0
    d.device = int(device)
    heads = {"dpt": _lib.MDE_HEAD_DPT, "encoder_taps": _lib.MDE_HEAD_ENCODER_TAPS, "dpt_exp_sky": _lib.MDE_HEAD_DPT_EXP_SKY}
    if head not in heads:
        raise ValueError(f"[MDET] unknown head {head!r}")
    d.head_mode = heads[head]
    d.tap_norm_mask = int(tap_norm_mask)
    if output not in ("model_grid", "source_grid"):
        raise ValueError(f"[MDET] unknown output {output!r}")
    d.output_mode = _lib.MDE_OUTPUT_SOURCE_GRID if output == "source_grid" else _lib.MDE_OUTPUT_MODEL_GRID
    if scale_dtype not in ("float64", "float32"):
        raise ValueError(f"[MDET] scale_dtype {scale_dtype!r}: float64 (depth_anything_v2) or float32 (depth_anything_ac)")
    d.flags = ((_lib.MDE_FLAG_SPLIT_K if split_k else 0) | (0 if pdl else _lib.MDE_FLAG_NO_PDL) |
               (0 if graph else _lib.MDE_FLAG_NO_GRAPH) | (_lib.MDE_FLAG_SCALE_F32 if scale_dtype == "float32" else 0) |
               (_lib.MDE_FLAG_NORMALISE_F32 if normalise_f32 else 0))
    if normalise_f32 and input_mode != "f32_nchw":
        raise ValueError("[MDET] normalise_f32 applies to the float32 input binding (the uint8 binding normalises through its table)")
    d.attn_poly = int(attn_poly)
    d.num_registers = int(registers if registers else meta.get("registers", 0))
    return d


class ExecutionContext:
    def __init__(self, engine: "Engine"):
        self._lib = _lib.load()
        self._engine = engine
        h = C.c_void_p()
        _lib.check(self._lib.mde_context_create(engine._h, C.byref(h)), "mde_context_create")
        self._h = h

    # -- IExecutionContext surface
    def set_tensor_address(self, name: str, ptr: int) -> bool:
        _lib.check(self._lib.mde_context_set_tensor_address(self._h, name.encode(), C.c_void_p(int(ptr))),
                   "set_tensor_address")
        return True

    def set_input_shape(self, name: str, shape: Sequence[int]) -> bool:
        dims = (C.c_int64 * len(shape))(*[int(s) for s in shape])
        _lib.check(self._lib.mde_context_set_input_shape(self._h, name.encode(), len(shape), dims), "set_input_shape")
        return True

    def set_gather(self, n_ranks: int, rank: int, peer_outputs: Sequence[int]) -> None:
        """Trunk-only engines: write the taps into every rank's gather buffer (device pointers mapped into this
        process) instead of the "output" binding -- the all-gather fused into the kernel that produces the taps."""
        arr = (C.c_void_p * max(1, len(peer_outputs)))(*[C.c_void_p(int(p)) for p in peer_outputs])
        _lib.check(self._lib.mde_context_set_gather(self._h, int(n_ranks), int(rank), arr), "set_gather")

    def execute_async_v3(self, stream_handle) -> bool:
        # core/profile.py:30-58 attaches a LayerTimer with `context.profiler = timer`; TensorRT then serialises the
        # layers and calls timer.report_layer_time(name, ms) for each.  Same contract here: with a profiler attached the
        # launches run one by one between CUDA events (no graph replay) and every launch is reported under its label.
        prof = getattr(self, "profiler", None)
        if prof is not None:
            for label, ms, _, _ in self.execute_timed(stream_handle):
                prof.report_layer_time(label, ms)
            return True
        _lib.check(self._lib.mde_context_enqueue(self._h, C.c_void_p(int(stream_handle))), "enqueue")
        return True

    def layer_information(self):
        """One dict per launch of the plan, in execution order -- what core/profile.py:60-75 `inspect` collects from
        TensorRT's engine inspector (keys Name / LayerType, plus the algorithmic FLOPs and bytes of the launch)."""
        n = self.launches_per_enqueue
        out = []
        buf = C.create_string_buffer(192)
        for i in range(n):
            fl, by = C.c_double(), C.c_double()
            _lib.check(self._lib.mde_context_op_info(self._h, i, buf, 192, C.byref(fl), C.byref(by)), "op_info")
            label = buf.value.decode()
            kind = label.split(" ")[0]
            layer_type = ("Convolution" if kind.startswith("conv") else "MatrixMultiply" if kind.startswith("gemm") else
                          "Attention" if kind == "attention" else "Normalization" if kind == "layernorm" else
                          "Resize" if kind in ("bilinear", "resize_depth", "upconv_head") else "Shuffle")
            prec = "Half" if self._engine._desc.precision == _lib.PRECISIONS["fp16"] else "BFloat16"
            out.append({"Name": label, "LayerType": layer_type, "Outputs": [{"Format/Datatype": prec}],
                        "AlgorithmicFlops": fl.value, "AlgorithmicBytes": by.value})
        return out

    # -- extras (not part of the TensorRT surface)
    @property
    def launches_per_enqueue(self) -> int:
        return int(self._lib.mde_context_launches_per_enqueue(self._h))

    def snapshot_block(self, block: int) -> None:
        _lib.check(self._lib.mde_context_snapshot_block(self._h, int(block)), "snapshot_block")

    def get_buffer(self, name: str) -> Tuple[int, int, int]:
        """(device pointer, bytes, dtype: 0 fp32 / 1 16-bit) of a named intermediate."""
        p, n, dt = C.c_void_p(), C.c_int64(), C.c_int32()
        _lib.check(self._lib.mde_context_get_buffer(self._h, name.encode(), C.byref(p), C.byref(n), C.byref(dt)),
                   "get_buffer")
        return int(p.value), int(n.value), int(dt.value)

    def execute_timed(self, stream_handle):
        """Profiling: one forward with CUDA events between launches -> [(label, ms, flops, bytes)]."""
        n = self.launches_per_enqueue
        ms = (C.c_float * n)()
        _lib.check(self._lib.mde_context_enqueue_timed(self._h, C.c_void_p(int(stream_handle)), ms, n), "enqueue_timed")
        out = []
        buf = C.create_string_buffer(192)
        for i in range(n):
            fl, by = C.c_double(), C.c_double()
            _lib.check(self._lib.mde_context_op_info(self._h, i, buf, 192, C.byref(fl), C.byref(by)), "op_info")
            out.append((buf.value.decode(), float(ms[i]), fl.value, by.value))
        return out

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.mde_context_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    TensorIOMode = TensorIOMode

    def __init__(self, desc: _lib.EngineDesc, meta: Optional[Mapping] = None):
        self._lib = _lib.load()
        self._desc = desc
        self.meta = dict(meta or {})
        h = C.c_void_p()
        _lib.check(self._lib.mde_engine_create(C.byref(desc), C.byref(h)), "mde_engine_create")
        self._h = h

    # -- weights
    def set_weight(self, name: str, array) -> None:
        a = np.ascontiguousarray(array, dtype=np.float32)
        dims = (C.c_int64 * a.ndim)(*a.shape)
        _lib.check(self._lib.mde_engine_set_weight(self._h, name.encode(), a.ctypes.data_as(C.c_void_p), a.ndim, dims),
                   f"set_weight({name})")

    def load_state_dict(self, state_dict: Mapping) -> None:
        """`meta["pos_interp"]` ("dinov2" | "bilinear") picks the rule the position embedding is resized with for a grid other
        than the trained one (Metric3D V2's export uses bilinear, weights.resize_pos_embed)."""
        gh = self._desc.input_h // self._desc.patch_size
        gw = self._desc.input_w // self._desc.patch_size
        for k, v in state_dict.items():
            a = v.detach().cpu().float().numpy() if hasattr(v, "detach") else np.asarray(v, dtype=np.float32)
            if k == "pretrained.pos_embed":
                a = W.resize_pos_embed(a, gh, gw, self.meta.get("pos_interp", "dinov2"))
            self.set_weight(k, a)

    def load_weights_file(self, path: str) -> None:
        _lib.check(self._lib.mde_engine_load_weights(self._h, path.encode()), f"load_weights({path})")
        gh = self._desc.input_h // self._desc.patch_size
        gw = self._desc.input_w // self._desc.patch_size
        # the trained grid is whatever the stored table holds (37 x 37 for the /14 checkpoints, 24 x 24 for ViT/16 at 384):
        # only that one tensor is read again, and only when this engine's grid differs from it
        pos = W.load_tensor(path, "pretrained.pos_embed")
        if pos.shape[1] != gh * gw + 1 or gh != gw:
            self.set_weight("pretrained.pos_embed", W.resize_pos_embed(pos, gh, gw, self.meta.get("pos_interp", "dinov2")))

    def finalize(self) -> "Engine":
        _lib.check(self._lib.mde_engine_finalize(self._h), "mde_engine_finalize")
        return self

    # -- ICudaEngine surface
    @property
    def num_io_tensors(self) -> int:
        return int(self._lib.mde_engine_num_io(self._h))

    def get_tensor_name(self, i: int) -> str:
        n = self._lib.mde_engine_io_name(self._h, int(i))
        if n is None:
            raise IndexError(f"engine has no I/O tensor {i}")
        return n.decode()

    def _index(self, name: str) -> int:
        for i in range(self.num_io_tensors):
            if self.get_tensor_name(i) == name:
                return i
        raise KeyError(f"engine has no tensor named {name!r}")

    def get_tensor_shape(self, name: str) -> Tuple[int, ...]:
        nd = C.c_int32()
        dims = (C.c_int64 * 8)()
        _lib.check(self._lib.mde_engine_io_shape(self._h, self._index(name), C.byref(nd), dims), "io_shape")
        return tuple(int(dims[i]) for i in range(nd.value))

    def get_tensor_profile_shape(self, name: str, profile_idx: int):
        s = self.get_tensor_shape(name)      # static engine: min == opt == max
        return (s, s, s)

    def get_tensor_dtype(self, name: str) -> np.dtype:
        return _NP_DTYPES[self._lib.mde_engine_io_dtype(self._h, self._index(name))]

    def get_tensor_mode(self, name: str) -> TensorIOMode:
        return TensorIOMode.INPUT if self._lib.mde_engine_io_is_input(self._h, self._index(name)) == 1 else TensorIOMode.OUTPUT

    def create_execution_context(self) -> ExecutionContext:
        return ExecutionContext(self)

    @property
    def workspace_bytes(self) -> int:
        return int(self._lib.mde_engine_workspace_bytes(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.mde_engine_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
