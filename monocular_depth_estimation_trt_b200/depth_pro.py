"""Depth Pro on the B200 kernels: the engine the reference builds from `depth_pro_1536x1536.onnx`
(models/depth_pro/onnx_export.py:15-59, spec.json: input float32 [1,3,1536,1536]; outputs "canonical_inverse_depth"
[1,1,1536,1536] and "fov_deg" [1]) and runs through core/common_runtime.py in models/depth_pro/onnx2trt.py:99-116.

`DepthProEngine` / `DepthProContext` expose the same slice of tensorrt.ICudaEngine / IExecutionContext as `engine.Engine`,
so `common.allocate_buffers` / `common.do_inference` drive it unchanged.  What runs underneath:

  input -> crop pyramid (one kernel: 25 + 9 + 1 crops of 384 x 384, the two lower levels resized on the fly)
        -> patch trunk   (ViT-L/16 trunk-only engine, crops sharded over the ranks, taps all-gathered by the producing kernel)
        -> image trunk, field-of-view trunk (same engine type, batch 1, on the quarter-resolution crop)
        -> patch merge (five maps) -> upsampling neck (1x1 GEMMs, ConvTranspose2d as GEMM + pixel-shuffle epilogue)
        -> fusion decoder (implicit-GEMM 3x3 convolutions, residual / ReLU fused in the epilogues; every 1x1 `out_conv` that
           follows a transposed convolution is folded into its weights: two linear maps, one GEMM)
        -> depth head (3x3, ConvTranspose2d, 3x3 + ReLU + 1x1 + ReLU fused) and field-of-view head (stride-2 3x3 convolutions
           as gather + GEMM, the final 6x6 convolution as a one-row GEMM)

Activations are NHWC 16-bit (fp16 or bf16, fp32 accumulation); the two outputs are fp32.  torch provides device memory
only: every launch below is a kernel of libmde_b200.so, there is no eager fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Mapping, Sequence, Tuple

import numpy as np

from . import _lib, engine as E, sharding as S, weights as W

GRID = 24            # tokens per crop side
CROP = 384           # crop side in pixels
TRUNKS = ("encoder.patch_encoder.", "encoder.image_encoder.", "fov.encoder.0.")


def tap_plan(depth: int, hook_blocks: Sequence[int]) -> Tuple[List[int], Tuple[int, int]]:
    """The trunk engine taps four blocks; Depth Pro needs the two hooked blocks and the last one.  -> (sorted tap list,
    positions of hook_blocks[0] / hook_blocks[1] in it).  (11, 5) on the 24-block trunk -> [5, 11, 22, 23], (1, 0)."""
    want = sorted(set(int(b) for b in hook_blocks) | {depth - 1})
    if len(want) != 3 or want[-1] != depth - 1 or want[0] < 0:
        raise ValueError(f"[MDET] hook blocks {tuple(hook_blocks)} must be two distinct blocks below the last one")
    filler = next(b for b in range(depth - 2, -1, -1) if b not in want)
    taps = sorted(want + [filler])
    return taps, (taps.index(int(hook_blocks[0])), taps.index(int(hook_blocks[1])))


# ------------------------------------------------------------------------------------------------ kernel calls
def _p(t) -> C.c_void_p:
    return C.c_void_p(0 if t is None else (t if isinstance(t, int) else t.data_ptr()))


class _Ops:
    """Thin callers of the mde_k_* entry points on torch CUDA tensors, all on one stream."""

    def __init__(self, precision: str):
        self.lib = _lib.load()
        self.prec = _lib.PRECISIONS[precision]
        self.stream = C.c_void_p(0)
        self.launches = 0

    def ep(self, bias=None, act=0, x=None, res1=None, res2=None, out=None, out_relu=None, ld_out=0, shuffle=None,
           head_w=None, head_b=0.0, head_out=None, gamma=None, accumulate_x=False, head_act=0):
        e = _lib.Epilogue()
        e.d_bias, e.act, e.d_x, e.accumulate_x, e.d_gamma = _p(bias), act, _p(x), int(accumulate_x), _p(gamma)
        e.d_res1, e.d_res2, e.d_out, e.d_out_relu, e.ld_out = _p(res1), _p(res2), _p(out), _p(out_relu), ld_out
        if shuffle:
            e.shuffle_s, e.shuffle_cout, e.shuffle_h, e.shuffle_w = shuffle
        e.d_head_w, e.head_b, e.head_scale, e.d_head_out = _p(head_w), head_b, 0.0, _p(head_out)
        e.head_act = int(head_act)
        return e

    def gemm(self, a, m, k, lda, b, n, ep):
        _lib.check(self.lib.mde_k_gemm(self.prec, _p(a), m, k, lda, _p(b), n, b.stride(0), C.byref(ep), self.stream), "mde_k_gemm")
        self.launches += 1

    def conv3x3(self, x, h, w, cin, wt, cout, ep, batch=1):
        _lib.check(self.lib.mde_k_conv3x3(self.prec, _p(x), batch, h, w, cin, _p(wt), cout, C.byref(ep), self.stream), "mde_k_conv3x3")
        self.launches += 1

    def im2col_s2(self, x, h, w, c, out, batch=1):
        _lib.check(self.lib.mde_k_im2col_s2(self.prec, _p(x), _p(out), batch, h, w, c, self.stream), "mde_k_im2col_s2")
        self.launches += 1

    def layernorm(self, x, w, b, out, rows, dim, eps=1e-6, drop=0, ntok=0):
        _lib.check(self.lib.mde_k_layernorm(self.prec, _p(x), _p(w), _p(b), _p(out), rows, dim, eps, drop, ntok, self.stream), "mde_k_layernorm")
        self.launches += 1

    def bilinear(self, x, out, batch, hi, wi, ho, wo, c, addend=None):
        if addend is None:
            _lib.check(self.lib.mde_k_bilinear(self.prec, _p(x), _p(out), batch, hi, wi, ho, wo, c, self.stream), "mde_k_bilinear")
        else:
            _lib.check(self.lib.mde_k_bilinear_add(self.prec, _p(x), _p(out), batch, hi, wi, ho, wo, c, _p(addend), self.stream), "mde_k_bilinear_add")
        self.launches += 1

    def assemble_tokens(self, patch, special, frames, tokens, n_special, dim, first_frame, out):
        _lib.check(self.lib.mde_k_assemble_tokens(self.prec, _p(patch), _p(special), frames, tokens, n_special, dim, first_frame, _p(out),
                                                  self.stream), "mde_k_assemble_tokens")
        self.launches += 1

    def attention(self, qkv, out, batch, ntok, heads):
        _lib.check(self.lib.mde_k_attention(self.prec, _p(qkv), _p(out), batch, ntok, heads, self.stream), "mde_k_attention")
        self.launches += 1

    def attention_kv(self, q, ldq, kv, ldkv, k_col0, v_col0, out, batch, ntok_q, ntok_kv, heads):
        _lib.check(self.lib.mde_k_attention_kv(self.prec, _p(q), ldq, _p(kv), ldkv, k_col0, v_col0, _p(out), batch, ntok_q, ntok_kv, heads,
                                               self.stream), "mde_k_attention_kv")
        self.launches += 1

    def qknorm_rope(self, qkv, rows, heads, qw, qb, kw, kb, eps, pos, cos_sin, max_pos, gather=(), gather_ld=0):
        arr = (C.c_void_p * max(1, len(gather)))(*[C.c_void_p(int(g)) for g in gather])
        _lib.check(self.lib.mde_k_qknorm_rope(self.prec, _p(qkv), rows, heads, _p(qw), _p(qb), _p(kw), _p(kb), eps, _p(pos), _p(cos_sin),
                                              max_pos, len(gather), arr, gather_ld, self.stream), "mde_k_qknorm_rope")
        self.launches += 1

    def merge(self, src, per_side, pad, dim, out):
        _lib.check(self.lib.mde_k_merge_patches(self.prec, _p(src), per_side, GRID, pad, dim, _p(out), self.stream), "mde_k_merge_patches")
        self.launches += 1

    def crops(self, src_ptr, src_u8, swap_rb, src_h, src_w, plan, out, mean=None, std=None):
        arr = (_lib.Crop * len(plan))(*[_lib.Crop(*c) for c in plan])
        m3 = (C.c_float * 3)(*mean) if mean is not None else None
        s3 = (C.c_float * 3)(*std) if std is not None else None
        _lib.check(self.lib.mde_k_resize_crops(_p(src_ptr), int(src_u8), int(swap_rb), src_h, src_w, arr, len(plan), int(out.shape[-2]),
                                               int(out.shape[-1]), m3, s3, _p(out), self.stream), "mde_k_resize_crops")
        self.launches += 1


# ------------------------------------------------------------------------------------------------ checkpoint contract
def required_tensors(encoder: str = "vitl", features: int = 256) -> Dict[str, Tuple[int, ...]]:
    """Name -> shape of every tensor the engine reads from a Depth Pro state dict, under the module tree the reference's
    export script instantiates (models/depth_pro/onnx_export.py:15-29: `encoder` with its three up-sampling stacks and the
    low-resolution fusion, `decoder`, `head`, `fov`; the three ViT trunks under timm's key names).  `export_checkpoint` checks a
    checkpoint against it before anything is built, so a wrong file fails with the list of what is missing."""
    c = W.ENCODERS[encoder]
    D, L, Fd = c["embed_dim"], c["depth"], int(features)
    s: Dict[str, Tuple[int, ...]] = {}
    for pre in TRUNKS:
        s[pre + "cls_token"] = (1, 1, D); s[pre + "pos_embed"] = (1, 1 + GRID * GRID, D)
        s[pre + "patch_embed.proj.weight"] = (D, 3, 16, 16); s[pre + "patch_embed.proj.bias"] = (D,)
        s[pre + "norm.weight"] = (D,); s[pre + "norm.bias"] = (D,)
        for i in range(L):
            b = f"{pre}blocks.{i}."
            for n in ("norm1", "norm2"):
                s[b + n + ".weight"] = (D,); s[b + n + ".bias"] = (D,)
            s[b + "attn.qkv.weight"] = (3 * D, D); s[b + "attn.qkv.bias"] = (3 * D,)
            s[b + "attn.proj.weight"] = (D, D); s[b + "attn.proj.bias"] = (D,)
            s[b + "ls1.gamma"] = (D,); s[b + "ls2.gamma"] = (D,)
            s[b + "mlp.fc1.weight"] = (4 * D, D); s[b + "mlp.fc1.bias"] = (4 * D,)
            s[b + "mlp.fc2.weight"] = (D, 4 * D); s[b + "mlp.fc2.bias"] = (D,)
    for name, mid, n_up in (("upsample_latent0", Fd, 3), ("upsample_latent1", Fd, 2), ("upsample0", D // 2, 1), ("upsample1", D, 1), ("upsample2", D, 1)):
        s[f"encoder.{name}.0.weight"] = (mid, D, 1, 1)
        for j in range(n_up):
            s[f"encoder.{name}.{j + 1}.weight"] = (mid, mid, 2, 2)
    s["encoder.upsample_lowres.weight"] = (D, D, 2, 2); s["encoder.upsample_lowres.bias"] = (D,)
    s["encoder.fuse_lowres.weight"] = (D, 2 * D, 1, 1); s["encoder.fuse_lowres.bias"] = (D,)
    dims = [Fd, Fd, D // 2, D, D]
    for i in range(1, 5):
        s[f"decoder.convs.{i}.weight"] = (Fd, dims[i], 3, 3)
    for i in range(5):
        f = f"decoder.fusions.{i}."
        for r in (("resnet1",) if i < 4 else ()) + ("resnet2",):
            for j in (1, 3):
                s[f + f"{r}.residual.{j}.weight"] = (Fd, Fd, 3, 3); s[f + f"{r}.residual.{j}.bias"] = (Fd,)
        if i > 0:
            s[f + "deconv.weight"] = (Fd, Fd, 2, 2)
        s[f + "out_conv.weight"] = (Fd, Fd, 1, 1); s[f + "out_conv.bias"] = (Fd,)
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


def export_checkpoint(checkpoint_path: str, out_path: str, encoder: str = "vitl", features: int = 256,
                      hook_blocks: Sequence[int] = (11, 5)) -> dict:
    """Stage `export` for an upstream Depth Pro checkpoint (models/depth_pro/onnx_export.py:15-22 loads
    `ml-depth-pro/checkpoints/depth_pro.pt` into the model it then exports): read the state dict, check it against
    `required_tensors`, keep those tensors and write the .mdew file `common.get_engine` builds the engine from."""
    import torch
    sd = torch.load(checkpoint_path, map_location="cpu", weights_only=True)
    if isinstance(sd, dict) and "state_dict" in sd and "head.0.weight" not in sd:
        sd = sd["state_dict"]
    sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
    need = required_tensors(encoder, features)
    missing = [k for k in need if k not in sd]
    wrong = [f"{k}: {tuple(sd[k].shape)} != {shape}" for k, shape in need.items() if k in sd and tuple(sd[k].shape) != shape]
    if missing or wrong:
        raise ValueError(f"[MDET] {checkpoint_path} is not a Depth Pro ({encoder}, {features} decoder features) checkpoint: "
                         f"{len(missing)} tensors missing (first: {missing[:3]}), {len(wrong)} with another shape (first: {wrong[:3]})")
    meta = W.describe_depth_pro(encoder, features, hook_blocks)
    meta["source_checkpoint_sha256"] = W.file_sha256(checkpoint_path)
    W.save(out_path, {k: sd[k] for k in need}, meta)
    return meta


# ------------------------------------------------------------------------------------------------ weight packing
def _t(v):
    import torch
    return v.detach().float().cpu() if hasattr(v, "detach") else torch.from_numpy(np.asarray(v, dtype=np.float32))


def pack_conv3x3(w, dtype):
    """[cout, cin, 3, 3] -> [cout, 9 * cin_pad], K index (ky*3+kx)*cin_pad + c (mde_k_conv3x3's layout)."""
    import torch
    cout, cin = w.shape[:2]
    cin_pad = (cin + 63) // 64 * 64
    p = torch.zeros(cout, 9, cin_pad)
    p[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, 9, cin)
    return p.reshape(cout, 9 * cin_pad).to(dtype).contiguous()


def pack_conv3x3_s2(w, dtype):
    """[cout, cin, 3, 3] -> [cout, 9 * cin], K index tap*cin + c (the gather of mde_k_im2col_s2)."""
    cout, cin = w.shape[:2]
    return w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).to(dtype).contiguous()


def pack_deconv2(w, dtype, then_1x1=None):
    """ConvTranspose2d(k=2, s=2) [cin, cout, 2, 2] -> [4 * cout', cin], row (ky*2+kx)*cout' + o.  `then_1x1` [cout', cout]
    is a 1x1 convolution applied right after it: the two linear maps are multiplied out in fp32 before the rounding."""
    cin, cout = w.shape[:2]
    m = w.permute(2, 3, 1, 0).reshape(4, cout, cin)                  # [tap, o, c]
    if then_1x1 is not None:
        m = (then_1x1.double() @ m.double()).float()                 # [tap, o', c]
    return m.reshape(-1, cin).to(dtype).contiguous()


class DepthProWeights:
    """Packed device weights of everything behind the trunks."""

    def __init__(self, sd: Mapping, embed_dim: int, features: int, dtype, device):
        import torch
        D, Fd = embed_dim, features
        g = lambda k: _t(sd[k])
        dev = lambda t: t.to(device)
        self.up = {}
        for name in ("upsample_latent0", "upsample_latent1", "upsample0", "upsample1", "upsample2"):
            layers = [dev(g(f"encoder.{name}.0.weight").flatten(1).to(dtype).contiguous())]
            j = 1
            while f"encoder.{name}.{j}.weight" in sd:
                layers.append(dev(pack_deconv2(g(f"encoder.{name}.{j}.weight"), dtype)))
                j += 1
            self.up[name] = layers
        self.up_lowres = dev(pack_deconv2(g("encoder.upsample_lowres.weight"), dtype))
        self.up_lowres_b = dev(g("encoder.upsample_lowres.bias").contiguous())
        self.fuse = dev(g("encoder.fuse_lowres.weight").flatten(1).to(dtype).contiguous())
        self.fuse_b = dev(g("encoder.fuse_lowres.bias").contiguous())
        self.convs = {i: dev(pack_conv3x3(g(f"decoder.convs.{i}.weight"), dtype)) for i in range(1, 5)}
        self.fus = {}
        for i in range(5):
            f = f"decoder.fusions.{i}."
            d = {}
            for r in (("resnet1",) if i < 4 else ()) + ("resnet2",):
                for j in (1, 3):
                    d[f"{r}.{j}.w"] = dev(pack_conv3x3(g(f + f"{r}.residual.{j}.weight"), dtype))
                    d[f"{r}.{j}.b"] = dev(g(f + f"{r}.residual.{j}.bias").contiguous())
            out_w, out_b = g(f + "out_conv.weight").flatten(1), g(f + "out_conv.bias")
            if i > 0:
                d["deconv.w"] = dev(pack_deconv2(g(f + "deconv.weight"), dtype, then_1x1=out_w))
            else:
                d["out.w"] = dev(out_w.to(dtype).contiguous())
            d["out.b"] = dev(out_b.contiguous())
            self.fus[i] = d
        self.h0 = dev(pack_conv3x3(g("head.0.weight"), dtype)); self.h0_b = dev(g("head.0.bias").contiguous())
        self.h1 = dev(pack_deconv2(g("head.1.weight"), dtype)); self.h1_b = dev(g("head.1.bias").contiguous())
        self.h2 = dev(pack_conv3x3(g("head.2.weight"), dtype)); self.h2_b = dev(g("head.2.bias").contiguous())
        self.h4 = dev(g("head.4.weight").flatten().contiguous()); self.h4_b = float(g("head.4.bias"))
        self.fov_lin = dev(g("fov.encoder.1.weight").to(dtype).contiguous()); self.fov_lin_b = dev(g("fov.encoder.1.bias").contiguous())
        self.fov_down = dev(pack_conv3x3_s2(g("fov.downsample.0.weight"), dtype)); self.fov_down_b = dev(g("fov.downsample.0.bias").contiguous())
        self.fov0 = dev(pack_conv3x3_s2(g("fov.head.0.weight"), dtype)); self.fov0_b = dev(g("fov.head.0.bias").contiguous())
        self.fov2 = dev(pack_conv3x3_s2(g("fov.head.2.weight"), dtype)); self.fov2_b = dev(g("fov.head.2.bias").contiguous())
        w4 = g("fov.head.4.weight")                                   # [1, F/8, 6, 6] -> one GEMM row, K index (y*6+x)*c + ch
        last = torch.zeros(8, w4[0].numel())
        last[0] = w4[0].permute(1, 2, 0).reshape(-1)
        self.fov4 = dev(last.to(dtype).contiguous())
        b4 = torch.zeros(8)
        b4[0] = g("fov.head.4.bias")[0]
        self.fov4_b = dev(b4)


# ------------------------------------------------------------------------------------------------ engine / context
class DepthProEngine:
    """Stands in for the tensorrt.ICudaEngine of models/depth_pro/onnx2trt.py:99 (`get_engine(...)`)."""

    TensorIOMode = E.TensorIOMode
    IO = (("input", True), ("canonical_inverse_depth", False), ("fov_deg", False))

    def __init__(self, state_dict: Mapping, encoder: str = "vitl", features: int = 256, precision: str = "fp16",
                 hook_blocks: Sequence[int] = (11, 5), image_size: int = 1536, world: int = 1, rank: int = 0,
                 gather: str = "fused", device: int = 0):
        import torch
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"[MDET] precision {precision!r} is not supported; use one of {sorted(_lib.PRECISIONS)}")
        if image_size != 4 * CROP:
            # upstream's fixed size (spec.json "caveats"): the quarter-resolution level must be exactly one crop
            raise ValueError(f"[MDET] Depth Pro runs at {4 * CROP} x {4 * CROP}, not {image_size}")
        if features % 64:
            raise ValueError(f"[MDET] decoder width {features} must be a multiple of 64")
        cfg = W.ENCODERS[encoder]
        self.encoder, self.features, self.precision = encoder, int(features), precision
        self.D, self.S, self.world, self.rank, self.gather_mode = cfg["embed_dim"], int(image_size), int(world), int(rank), gather
        self.dtype = torch.bfloat16 if precision == "bf16" else torch.float16
        self.device = torch.device("cuda", device)
        self.plan = [(side, side, y0, x0) for _, side, y0, x0 in S.pyramid_plan(self.S)]
        self.n_crops = len(self.plan)
        self.per_rank, self.bounds = S.shard_bounds(self.n_crops, self.world)
        taps, self.hook_taps = tap_plan(cfg["depth"], hook_blocks)
        meta = W.describe(encoder, CROP, CROP, None, patch_size=16)
        meta["taps"] = taps
        self.trunks: List[E.Engine] = []
        try:
            for prefix, batch in zip(TRUNKS, (self.per_rank, 1, 1)):
                sd = {"pretrained." + k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
                if not sd:
                    raise ValueError(f"[MDET] the state dict holds no '{prefix}*' tensors")
                eng = E.Engine(E.make_desc(meta, precision=precision, batch=batch, head="encoder_taps", tap_norm_mask=0x8, device=device), meta)
                self.trunks.append(eng)
                eng.load_state_dict(sd)
                eng.finalize()
            self.weights = DepthProWeights(state_dict, self.D, self.features, self.dtype, self.device)
        except Exception:
            self.close()
            raise

    # -- ICudaEngine surface (core/common_runtime.py:136-171)
    @property
    def num_io_tensors(self) -> int:
        return len(self.IO)

    def get_tensor_name(self, i: int) -> str:
        return self.IO[i][0]

    def get_tensor_shape(self, name: str) -> Tuple[int, ...]:
        return {"input": (1, 3, self.S, self.S), "canonical_inverse_depth": (1, 1, self.S, self.S), "fov_deg": (1,)}[name]

    def get_tensor_profile_shape(self, name: str, profile_idx: int):
        s = self.get_tensor_shape(name)
        return (s, s, s)

    def get_tensor_dtype(self, name: str) -> np.dtype:
        self.get_tensor_shape(name)
        return np.dtype(np.float32)

    def get_tensor_mode(self, name: str):
        return E.TensorIOMode.INPUT if dict(self.IO)[name] else E.TensorIOMode.OUTPUT

    def create_execution_context(self) -> "DepthProContext":
        return DepthProContext(self)

    def close(self) -> None:
        for t in getattr(self, "trunks", []):
            t.close()
        self.trunks = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class DepthProContext:
    """Stands in for the IExecutionContext: `set_tensor_address` for the three bindings, `execute_async_v3(stream)`.
    Every buffer is allocated here; `execute_async_v3` only launches kernels (world 1: no host synchronisation either)."""

    def __init__(self, engine: DepthProEngine):
        import torch
        e = self.e = engine
        self.ops = _Ops(e.precision)
        self.addr: Dict[str, int] = {}
        dev, dt, D, Fd = e.device, e.dtype, e.D, e.features
        n = GRID * (e.S // CROP)                                      # 96: side of the full-resolution token map
        self.n = n
        z16 = lambda *shape: torch.zeros(*shape, dtype=dt, device=dev)
        self.crops_in = torch.zeros(e.per_rank, 3, CROP, CROP, dtype=torch.float32, device=dev)
        self.low_in = torch.zeros(1, 3, CROP, CROP, dtype=torch.float32, device=dev)
        self.patch = S.ShardedPatchEncoder(e.trunks[0], n_items=e.n_crops, world=e.world, rank=e.rank, mode=e.gather_mode,
                                           external_sync=True)      # the PeerSync hand-shake below orders the rounds on the stream
        self.sync = S.PeerSync(e.world, e.rank) if (e.world > 1 and e.gather_mode == "fused") else None
        self.ctx_img, self.ctx_fov = e.trunks[1].create_execution_context(), e.trunks[2].create_execution_context()
        self.img_taps, self.fov_taps = z16(4, 1, GRID * GRID, D), z16(4, 1, GRID * GRID, D)
        for ctx, buf in ((self.ctx_img, self.img_taps), (self.ctx_fov, self.fov_taps)):
            ctx.set_tensor_address("input", self.low_in.data_ptr())
            ctx.set_tensor_address("output", buf.data_ptr())
        self.f24, self.f48, self.f96 = z16(n // 4, n // 4, D), z16(n // 2, n // 2, D), z16(n, n, D)
        self.hook_a, self.hook_b = z16(n, n, D), z16(n, n, D)
        b = {}
        side = {0: 8 * n, 1: 4 * n, 2: 2 * n, 3: n, 4: n // 2}          # decoder level -> map side (768 ... 48)
        self.side = side
        # neck
        b["lat0.p"] = z16(n * n, Fd); b["lat0.1"] = z16(4 * n * n, Fd); b["lat0.2"] = z16(16 * n * n, Fd)
        b["lat1.p"] = z16(n * n, Fd); b["lat1.1"] = z16(4 * n * n, Fd)
        b["x0.p"] = z16(n * n, D // 2); b["x1.p"] = z16(n * n // 4, D); b["x2.p"] = z16(n * n // 16, D)
        b["cat"] = z16(side[4] ** 2, 2 * D); b["xg"] = z16(side[4] ** 2, D)
        enc_dims = [Fd, Fd, D // 2, D, D]
        for i in range(5):
            if i > 0:
                b[f"enc{i}"] = z16(side[i] ** 2, enc_dims[i]) if i < 4 else b["xg"]
            b[f"proj{i}"] = z16(side[i] ** 2, Fd)
            b[f"proj{i}.relu"] = z16(side[i] ** 2, Fd)
            b[f"t{i}.relu"] = z16(side[i] ** 2, Fd)                   # relu(conv1(.)) inside a residual unit
            b[f"s{i}"] = z16(side[i] ** 2, Fd); b[f"s{i}.relu"] = z16(side[i] ** 2, Fd)
            b[f"y{i}"] = z16(side[i] ** 2, Fd)
            if i < 4:
                b[f"feat{i}"] = z16(side[i] ** 2, Fd)                 # what the fusion below hands up
        b["features"] = z16(side[0] ** 2, Fd)
        b["h0"] = z16(side[0] ** 2, Fd // 2); b["h1"] = z16(4 * side[0] ** 2, Fd // 2)
        # field of view
        g2 = GRID * GRID
        b["fov.lin"] = z16(g2, Fd // 2); b["fov.col0"] = z16(g2, 9 * Fd); b["fov.f0"] = z16(g2, Fd // 2)
        b["fov.col1"] = z16(g2 // 4, 9 * Fd // 2); b["fov.f1"] = z16(g2 // 4, Fd // 4)
        b["fov.col2"] = z16(g2 // 16, 9 * Fd // 4); b["fov.f2"] = z16(g2 // 16, Fd // 8)
        self.fov_out = torch.zeros(1, 8, dtype=torch.float32, device=dev)
        self.b = b
        self.launches_per_enqueue = 0
        # the two batch-1 trunks fill a third of the SMs each: they run on side streams, beside each other and the patch trunk
        from .common_runtime import cuda_call, cudart
        self.side_streams = [cuda_call(cudart.cudaStreamCreateWithFlags(cudart.cudaStreamNonBlocking)) for _ in range(2)]
        self.events = [cuda_call(cudart.cudaEventCreateWithFlags(cudart.cudaEventDisableTiming)) for _ in range(3)]
        self.overlap_trunks = True
        self.profile_stages = False       # True: CUDA events at the stage boundaries of the next execute (see stage_times())
        self._marks = []
        self.use_graph = True
        self._graph_exec = None
        self._graph_key = None
        self._graph_failed = False

    # -- IExecutionContext surface
    def set_tensor_address(self, name: str, ptr: int) -> bool:
        self.e.get_tensor_shape(name)
        self.addr[name] = int(ptr)
        return True

    def set_input_shape(self, name: str, shape: Sequence[int]) -> bool:
        if tuple(int(s) for s in shape) != self.e.get_tensor_shape(name):
            raise ValueError(f"[MDET] {name}: engines are static, shape {tuple(shape)} != {self.e.get_tensor_shape(name)}")
        return True

    def execute_async_v3(self, stream_handle) -> bool:
        """One forward on `stream_handle`.  The ~585 launches (three trunks on three streams, neck, decoder, two heads) are
        recorded into ONE CUDA graph the first time they run with a given set of bindings and replayed afterwards, like the
        C-ABI engine does for its own plan (csrc/engine.cu `mde_context_enqueue`): issuing them through ctypes costs the host
        more than the device needs to run them.  Not for the NCCL baseline on several ranks, the legacy stream, or while
        `profile_stages` wants events between the stages; `use_graph = False` turns it off."""
        missing = [n for n, _ in self.e.IO if not self.addr.get(n)]
        if missing:
            raise RuntimeError(f"[MDET] execute before set_tensor_address for {missing}")
        sh = int(stream_handle)
        graphable = (self.use_graph and not self.profile_stages and sh != 0 and not self._graph_failed
                     and self.e.gather_mode == "fused")      # the NCCL baseline issues collectives / copies on torch's stream
        if not graphable:
            return self._enqueue(sh)
        from .common_runtime import cuda_call, cudart
        key = tuple(sorted(self.addr.items())) + (self.overlap_trunks,)
        if self._graph_exec is None or key != self._graph_key:
            self.release_graph()
            cuda_call(cudart.cudaStreamBeginCapture(sh, cudart.cudaStreamCaptureMode.cudaStreamCaptureModeThreadLocal))
            ok = False
            try:
                self._enqueue(sh)
                ok = True
            finally:
                err, graph = cudart.cudaStreamEndCapture(sh)
                if not ok or int(err) != 0:
                    cudart.cudaGetLastError()
                    self._graph_failed = True           # a forward that cannot be recorded runs launch by launch from now on
            if self._graph_failed:
                return self._enqueue(sh)
            self._graph_exec = cuda_call(cudart.cudaGraphInstantiate(graph, 0))
            cuda_call(cudart.cudaGraphDestroy(graph))
            self._graph_key = key
        cuda_call(cudart.cudaGraphLaunch(self._graph_exec, sh))
        return True

    def release_graph(self) -> None:
        from .common_runtime import cudart
        if getattr(self, "_graph_exec", None) is not None:
            cudart.cudaGraphExecDestroy(self._graph_exec)
        self._graph_exec = None
        self._graph_key = None

    def _enqueue(self, sh: int) -> bool:
        e, o, b, w, n = self.e, self.ops, self.b, self.e.weights, self.n
        D, Fd, side = e.D, e.features, self.side
        o.stream = C.c_void_p(sh)
        o.launches = 0
        from .common_runtime import cuda_call, cudart
        self._marks = []

        def mark(label):
            if self.profile_stages:
                ev = cuda_call(cudart.cudaEventCreate())
                cuda_call(cudart.cudaEventRecord(ev, sh))
                self._marks.append((label, ev))

        mark("start")
        # 1. crops: this rank's share of the pyramid + the quarter-resolution image for the two batch-1 trunks
        first, count = e.bounds[e.rank]
        if count > 0:
            o.crops(self.addr["input"], False, False, e.S, e.S, e.plan[first:first + count], self.crops_in[:count])
        o.crops(self.addr["input"], False, False, e.S, e.S, e.plan[-1:], self.low_in)
        mark("crops")
        # 2. trunks
        if self.overlap_trunks:
            cuda_call(cudart.cudaEventRecord(self.events[0], sh))
            for st, ctx, done in ((self.side_streams[0], self.ctx_img, self.events[1]), (self.side_streams[1], self.ctx_fov, self.events[2])):
                cuda_call(cudart.cudaStreamWaitEvent(st, self.events[0], 0))
                ctx.execute_async_v3(int(st))
                cuda_call(cudart.cudaEventRecord(done, st))
        if self.sync is not None:
            self.sync.wait_acks(sh)
        self.patch.enqueue(self.crops_in.data_ptr(), sh)
        if self.sync is not None:
            self.sync.signal_ready(sh)
        mark("patch trunk")
        if not self.overlap_trunks:
            self.ctx_img.execute_async_v3(sh)
            self.ctx_fov.execute_async_v3(sh)
        mark("image + fov trunks")
        if self.sync is not None:
            self.sync.wait_ready(sh)                                   # (the NCCL baseline is ordered on the stream by NCCL itself)
        taps = self.patch.gathered()
        # 3. merge the crops' tokens into the five maps
        fin, (ha, hb) = taps[3], e.hook_taps
        o.merge(fin[34:35], 1, 0, D, self.f24)                       # crop order: 25 full-resolution, 9 half, 1 quarter
        o.merge(fin[25:34], 3, 6, D, self.f48)
        o.merge(fin[0:25], 5, 3, D, self.f96)
        o.merge(taps[ha][0:25], 5, 3, D, self.hook_a)
        o.merge(taps[hb][0:25], 5, 3, D, self.hook_b)
        if self.sync is not None:
            self.sync.signal_acks(sh)
        mark("merge")
        # 4. neck: 1x1 projection, then ConvTranspose2d layers (GEMM with the 2x2 pixel shuffle in the epilogue)
        def project(src, rows, wt, out):
            o.gemm(src, rows, D, D, wt, wt.shape[0], o.ep(out=out, ld_out=wt.shape[0]))

        def deconv(src, h, cin, wt, out, ld_out=None, bias=None):
            cout = wt.shape[0] // 4
            o.gemm(src, h * h, cin, cin, wt, 4 * cout, o.ep(bias=bias, out=out, ld_out=ld_out or cout, shuffle=(2, cout, h, h)))

        u = w.up
        project(self.hook_b, n * n, u["upsample_latent0"][0], b["lat0.p"])
        deconv(b["lat0.p"], n, Fd, u["upsample_latent0"][1], b["lat0.1"])
        deconv(b["lat0.1"], 2 * n, Fd, u["upsample_latent0"][2], b["lat0.2"])
        # the last layer of latent0 is the decoder's level-0 input as it stands (decoder.convs.0 is the identity)
        o.gemm(b["lat0.2"], 16 * n * n, Fd, Fd, u["upsample_latent0"][3], 4 * Fd,
               o.ep(out=b["proj0"], out_relu=b["proj0.relu"], ld_out=Fd, shuffle=(2, Fd, 4 * n, 4 * n)))
        project(self.hook_a, n * n, u["upsample_latent1"][0], b["lat1.p"])
        deconv(b["lat1.p"], n, Fd, u["upsample_latent1"][1], b["lat1.1"])
        deconv(b["lat1.1"], 2 * n, Fd, u["upsample_latent1"][2], b["enc1"])
        project(self.f96, n * n, u["upsample0"][0], b["x0.p"])
        deconv(b["x0.p"], n, D // 2, u["upsample0"][1], b["enc2"])
        project(self.f48, n * n // 4, u["upsample1"][0], b["x1.p"])
        deconv(b["x1.p"], n // 2, D, u["upsample1"][1], b["enc3"])
        project(self.f24, n * n // 16, u["upsample2"][0], b["x2.p"])
        cat = b["cat"]
        if self.overlap_trunks:
            cuda_call(cudart.cudaStreamWaitEvent(sh, self.events[1], 0))      # image trunk done
        deconv(b["x2.p"], n // 4, D, u["upsample2"][1], cat, ld_out=2 * D)
        deconv(self.img_taps[3, 0], n // 4, D, w.up_lowres, cat[:, D:], ld_out=2 * D, bias=w.up_lowres_b)
        o.gemm(cat, side[4] ** 2, 2 * D, 2 * D, w.fuse, D, o.ep(bias=w.fuse_b, out=b["xg"], ld_out=D))
        mark("neck")
        # 5. decoder: project every level to `features` channels, fuse from the lowest resolution up
        enc_dims = [Fd, Fd, D // 2, D, D]
        for i in range(1, 5):
            o.conv3x3(b[f"enc{i}"], side[i], side[i], enc_dims[i], w.convs[i], Fd,
                      o.ep(out=b[f"proj{i}"], out_relu=b[f"proj{i}.relu"], ld_out=Fd))
        feat = None
        for i in (4, 3, 2, 1, 0):
            f, hw = w.fus[i], side[i]
            if feat is None:
                s, s_relu = b[f"proj{i}"], b[f"proj{i}.relu"]
            else:
                # feat + resnet1(proj): conv2's epilogue adds its own input (the unit's skip) and the running features
                o.conv3x3(b[f"proj{i}.relu"], hw, hw, Fd, f["resnet1.1.w"], Fd, o.ep(bias=f["resnet1.1.b"], out_relu=b[f"t{i}.relu"], ld_out=Fd))
                s, s_relu = b[f"s{i}"], b[f"s{i}.relu"]
                o.conv3x3(b[f"t{i}.relu"], hw, hw, Fd, f["resnet1.3.w"], Fd,
                          o.ep(bias=f["resnet1.3.b"], res1=b[f"proj{i}"], res2=feat, out=s, out_relu=s_relu, ld_out=Fd))
            o.conv3x3(s_relu, hw, hw, Fd, f["resnet2.1.w"], Fd, o.ep(bias=f["resnet2.1.b"], out_relu=b[f"t{i}.relu"], ld_out=Fd))
            o.conv3x3(b[f"t{i}.relu"], hw, hw, Fd, f["resnet2.3.w"], Fd, o.ep(bias=f["resnet2.3.b"], res1=s, out=b[f"y{i}"], ld_out=Fd))
            if i > 0:
                feat = b[f"feat{i - 1}"]
                o.gemm(b[f"y{i}"], hw * hw, Fd, Fd, f["deconv.w"], 4 * Fd, o.ep(bias=f["out.b"], out=feat, ld_out=Fd, shuffle=(2, Fd, hw, hw)))
            else:
                o.gemm(b["y0"], hw * hw, Fd, Fd, f["out.w"], Fd, o.ep(bias=f["out.b"], out=b["features"], ld_out=Fd))
        mark("decoder")
        # 6. depth head
        hw = side[0]
        o.conv3x3(b["features"], hw, hw, Fd, w.h0, Fd // 2, o.ep(bias=w.h0_b, out=b["h0"], ld_out=Fd // 2))
        o.gemm(b["h0"], hw * hw, Fd // 2, Fd // 2, w.h1, 2 * Fd, o.ep(bias=w.h1_b, out=b["h1"], ld_out=Fd // 2, shuffle=(2, Fd // 2, hw, hw)))
        o.conv3x3(b["h1"], 2 * hw, 2 * hw, Fd // 2, w.h2, 32,
                  o.ep(bias=w.h2_b, head_w=w.h4, head_b=w.h4_b, head_out=self.addr["canonical_inverse_depth"], ld_out=32))
        mark("depth head")
        # 7. field-of-view head
        g = GRID
        if self.overlap_trunks:
            cuda_call(cudart.cudaStreamWaitEvent(sh, self.events[2], 0))      # field-of-view trunk done
        o.gemm(self.fov_taps[3, 0], g * g, D, D, w.fov_lin, Fd // 2, o.ep(bias=w.fov_lin_b, out=b["fov.lin"], ld_out=Fd // 2))
        o.im2col_s2(b["proj4"], 2 * g, 2 * g, Fd, b["fov.col0"])
        o.gemm(b["fov.col0"], g * g, 9 * Fd, 9 * Fd, w.fov_down, Fd // 2, o.ep(bias=w.fov_down_b, act=2, res1=b["fov.lin"], out=b["fov.f0"], ld_out=Fd // 2))
        o.im2col_s2(b["fov.f0"], g, g, Fd // 2, b["fov.col1"])
        o.gemm(b["fov.col1"], g * g // 4, 9 * Fd // 2, 9 * Fd // 2, w.fov0, Fd // 4, o.ep(bias=w.fov0_b, act=2, out=b["fov.f1"], ld_out=Fd // 4))
        o.im2col_s2(b["fov.f1"], g // 2, g // 2, Fd // 4, b["fov.col2"])
        o.gemm(b["fov.col2"], g * g // 16, 9 * Fd // 4, 9 * Fd // 4, w.fov2, Fd // 8, o.ep(bias=w.fov2_b, act=2, out=b["fov.f2"], ld_out=Fd // 8))
        k = (g // 4) ** 2 * (Fd // 8)
        o.gemm(b["fov.f2"], 1, k, k, w.fov4, 8, o.ep(bias=w.fov4_b, x=self.fov_out, ld_out=8))
        cuda_call(cudart.cudaMemcpyAsync(self.addr["fov_deg"], self.fov_out.data_ptr(), 4,
                                         cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice, sh))
        mark("fov head")
        self.launches_per_enqueue = o.launches + self.patch.ctx.launches_per_enqueue + self.ctx_img.launches_per_enqueue + self.ctx_fov.launches_per_enqueue
        return True

    def stage_times(self):
        """[(stage, ms)] of the last execute run with `profile_stages = True` (call after the stream has synchronised)."""
        from .common_runtime import cuda_call, cudart
        out = [(self._marks[i + 1][0], float(cuda_call(cudart.cudaEventElapsedTime(self._marks[i][1], self._marks[i + 1][1]))))
               for i in range(len(self._marks) - 1)]
        for _, ev in self._marks:
            cuda_call(cudart.cudaEventDestroy(ev))
        self._marks = []
        return out

    def get_buffer(self, name: str):
        """Intermediate tensors for the per-stage parity gates (valid after execute + stream sync)."""
        return self.b[name]

    def close(self) -> None:
        self.release_graph()
        for c in (getattr(self, "ctx_img", None), getattr(self, "ctx_fov", None)):
            if c is not None:
                c.close()
        if getattr(self, "patch", None) is not None:
            self.patch.close()
            self.patch = None
        if getattr(self, "sync", None) is not None:
            self.sync.close()
            self.sync = None
        self.ctx_img = self.ctx_fov = None
        from .common_runtime import cudart
        for st in getattr(self, "side_streams", []):
            cudart.cudaStreamDestroy(st)
        for ev in getattr(self, "events", []):
            cudart.cudaEventDestroy(ev)
        self.side_streams, self.events = [], []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def postprocess(inv_ptr: int, fov_ptr: int, size: int, src_h: int, src_w: int, depth_out, f_px_out=None, stream_handle: int = 0) -> None:
    """models/depth_pro/onnx2trt.py:118-134 on the device: the engine's two outputs -> metric depth float32 [src_h, src_w]."""
    lib = _lib.load()
    _lib.check(lib.mde_k_depth_pro_post(C.c_void_p(int(inv_ptr)), C.c_void_p(int(fov_ptr)), size, size, src_h, src_w, _p(depth_out),
                                        _p(f_px_out), C.c_void_p(int(stream_handle))), "mde_k_depth_pro_post")


def preprocess_u8(src_u8, size: int, out, swap_rb: bool = False, stream_handle: int = 0) -> None:
    """models/depth_pro/onnx2trt.py:56-74 on the device: uint8 [H, W, 3] -> ToTensor -> Normalize(0.5, 0.5) ->
    interpolate(bilinear, align_corners=False) to size x size -> float32 [1, 3, size, size] (spec.json's input binding)."""
    ops = _Ops("fp16")
    ops.stream = C.c_void_p(int(stream_handle))
    h, w = int(src_u8.shape[0]), int(src_u8.shape[1])
    ops.crops(src_u8.data_ptr(), True, swap_rb, h, w, [(size, size, 0, 0)], out.reshape(1, 3, size, size), mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5))
