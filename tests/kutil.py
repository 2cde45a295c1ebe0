"""Call the single-kernel C-ABI entry points (mde_k_*) on torch CUDA tensors.  Test helper only:
torch provides device memory and the current stream; every kernel is the library's own."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from monocular_depth_estimation_trt_b200 import _lib

TORCH_DT = {"fp16": torch.float16, "bf16": torch.bfloat16}


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def epilogue(bias=None, gamma=None, act=0, x=None, accumulate_x=False, res1=None, res2=None, out=None,
             out_relu=None, ld_out=0, tokens=0, pos=None, shuffle=None, head_w=None, head_b=0.0,
             head_scale=0.0, head_out=None, gather=None):
    ep = _lib.Epilogue()
    ep.d_bias, ep.d_gamma, ep.act = ptr(bias), ptr(gamma), act
    ep.d_x, ep.accumulate_x = ptr(x), int(accumulate_x)
    ep.d_res1, ep.d_res2, ep.d_out, ep.d_out_relu = ptr(res1), ptr(res2), ptr(out), ptr(out_relu)
    ep.ld_out, ep.tokens, ep.d_pos = ld_out, tokens, ptr(pos)
    if shuffle:
        ep.shuffle_s, ep.shuffle_cout, ep.shuffle_h, ep.shuffle_w = shuffle
    ep.d_head_w, ep.head_b, ep.head_scale, ep.d_head_out = ptr(head_w), head_b, head_scale, ptr(head_out)
    if gather:       # (col0, ld, [device pointers, already offset to this rank's first row])
        ep.gather_col0, ep.gather_ld, ptrs = gather
        ep.gather_n = len(ptrs)
        for i, p_ in enumerate(ptrs):
            ep.d_gather[i] = int(p_)
    return ep


def gemm(precision, a, b, ep, m=None, k=None, n=None):
    """a [M, lda], b [N, ldb] 16-bit row-major."""
    lib = _lib.load()
    m = a.shape[0] if m is None else m
    k = a.shape[1] if k is None else k
    n = b.shape[0] if n is None else n
    _lib.check(lib.mde_k_gemm(_lib.PRECISIONS[precision], ptr(a), m, k, a.stride(0), ptr(b), n, b.stride(0),
                              C.byref(ep), stream()), "mde_k_gemm")


def conv3x3(precision, x_nhwc, w_packed, cout, ep):
    lib = _lib.load()
    B, H, W_, Cin = x_nhwc.shape
    _lib.check(lib.mde_k_conv3x3(_lib.PRECISIONS[precision], ptr(x_nhwc), B, H, W_, Cin, ptr(w_packed), cout,
                                 C.byref(ep), stream()), "mde_k_conv3x3")


def pack_conv3x3(w, dtype):
    """[cout, cin, 3, 3] fp32 -> [cout, 9*cin_pad] with K index (ky*3+kx)*cin_pad + c."""
    cout, cin = w.shape[:2]
    cin_pad = (cin + 63) // 64 * 64
    p = torch.zeros(cout, 9, cin_pad, dtype=dtype, device=w.device)
    p[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, 9, cin).to(dtype)
    return p.reshape(cout, 9 * cin_pad).contiguous()


def attention(precision, qkv, batch, ntok, heads, variant="tc"):
    """variant: "tc" (whatever the library picks), "tc:<n>" (the one-query-tile kernel with n/8 of the exponentials on the FMA
    pipe), "q3" / "q3:<n>" (the three-query-tile persistent kernel), "mma" (the independent mma.sync cross-check)."""
    lib = _lib.load()
    out = torch.empty(batch * ntok, heads * 64, dtype=qkv.dtype, device=qkv.device)
    if variant.startswith("q3"):
        poly = int(variant[3:]) if variant.startswith("q3:") else -1
        _lib.check(lib.mde_k_attention_q3(_lib.PRECISIONS[precision], ptr(qkv), ptr(out), batch, ntok, heads, poly, stream()),
                   "mde_k_attention_q3")
        return out
    if variant.startswith("tc:"):
        _lib.check(lib.mde_k_attention_poly(_lib.PRECISIONS[precision], ptr(qkv), ptr(out), batch, ntok, heads, int(variant[3:]), stream()),
                   "mde_k_attention_poly")
        return out
    fn = {"tc": lib.mde_k_attention, "mma": lib.mde_k_attention_mma}[variant]
    _lib.check(fn(_lib.PRECISIONS[precision], ptr(qkv), ptr(out), batch, ntok, heads, stream()), "mde_k_attention")
    return out


def attention_kv(precision, q, ldq, kv, ldkv, k_col0, v_col0, batch, ntok_q, ntok_kv, heads):
    """q: tensor or device pointer of [batch*ntok_q, ldq]; kv: tensor or device pointer of [batch*ntok_kv, ldkv]."""
    lib = _lib.load()
    dt = q.dtype if hasattr(q, "dtype") else TORCH_DT[precision]
    out = torch.empty(batch * ntok_q, heads * 64, dtype=dt, device="cuda")
    qp = ptr(q) if hasattr(q, "data_ptr") else int(q)
    kp = ptr(kv) if hasattr(kv, "data_ptr") else int(kv)
    _lib.check(lib.mde_k_attention_kv(_lib.PRECISIONS[precision], qp, ldq, kp, ldkv, k_col0, v_col0, ptr(out), batch, ntok_q,
                                      ntok_kv, heads, stream()), "mde_k_attention_kv")
    return out


def layernorm(precision, x, w, b, eps=1e-6, drop_cls=False, ntok=0):
    lib = _lib.load()
    rows, dim = x.shape
    orows = rows - rows // ntok if drop_cls else rows
    out = torch.empty(orows, dim, dtype=TORCH_DT[precision], device=x.device)
    _lib.check(lib.mde_k_layernorm(_lib.PRECISIONS[precision], ptr(x), ptr(w), ptr(b), ptr(out), rows, dim, eps,
                                   int(drop_cls), ntok, stream()), "mde_k_layernorm")
    return out


def bilinear(precision, x_nhwc, ho, wo):
    lib = _lib.load()
    B, H, W_, Cc = x_nhwc.shape
    out = torch.empty(B, ho, wo, Cc, dtype=x_nhwc.dtype, device=x_nhwc.device)
    _lib.check(lib.mde_k_bilinear(_lib.PRECISIONS[precision], ptr(x_nhwc), ptr(out), B, H, W_, ho, wo, Cc, stream()),
               "mde_k_bilinear")
    return out


def im2col_s2(precision, x_nhwc):
    lib = _lib.load()
    B, H, W_, Cc = x_nhwc.shape
    ho, wo = (H - 1) // 2 + 1, (W_ - 1) // 2 + 1
    out = torch.empty(B * ho * wo, 9 * Cc, dtype=x_nhwc.dtype, device=x_nhwc.device)
    _lib.check(lib.mde_k_im2col_s2(_lib.PRECISIONS[precision], ptr(x_nhwc), ptr(out), B, H, W_, Cc, stream()),
               "mde_k_im2col_s2")
    return out


def upconv_head(precision, z_nhwc, ho, wo, bias, head_w, head_b, head_scale):
    """z [B, hs, ws, ldz] 16-bit (first 288 channels live) -> depth [B, ho, wo] fp32."""
    lib = _lib.load()
    B, hs, ws, ldz = z_nhwc.shape
    out = torch.full((B, ho, wo), float("nan"), dtype=torch.float32, device=z_nhwc.device)
    _lib.check(lib.mde_k_upconv_head(_lib.PRECISIONS[precision], ptr(z_nhwc), ldz, B, hs, ws, ho, wo, ptr(bias), ptr(head_w),
                                     float(head_b), float(head_scale), ptr(out), stream()), "mde_k_upconv_head")
    return out


def resize_depth(depth, ho, wo, lo=1e-3, hi=1e3):
    lib = _lib.load()
    B, h, w = depth.shape
    out = torch.full((B, ho, wo), float("nan"), dtype=torch.float32, device=depth.device)
    _lib.check(lib.mde_k_resize_depth(ptr(depth), B, h, w, ptr(out), ho, wo, float(lo), float(hi), stream()), "mde_k_resize_depth")
    return out


def im2col_f32(precision, x_nchw, patch, kpad):
    lib = _lib.load()
    B, _, H, W_ = x_nchw.shape
    rows = B * (H // patch) * (W_ // patch)
    out = torch.empty(rows, kpad, dtype=TORCH_DT[precision], device=x_nchw.device)
    _lib.check(lib.mde_k_im2col_f32(_lib.PRECISIONS[precision], ptr(x_nchw), B, H, W_, patch, kpad, ptr(out), stream()),
               "mde_k_im2col_f32")
    return out


def preprocess_u8(precision, src_u8, dst_h, dst_w, patch=14, kpad=640, swap_rb=True,
                  mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), want_cols=True, want_nchw=True):
    """src_u8 [B, H, W, 3] uint8 cuda tensor -> (cols or None, nchw or None)."""
    lib = _lib.load()
    B, H, W_, _ = src_u8.shape
    rows = B * (dst_h // patch) * (dst_w // patch)
    cols = torch.empty(rows, kpad, dtype=TORCH_DT[precision], device=src_u8.device) if want_cols else None
    nchw = torch.empty(B, 3, dst_h, dst_w, dtype=torch.float32, device=src_u8.device) if want_nchw else None
    m3 = (C.c_double * 3)(*mean)
    s3 = (C.c_double * 3)(*std)
    _lib.check(lib.mde_k_preprocess_u8(_lib.PRECISIONS[precision], ptr(src_u8), B, H, W_, dst_h, dst_w, patch, kpad,
                                       int(swap_rb), m3, s3, ptr(cols), ptr(nchw), stream()), "mde_k_preprocess_u8")
    return cols, nchw


def preprocess_u8_pad(precision, src_u8, dst_h, dst_w, patch=14, kpad=640, swap_rb=True, pad_rgb=(123.675, 116.28, 103.53),
                      mean=None, std=None, want_cols=True, want_nchw=True):
    """Keep-ratio + centre pad (Metric3D V2's input contract); mean/std None = no normalisation."""
    lib = _lib.load()
    B, H, W_, _ = src_u8.shape
    rows = B * (dst_h // patch) * (dst_w // patch)
    cols = torch.empty(rows, kpad, dtype=TORCH_DT[precision], device=src_u8.device) if want_cols else None
    nchw = torch.full((B, 3, dst_h, dst_w), float("nan"), dtype=torch.float32, device=src_u8.device) if want_nchw else None
    p3 = (C.c_double * 3)(*pad_rgb)
    m3 = (C.c_double * 3)(*mean) if mean is not None else None
    s3 = (C.c_double * 3)(*std) if std is not None else None
    _lib.check(lib.mde_k_preprocess_u8_pad(_lib.PRECISIONS[precision], ptr(src_u8), B, H, W_, dst_h, dst_w, patch, kpad,
                                           int(swap_rb), p3, m3, s3, ptr(cols), ptr(nchw), stream()), "mde_k_preprocess_u8_pad")
    return cols, nchw


def rel_err(got, ref):
    """The larger of two error measures: the global one (max |d| / max |ref|) and the per-element one
    (max_i |d_i| / (|ref_i| + rms(ref))), which a localised error in a small-magnitude output cannot hide behind -- an
    element is held to its own magnitude plus one rms of the tensor (the floor that fp32 accumulation-order noise of a
    cancelling sum needs)."""
    got, ref = got.float(), ref.float()
    d = (got - ref).abs()
    glob = d.max() / ref.abs().max().clamp_min(1e-30)
    rms = (ref.double() ** 2).mean().sqrt().float().clamp_min(1e-30)
    elem = (d / (ref.abs() + rms)).max()
    return float(torch.maximum(glob, elem))


def rms_rel(got, ref):
    got, ref = got.double(), ref.double()
    return float(((got - ref) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt().clamp_min(1e-30))


def preprocess_u8_square_pad_cubic(src_u8, dst_h, dst_w, swap_rb=True, pad_value=255):
    """VGGT's input contract: src_u8 [B, H, W, 3] uint8 cuda tensor -> float32 [B, 3, dst_h, dst_w]."""
    lib = _lib.load()
    B, H, W_, _ = src_u8.shape
    out = torch.full((B, 3, dst_h, dst_w), float("nan"), dtype=torch.float32, device=src_u8.device)
    _lib.check(lib.mde_k_preprocess_u8_square_pad_cubic(ptr(src_u8), B, H, W_, dst_h, dst_w, int(swap_rb), pad_value, ptr(out), stream()),
               "mde_k_preprocess_u8_square_pad_cubic")
    return out


def preprocess_u8_cubic_f32(src_u8, dst_h, dst_w, swap_rb=True, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """Depth-Anything-AC's `native` profile: src_u8 [B, H, W, 3] uint8 cuda tensor -> float32 [B, 3, dst_h, dst_w]."""
    lib = _lib.load()
    B, H, W_, _ = src_u8.shape
    out = torch.full((B, 3, dst_h, dst_w), float("nan"), dtype=torch.float32, device=src_u8.device)
    m3 = (C.c_double * 3)(*mean) if mean is not None else None
    s3 = (C.c_double * 3)(*std) if std is not None else None
    _lib.check(lib.mde_k_preprocess_u8_cubic_f32(ptr(src_u8), B, H, W_, dst_h, dst_w, int(swap_rb), m3, s3, ptr(out), stream()),
               "mde_k_preprocess_u8_cubic_f32")
    return out
