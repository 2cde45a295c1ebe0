"""ORACLE (test infrastructure, never shipped, never on the product path).

numpy restatement of the reference's host preprocessing for the stretch + ImageNet family
(Depth Anything V2/V3, Distill Any Depth) and the keep-ratio + pad form of Metric3D V2 (`preprocess_pad_none`): /root/reference/core/preprocess.py:412-430
`preprocess` with MODELS['depth_anything_v2'] (:463-468):

    cvtColor BGR->RGB (:420) -> resize_stretch = cv2.resize(INTER_LINEAR) on uint8 (:144-154)
    -> scale: uint8 -> float64 / 255.0 (:294-305) -> post_resize keep_ratio: identity at the built
    size (:465-467; tests/test_preprocess.py:175-188) -> standardize (x - mean)/std in float64
    (:308-328) -> to_nchw: contiguous float32 (:337-342)

cv2's 8-bit bilinear is itself a third-party algorithm (OpenCV 4.13, modules/imgproc/src/
resize.cpp: `resizeGeneric_`/`HResizeLinear`/`VResizeLinear` with INTER_RESIZE_COEF_BITS = 11);
`resize_linear_u8` restates it in integer arithmetic and is pinned against the cv2 build in this
image by tests/test_oracle_preprocess.py (bit-exact on every shape pair tried), and the whole
pipeline is pinned against the reference module itself (imported from /root/reference where that
exists) and against tests/golden/preprocess_*.npz.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import numpy as np

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # core/preprocess.py:39-40
IMAGENET_STD = (0.229, 0.224, 0.225)
COEF_SCALE = 2048                       # 1 << INTER_RESIZE_COEF_BITS


def _coeffs(dst: int, src: int, clamp: bool):
    """Per destination index: source index and the two int16 weights.

    coordinates in float64, fraction in float32, weights rint(f * 2048) (round-half-even).
    x axis (`clamp=True`): out-of-range taps collapse onto the border pixel with weight 2048/0.
    y axis: the fraction is kept and the *row indices* are clamped by the caller instead."""
    inv = np.float64(dst) / np.float64(src)
    scale = np.float64(1.0) / inv
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp:
        lo = s < 0
        f[lo] = 0.0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0.0
        s[hi] = src - 1
    w0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int64)
    w1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int64)
    return s, w0, w1


def resize_linear_u8(img: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_LINEAR) for HxWxC uint8."""
    assert img.dtype == np.uint8 and img.ndim == 3
    src_h, src_w = img.shape[:2]
    if (src_h, src_w) == (dst_h, dst_w):
        return img.copy()
    if src_w == 2 * dst_w and src_h == 2 * dst_h:
        # OpenCV switches INTER_LINEAR to the INTER_AREA fast path when both scales are exactly 2
        a = img.astype(np.int64)
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx, a0, a1 = _coeffs(dst_w, src_w, clamp=True)
    sy, b0, b1 = _coeffs(dst_h, src_h, clamp=False)
    sx1 = np.minimum(sx + 1, src_w - 1)
    y0 = np.clip(sy, 0, src_h - 1)
    y1 = np.clip(sy + 1, 0, src_h - 1)
    src = img.astype(np.int64)
    # horizontal pass, un-shifted 11-bit products
    h = src[:, sx, :] * a0[None, :, None] + src[:, sx1, :] * a1[None, :, None]        # [src_h, dst_w, C]
    r0, r1 = h[y0], h[y1]
    out = (((b0[:, None, None] * (r0 >> 4)) >> 16) + ((b1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def norm_lut(mean=IMAGENET_MEAN, std=IMAGENET_STD) -> np.ndarray:
    """[3, 256] float32: (v / 255.0 - mean_c) / std_c evaluated in float64, rounded once."""
    v = np.arange(256, dtype=np.float64) / 255.0
    return np.stack([((v - np.float64(m)) / np.float64(s)) for m, s in zip(mean, std)]).astype(np.float32)


def preprocess_stretch_imagenet(img_bgr: np.ndarray, dst_h: int, dst_w: int, scale_dtype: str = "float64") -> np.ndarray:
    """BGR uint8 HxWx3 -> float32 [1, 3, dst_h, dst_w], byte-exact with
    core.preprocess.preprocess_for(img, 'depth_anything_v2', (dst_h, dst_w))[0]; with scale_dtype="float32" (the division by
    255 happens in float32, core/preprocess.py:294-305, :470-476 `da_ac`) with ...(img, 'depth_anything_ac', (dst_h, dst_w))[0]
    at the square bench size, where the model's keep-ratio post-resize is the identity."""
    rgb = img_bgr[:, :, ::-1]
    small = resize_linear_u8(np.ascontiguousarray(rgb), dst_h, dst_w)
    x = small.astype(np.dtype(scale_dtype)) / 255.0
    x = (x - np.asarray(IMAGENET_MEAN, np.float64)) / np.asarray(IMAGENET_STD, np.float64)
    return np.ascontiguousarray(x.transpose(2, 0, 1)[None]).astype(np.float32)


def keep_ratio_size(src_h: int, src_w: int, target: int = 518, multiple: int = 14, rounding: str = "ceil"):
    """Network input size of the keep-ratio rule (core/preprocess.py:157-171 `resize_keep_ratio`, bound="lower", with
    :112-137 `_round_to_multiple`): the short side goes to `target`, both sides snap to a multiple -- always up for
    depth_anything_ac ("ceil": 4:3 -> 518 x 700), to the nearest but not below the target for depth_anything_v2 ("constrain":
    4:3 -> 518 x 686)."""
    scale = target / min(src_h, src_w)

    def snap(x):
        if rounding == "constrain":
            y = int(np.round(x / multiple) * multiple)
            if y < target:
                y = int(np.ceil(x / multiple) * multiple)
            return max(y, multiple)
        q = x / multiple
        q = np.ceil(q) if rounding == "ceil" else np.floor(q) if rounding == "floor" else np.round(q)
        return max(int(q) * multiple, multiple)

    return snap(src_h * scale), snap(src_w * scale)


IMAGENET_PAD = (123.675, 116.28, 103.53)   # core/preprocess.py `_IMAGENET_PAD`: metric3d_v2 pads with the mean colour


def pad_geometry(src_h: int, src_w: int, dst_h: int, dst_w: int):
    """core/preprocess.py:191-219 `resize_pad` with rounding='trunc', center=True (the metric3d_v2 spec, :487-491):
    -> (inner_h, inner_w, top, left)."""
    scale = min(dst_h / src_h, dst_w / src_w)
    inner_h, inner_w = int(src_h * scale), int(src_w * scale)
    return inner_h, inner_w, (dst_h - inner_h) // 2, (dst_w - inner_w) // 2


def preprocess_pad_none(img_bgr: np.ndarray, dst_h: int, dst_w: int, pad_rgb=IMAGENET_PAD) -> np.ndarray:
    """BGR uint8 HxWx3 -> float32 [1, 3, dst_h, dst_w] in 0..255 units, byte-exact with
    core.preprocess.preprocess_for(img, 'metric3d_v2', (dst_h, dst_w))[0]: BGR->RGB, keep-ratio INTER_LINEAR resize to the
    truncated inner size, centre pad with the mean colour (cv2.copyMakeBorder saturate-casts it: round half to even),
    no normalisation (the model's own first op does it), float32."""
    rgb = np.ascontiguousarray(img_bgr[:, :, ::-1])
    inner_h, inner_w, top, left = pad_geometry(rgb.shape[0], rgb.shape[1], dst_h, dst_w)
    canvas = np.empty((dst_h, dst_w, 3), dtype=np.uint8)
    canvas[:] = np.clip(np.rint(np.asarray(pad_rgb, np.float64)), 0, 255).astype(np.uint8)
    canvas[top:top + inner_h, left:left + inner_w] = resize_linear_u8(rgb, inner_h, inner_w)
    return np.ascontiguousarray(canvas.transpose(2, 0, 1)[None]).astype(np.float32)


def im2col(x_nchw: np.ndarray, patch: int = 14, kpad: int | None = None) -> np.ndarray:
    """[B,3,H,W] -> [B*gh*gw, kpad]; row = (b, gy, gx), column = c*p*p + ky*p + kx -- the layout a
    conv with kernel == stride == patch reads, zero padded to kpad."""
    B, Cc, H, W_ = x_nchw.shape
    gh, gw = H // patch, W_ // patch
    t = x_nchw.reshape(B, Cc, gh, patch, gw, patch).transpose(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, Cc * patch * patch)
    if kpad is None or kpad == t.shape[1]:
        return np.ascontiguousarray(t)
    out = np.zeros((t.shape[0], kpad), dtype=t.dtype)
    out[:, :t.shape[1]] = t
    return out


def metric3d_postprocess(depth, src_h: int, src_w: int, size=(616, 1064), focal_px=None):
    """models/metric3d_v2/onnx2trt.py:148-158 restated with torch, as the script runs it: un-pad with the geometry of
    tools/evaluate_gt.py:133-139, F.interpolate(mode='bilinear') to the source size, clamp(0, 300); `focal_px` applies
    canonical * focal * scale / 1000 before the clamp (evaluate_gt.py:162-184)."""
    import torch
    scale = min(size[0] / src_h, size[1] / src_w)
    rh, rw = int(src_h * scale), int(src_w * scale)
    pad_h, pad_w = size[0] - rh, size[1] - rw
    pad = (pad_h // 2, pad_h - pad_h // 2, pad_w // 2, pad_w - pad_w // 2)
    d = torch.as_tensor(depth, dtype=torch.float32).reshape(size)
    d = d[pad[0]:d.shape[0] - pad[1], pad[2]:d.shape[1] - pad[3]]
    d = torch.nn.functional.interpolate(d[None, None], (src_h, src_w), mode="bilinear").squeeze()
    if focal_px is not None:
        d = d * (focal_px * scale / 1000.0)
    return torch.clamp(d, 0, 300)


# ------------------------------------------------------------------ VGGT / StreamVGGT: white square pad, then INTER_CUBIC
def _cubic_tables(ssize: int, dsize: int):
    """Source index of the second tap and the four 11-bit coefficients per destination index (cv2 `resize`, INTER_CUBIC on
    8-bit data): fx = (float)((d + 0.5) * scale - 0.5), s = floor(fx), Keys kernel with A = -0.75 evaluated in fp32 in
    `interpolateCubic`'s order, coefficients saturate_cast<short>(c * 2048) (round half to even), no re-normalisation."""
    f32 = np.float32
    A, one = f32(-0.75), f32(1)
    d = np.arange(dsize, dtype=np.float64)
    fx = ((d + 0.5) * (ssize / dsize) - 0.5).astype(f32)
    s = np.floor(fx).astype(np.int64)
    x = (fx - s.astype(f32)).astype(f32)
    c0 = ((A * (x + one) - f32(5) * A) * (x + one) + f32(8) * A) * (x + one) - f32(4) * A
    c1 = ((A + f32(2)) * x - (A + f32(3))) * x * x + one
    c2 = ((A + f32(2)) * (one - x) - (A + f32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    co = np.rint(np.stack([c0, c1, c2, c3], axis=1).astype(f32) * f32(2048)).astype(np.int64)
    return s, co


def resize_cubic_u8(img: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_CUBIC) for HxWxC uint8 -- OpenCV's OWN implementation
    (imgproc/resize.cpp: integer horizontal pass with replicated borders, vertical pass in fp32 as `VResizeCubicVec_32s8u` does
    it -- products of int rows and beta / 2^22, summed right to left without fusion, rounded half to even -- and the scalar
    integer tail `(sum + 2^21) >> 22` for the last (W * C) % 8 elements of a row).  Bit-exact with cv2 4.13 when IPP is off
    (`cv2.ipp.setUseIPP(False)`); opencv-python wheels route 8-bit cubic through Intel IPP by default, whose closed
    implementation differs from this by one level in ~4.5 % of the pixels (tests/test_oracle_preprocess.py shows both)."""
    assert img.dtype == np.uint8 and img.ndim == 3
    src_h, src_w, cn = img.shape
    if (src_h, src_w) == (dst_h, dst_w):
        return img.copy()
    xi, xa = _cubic_tables(src_w, dst_w)
    yi, ya = _cubic_tables(src_h, dst_h)
    S = img.astype(np.int64)
    H = np.zeros((src_h, dst_w, cn), np.int64)
    for k in range(4):
        H += S[:, np.clip(xi + k - 1, 0, src_w - 1), :] * xa[:, k][None, :, None]
    R = [H[np.clip(yi + k - 1, 0, src_h - 1)] for k in range(4)]
    f32 = np.float32
    scale = f32(1.0) / f32(2048 * 2048)
    b = [(ya[:, k].astype(f32) * scale)[:, None, None] for k in range(4)]
    Rf = [r.astype(f32) for r in R]
    x = Rf[0] * b[0] + (Rf[1] * b[1] + (Rf[2] * b[2] + Rf[3] * b[3]))
    out = np.clip(np.rint(x), 0, 255).astype(np.uint8)
    tail = (dst_w * cn) % 8                                     # elements past the last full 8-lane vector: integer path
    if tail:
        O = sum(R[k] * ya[:, k][:, None, None] for k in range(4))
        exact = np.clip((O + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
        flat, flat_exact = out.reshape(dst_h, dst_w * cn), exact.reshape(dst_h, dst_w * cn)
        flat[:, dst_w * cn - tail:] = flat_exact[:, dst_w * cn - tail:]
    return out


# ------------------------------------------------------------------ Depth-Anything-AC `native` profile: float32 INTER_CUBIC
def _cubic_tables_f32(ssize: int, dsize: int):
    """Source index of the second tap and the four float32 Keys coefficients (A = -0.75) per destination index, as cv2 computes
    them for 32-bit data (imgproc/resize.cpp `interpolateCubic`, no fixed-point rounding)."""
    f32 = np.float32
    A, one = f32(-0.75), f32(1)
    d = np.arange(dsize, dtype=np.float64)
    fx = ((d + 0.5) * (ssize / dsize) - 0.5).astype(f32)
    s = np.floor(fx).astype(np.int64)
    x = (fx - s.astype(f32)).astype(f32)
    c0 = ((A * (x + one) - f32(5) * A) * (x + one) + f32(8) * A) * (x + one) - f32(4) * A
    c1 = ((A + f32(2)) * x - (A + f32(3))) * x * x + one
    c2 = ((A + f32(2)) * (one - x) - (A + f32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return s, np.stack([c0, c1, c2, c3], axis=1).astype(f32)


def resize_cubic_f32(img: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h), interpolation=cv2.INTER_CUBIC) for HxWxC float32 (imgproc/resize.cpp
    `HResizeCubic<float>` then `VResizeCubic<float>`): four-tap horizontal sums with replicated borders, taps added left to
    right; four-tap vertical sums, added right to left in the 4-lane vector loop and left to right in the scalar loop that
    finishes a row ((W * C) % 4 elements); every product and sum rounded to float32, nothing fused.  Bit-exact
    with cv2 4.x when IPP is off (`cv2.ipp.setUseIPP(False)`); opencv-python wheels route float cubic through Intel IPP by
    default, whose closed implementation differs from this by up to 5e-5 on 0..1 data (tests/test_oracle_preprocess.py)."""
    assert img.dtype == np.float32 and img.ndim == 3
    src_h, src_w, _ = img.shape
    if (src_h, src_w) == (dst_h, dst_w):
        return img.copy()
    xi, xa = _cubic_tables_f32(src_w, dst_w)
    yi, ya = _cubic_tables_f32(src_h, dst_h)
    H = None
    for k in range(4):
        term = img[:, np.clip(xi + k - 1, 0, src_w - 1), :] * xa[:, k][None, :, None]
        H = term if H is None else (H + term)
    R = [H[np.clip(yi + k - 1, 0, src_h - 1)] * ya[:, k][:, None, None] for k in range(4)]
    out = (R[0] + (R[1] + (R[2] + R[3]))).astype(np.float32)          # `VResizeCubicVec_32f`: four products added right to left
    tail = (dst_w * img.shape[2]) % 4                                 # elements past the last full 4-lane vector of a row: the
    if tail:                                                          # scalar loop, which adds them left to right
        flat = out.reshape(dst_h, -1)
        ltr = (((R[0] + R[1]) + R[2]) + R[3]).astype(np.float32).reshape(dst_h, -1)
        flat[:, -tail:] = ltr[:, -tail:]
    return out


def preprocess_keep_ratio_cubic_f32(img_bgr: np.ndarray, target: int = 518, multiple: int = 14, mean=IMAGENET_MEAN,
                                    std=IMAGENET_STD) -> np.ndarray:
    """depth_anything_ac's `native` profile (core/preprocess.py:470-476 `da_ac(h, w, stretch=False)`, models/depth_anything_ac/
    onnx2trt.py:50-75): no uint8 resize; BGR -> RGB; float32 / 255; cv2 float INTER_CUBIC to the keep-ratio "ceil" size (short
    side `target`, both sides rounded up to `multiple`: 480 x 640 -> 518 x 700); (x - mean) / std in float64; float32
    [1, 3, H, W]."""
    assert img_bgr.dtype == np.uint8 and img_bgr.ndim == 3
    h, w = keep_ratio_size(img_bgr.shape[0], img_bgr.shape[1], target, multiple, "ceil")
    x = img_bgr[:, :, ::-1].astype(np.float32) / np.float32(255.0)
    x = resize_cubic_f32(np.ascontiguousarray(x), h, w)
    x = (x.astype(np.float64) - np.asarray(mean, np.float64)) / np.asarray(std, np.float64)
    return np.ascontiguousarray(x.astype(np.float32).transpose(2, 0, 1)[None])


def square_pad_geometry(src_h: int, src_w: int):
    """core/preprocess.py:222-265 `resize_square_pad` with symmetric=True: (top, left) and the padded size -- the SAME count on
    both sides, so an odd difference leaves the canvas one pixel short of square, as upstream does."""
    m = max(src_h, src_w)
    left, top = (m - src_w) // 2, (m - src_h) // 2
    return top, left, src_h + 2 * top, src_w + 2 * left


def preprocess_square_pad_cubic(img_bgr: np.ndarray, dst_h: int, dst_w: int, pad_value: int = 255) -> np.ndarray:
    """BGR uint8 HxWx3 -> float32 [1, 1, 3, dst_h, dst_w] in 0..1: core.preprocess.preprocess_for(img, 'vggt', (dst_h, dst_w))[0]
    (core/preprocess.py:493-498: BGR->RGB, white square pad at source resolution, ONE cubic resize, / 255 in float32, rank 5),
    with cv2's own cubic (see resize_cubic_u8)."""
    rgb = np.ascontiguousarray(img_bgr[:, :, ::-1])
    top, left, ph, pw = square_pad_geometry(rgb.shape[0], rgb.shape[1])
    canvas = np.full((ph, pw, 3), pad_value, dtype=np.uint8)
    canvas[top:top + rgb.shape[0], left:left + rgb.shape[1]] = rgb
    small = resize_cubic_u8(canvas, dst_h, dst_w)
    x = small.astype(np.float32) / np.float32(255.0)
    return np.ascontiguousarray(x.transpose(2, 0, 1)[None, None])


def vggt_postprocess(depth, net_h: int, net_w: int, src_h: int, src_w: int) -> np.ndarray:
    """tools/evaluate_gt.py:240-262 `_square_pad_depth` as it stands there (numpy + cv2): crop the box the source frame occupies
    inside the padded square (rounded corners of `_square_pad_geometry`, :192-208), cv2.INTER_LINEAR back to the source size on
    float64, NaN where the value is not above 1e-6."""
    import cv2
    d = np.asarray(depth).reshape(net_h, net_w).astype(np.float64)
    m = max(src_w, src_h)
    left, top = (m - src_w) // 2, (m - src_h) // 2
    s = net_w / m
    x1, y1, x2, y2 = left * s, top * s, (left + src_w) * s, (top + src_h) * s
    d = d[int(round(y1)):int(round(y2)), int(round(x1)):int(round(x2))]
    d = cv2.resize(d, (src_w, src_h), interpolation=cv2.INTER_LINEAR)
    return np.where(d > 1e-6, d, np.nan)
