"""The reference scripts' post-processing as device kernels: what each `models/<name>/onnx2trt.py`
does on the CPU with torch after `do_inference`, here on the output binding where it lies.

  depth_anything_v2   resize back (align_corners=True) + clamp       -> engine output mode "source_grid" / mde_k_resize_depth
  depth_pro           f_px from fov, scale, resize, 1 / clamp        -> depth_pro.postprocess / mde_k_depth_pro_post
  metric3d_v2         un-pad, resize back (align_corners=False), [x canonical-to-metric], clamp(0, 300)   -> below
  vggt / streamvggt   cut the source frame's box out of the padded square, resize back, non-depths -> NaN     -> below
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

from . import _lib

METRIC3D_SIZE = (616, 1064)


def metric3d_geometry(src_h: int, src_w: int, size: Tuple[int, int] = METRIC3D_SIZE):
    """The keep-ratio resize factor and the centre padding Metric3D V2's input carries (models/metric3d_v2/onnx2trt.py:70-93,
    tools/evaluate_gt.py:133-139): -> (scale, (inner_h, inner_w), (top, bottom, left, right)); int() truncation as there."""
    scale = min(size[0] / src_h, size[1] / src_w)
    rh, rw = int(src_h * scale), int(src_w * scale)
    pad_h, pad_w = size[0] - rh, size[1] - rw
    return scale, (rh, rw), (pad_h // 2, pad_h - pad_h // 2, pad_w // 2, pad_w - pad_w // 2)


def metric3d_postprocess(depth_ptr: int, src_h: int, src_w: int, out, size: Tuple[int, int] = METRIC3D_SIZE,
                         focal_px: Optional[float] = None, stream_handle: int = 0) -> None:
    """models/metric3d_v2/onnx2trt.py:148-158 on the device.  `depth_ptr`: the engine's float32 [size] output; `out`: float32
    [src_h, src_w] device tensor.  `focal_px` None keeps the script's canonical depth; a focal length in pixels applies the
    de-canonical transform canonical * focal * scale / 1000 before the clamp (tools/evaluate_gt.py:162-184)."""
    scale, (rh, rw), (top, _, left, _) = metric3d_geometry(src_h, src_w, size)
    mul = 1.0 if focal_px is None else float(focal_px) * scale / 1000.0
    first = int(depth_ptr) + 4 * (top * size[1] + left)
    _lib.check(_lib.load().mde_k_resize_depth_halfpixel(C.c_void_p(first), size[1], rh, rw, C.c_void_p(out.data_ptr()), src_h, src_w,
                                                        mul, 0.0, 300.0, C.c_void_p(int(stream_handle))), "mde_k_resize_depth_halfpixel")


def vggt_box(src_h: int, src_w: int, net: int):
    """Where the source frame lands inside the padded square on the network's grid (tools/evaluate_gt.py:192-208
    `_square_pad_geometry`; core/preprocess.py:254-265 carries the same four floats): scale = net / max_dim, the model's own
    arithmetic even when the padded image is one pixel short of square."""
    m = max(src_w, src_h)
    left, top = (m - src_w) // 2, (m - src_h) // 2
    s = net / m
    return left * s, top * s, (left + src_w) * s, (top + src_h) * s


def vggt_postprocess(depth_ptr: int, net_h: int, net_w: int, src_h: int, src_w: int, out, stream_handle: int = 0) -> None:
    """tools/evaluate_gt.py:240-262 `_square_pad_depth` on the device.  `depth_ptr`: one frame of the engine's float32
    [net_h, net_w] depth output; `out`: float32 [src_h, src_w] device tensor (NaN where the value is not above 1e-6)."""
    x1, y1, x2, y2 = vggt_box(src_h, src_w, net_w)
    r0, r1, c0, c1 = int(round(y1)), int(round(y2)), int(round(x1)), int(round(x2))
    first = int(depth_ptr) + 4 * (r0 * net_w + c0)
    _lib.check(_lib.load().mde_k_resize_depth_halfpixel_nan(C.c_void_p(first), net_w, r1 - r0, c1 - c0, C.c_void_p(out.data_ptr()), src_h, src_w,
                                                            1e-6, C.c_void_p(int(stream_handle))), "mde_k_resize_depth_halfpixel_nan")
