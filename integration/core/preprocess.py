"""`core.preprocess` where the reference checkout is absent: `preprocess_for` for the models of the hot path over the pinned
restatement (oracle/preprocess_np.py: byte-exact with the reference module on its golden vectors)."""
import os
import sys
from dataclasses import dataclass

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from oracle import preprocess_np as _P  # noqa: E402


@dataclass
class Geometry:            # the fields of core/preprocess.py:59-84 the scripts of the hot path read
    src_h: int
    src_w: int
    dst_h: int
    dst_w: int
    inner_h: int = 0
    inner_w: int = 0
    pad_top: int = 0
    pad_left: int = 0


def preprocess_for(img_bgr, model, size, **kw):
    h, w = size
    if model in ("depth_anything_v2", "distill_any_depth"):
        return _P.preprocess_stretch_imagenet(img_bgr, h, w), Geometry(img_bgr.shape[0], img_bgr.shape[1], h, w, h, w)
    if model == "metric3d_v2":
        ih, iw, top, left = _P.pad_geometry(img_bgr.shape[0], img_bgr.shape[1], h, w)
        return _P.preprocess_pad_none(img_bgr, h, w), Geometry(img_bgr.shape[0], img_bgr.shape[1], h, w, ih, iw, top, left)
    raise KeyError(f"no preprocessing restated for {model!r}")
