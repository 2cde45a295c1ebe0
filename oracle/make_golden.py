"""Generate tests/golden/*.npz.  Run in the survey/build container, where /root/reference exists:

    python oracle/make_golden.py

preprocess_golden.npz  outputs of the REFERENCE module itself (imported from /root/reference:
                       core.preprocess.preprocess_for(img, 'depth_anything_v2', size)) on the
                       reference's synthetic-input convention (tests/test_preprocess.py:47-51):
                       full tensors at small target sizes, sha256 of the float32 bytes at 518x518; the same for
                       'metric3d_v2' (keep-ratio + pad, 616x1064) together with the geometry it reports, and for 'vggt'
                       (white square pad + cubic resize, IPP switched off: see oracle/preprocess_np.py resize_cubic_u8), and for
                       'depth_anything_ac' (float32 division by 255) with its keep-ratio network sizes, and its
                       `native` profile (stretch=False: float32 cubic resize to the keep-ratio size, IPP off).
dav2_vits_golden.npz   the oracle's own ViT-S 518x518 batch-1 forward (BASELINE config 1) with the
                       seeded, calibrated init: a 7x-strided subsample of the depth map and summary
                       statistics.  It pins the oracle against drift between hosts; the oracle
                       itself is pinned against transformers' DepthAnything in
                       tests/test_oracle_model.py.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

SOURCES = [(480, 640), (720, 1280), (500, 500), (1036, 1036), (300, 777)]
SMALL_TARGETS = [(70, 84), (56, 56)]


def synthetic(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def main():
    os.makedirs(OUT, exist_ok=True)
    sys.path.insert(0, "/root/reference")
    from core import preprocess as ref   # the reference module, unmodified

    blob = {}
    for i, (h, w) in enumerate(SOURCES):
        img = synthetic(i, h, w)
        for (th, tw) in SMALL_TARGETS:
            t, _ = ref.preprocess_for(img, "depth_anything_v2", (th, tw))
            blob[f"full_seed{i}_{h}x{w}_to_{th}x{tw}"] = t
        t, _ = ref.preprocess_for(img, "depth_anything_v2", (518, 518))
        digest = hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest()
        blob[f"sha_seed{i}_{h}x{w}_to_518x518"] = np.frombuffer(bytes.fromhex(digest), dtype=np.uint8)
    # Metric3D V2: keep-ratio resize (truncated inner size) + centre pad with the mean colour, no normalisation
    for i, (h, w) in enumerate(SOURCES):
        img = synthetic(i, h, w)
        for (th, tw) in [(70, 98), (56, 56)]:
            t, _ = ref.preprocess_for(img, "metric3d_v2", (th, tw))
            blob[f"m3d_full_seed{i}_{h}x{w}_to_{th}x{tw}"] = t
        t, geom = ref.preprocess_for(img, "metric3d_v2", (616, 1064))
        digest = hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest()
        blob[f"m3d_sha_seed{i}_{h}x{w}_to_616x1064"] = np.frombuffer(bytes.fromhex(digest), dtype=np.uint8)
        blob[f"m3d_geom_seed{i}_{h}x{w}_to_616x1064"] = np.array([geom.inner_h, geom.inner_w, geom.pad_top, geom.pad_left], dtype=np.int64)
    # VGGT / StreamVGGT: white square pad + ONE cubic resize + / 255 (float32, rank 5).  Generated with IPP off: the
    # opencv-python wheel routes 8-bit cubic through Intel IPP by default, OpenCV's own path is what can be restated.
    import cv2
    cv2.ipp.setUseIPP(False)
    for i, (h, w) in enumerate(SOURCES + [(501, 500), (33, 57)]):
        img = synthetic(i, h, w)
        for (th, tw) in [(70, 70), (56, 84)]:
            t, _ = ref.preprocess_for(img, "vggt", (th, tw))
            blob[f"vggt_full_seed{i}_{h}x{w}_to_{th}x{tw}"] = t
        t, geom = ref.preprocess_for(img, "vggt", (518, 518))
        digest = hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest()
        blob[f"vggt_sha_seed{i}_{h}x{w}_to_518x518"] = np.frombuffer(bytes.fromhex(digest), dtype=np.uint8)
        blob[f"vggt_box_seed{i}_{h}x{w}_to_518x518"] = np.array(geom.box, dtype=np.float64)
    cv2.ipp.setUseIPP(True)
    # Depth-Anything-AC at the square bench size: the stretch of depth_anything_v2 with the division by 255 in float32
    # (core/preprocess.py:470-476), and the network size its keep-ratio rule ("ceil") gives for each source frame
    for i, (h, w) in enumerate(SOURCES):
        img = synthetic(i, h, w)
        t, _ = ref.preprocess_for(img, "depth_anything_ac", (56, 56))
        blob[f"ac_full_seed{i}_{h}x{w}_to_56x56"] = t
        t, _ = ref.preprocess_for(img, "depth_anything_ac", (518, 518))
        digest = hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest()
        blob[f"ac_sha_seed{i}_{h}x{w}_to_518x518"] = np.frombuffer(bytes.fromhex(digest), dtype=np.uint8)
        sizes = []
        for rounding in ("ceil", "constrain"):
            kw = {"min_val": 518} if rounding == "constrain" else {}
            sc = 518 / min(h, w)
            sizes += [ref._round_to_multiple(h * sc, 14, rounding, **kw), ref._round_to_multiple(w * sc, 14, rounding, **kw)]
        blob[f"ac_keep_ratio_seed{i}_{h}x{w}"] = np.array(sizes, dtype=np.int64)        # [ceil h, ceil w, constrain h, constrain w]
    # Depth-Anything-AC `native` profile (stretch=False): float32 / 255, cv2 float INTER_CUBIC to the keep-ratio "ceil" size,
    # ImageNet statistics in float64.  IPP off for the same reason as above (float cubic goes through IPP in the wheel).
    cv2.ipp.setUseIPP(False)
    for i, (h, w) in enumerate(SOURCES + [(37, 53)]):
        img = synthetic(i, h, w)
        t, geom = ref.preprocess_for(img, "depth_anything_ac", (56, 56), stretch=False)
        blob[f"acn_full_seed{i}_{h}x{w}_target56"] = t
        t, geom = ref.preprocess_for(img, "depth_anything_ac", (518, 518), stretch=False)
        digest = hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest()
        blob[f"acn_sha_seed{i}_{h}x{w}_target518"] = np.frombuffer(bytes.fromhex(digest), dtype=np.uint8)
        blob[f"acn_size_seed{i}_{h}x{w}_target518"] = np.array([geom.dst_h, geom.dst_w], dtype=np.int64)
    cv2.ipp.setUseIPP(True)
    np.savez_compressed(os.path.join(OUT, "preprocess_golden.npz"), **blob)
    print("wrote preprocess_golden.npz", len(blob), "entries")

    import torch
    from oracle import dav2_torch as O
    from oracle import preprocess_np as P
    torch.manual_seed(0)
    x = torch.from_numpy(P.preprocess_stretch_imagenet(synthetic(0, 480, 640), 518, 518))
    sd = O.init_state_dict("vits", seed=0)
    m, s = O.calibrate_head(sd, x, "vits")
    d = O.forward(sd, x, "vits", max_depth=20.0)[0].numpy()
    np.savez_compressed(os.path.join(OUT, "dav2_vits_golden.npz"),
                        depth_stride7=d[::7, ::7].astype(np.float32),
                        stats=np.array([d.min(), d.max(), d.mean(), d.std(), m, s], dtype=np.float64))
    print("wrote dav2_vits_golden.npz", d.shape, d.min(), d.max(), d.mean(), d.std(), "precalib", m, s)


if __name__ == "__main__":
    main()
