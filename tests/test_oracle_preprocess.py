"""Pins oracle/preprocess_np.py (the CPU restatement of core/preprocess.py's stretch + ImageNet
pipeline) against: the committed golden vectors made by the REFERENCE module itself
(oracle/make_golden.py), the cv2 build in this image, and -- where /root/reference exists -- the
reference module imported live."""
import hashlib
import os
import sys

import numpy as np
import pytest

from oracle import preprocess_np as P

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "preprocess_golden.npz")
SOURCES = [(480, 640), (720, 1280), (500, 500), (1036, 1036), (300, 777)]


def synthetic(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def test_against_reference_golden_vectors():
    g = np.load(GOLDEN)
    n_full = n_sha = 0
    for i, (h, w) in enumerate(SOURCES):
        img = synthetic(i, h, w)
        for th, tw in [(70, 84), (56, 56)]:
            ref = g[f"full_seed{i}_{h}x{w}_to_{th}x{tw}"]
            got = P.preprocess_stretch_imagenet(img, th, tw)
            assert got.dtype == np.float32 and got.shape == ref.shape
            assert np.array_equal(got, ref)                      # byte-exact, not close
            n_full += 1
        got = P.preprocess_stretch_imagenet(img, 518, 518)
        sha = hashlib.sha256(np.ascontiguousarray(got).tobytes()).digest()
        assert sha == g[f"sha_seed{i}_{h}x{w}_to_518x518"].tobytes()
        n_sha += 1
    assert n_full == 10 and n_sha == 5


@pytest.mark.parametrize("src", [(480, 640), (2268, 3024), (1036, 1036), (259, 259), (37, 41), (1, 1), (2, 3),
                                 (1000, 333), (519, 517), (1035, 1037), (518, 518)])
@pytest.mark.parametrize("dst", [(518, 518), (616, 1064), (384, 384), (14, 14)])
def test_resize_restatement_bit_exact_vs_cv2(src, dst):
    cv2 = pytest.importorskip("cv2")
    img = synthetic(7, *src)
    assert np.array_equal(cv2.resize(img, (dst[1], dst[0]), interpolation=cv2.INTER_LINEAR),
                          P.resize_linear_u8(img, *dst))


def test_exact_2x_downscale_uses_area_rule():
    """cv2 silently switches INTER_LINEAR to the INTER_AREA fast path at exactly 2x (edge case of a-1)."""
    cv2 = pytest.importorskip("cv2")
    img = synthetic(3, 1036, 1036)
    out = P.resize_linear_u8(img, 518, 518)
    assert np.array_equal(out, cv2.resize(img, (518, 518), interpolation=cv2.INTER_LINEAR))
    a = img.astype(np.int64)
    assert np.array_equal(out, ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8))


def test_lut_is_the_whole_float_stage():
    lut = P.norm_lut()
    assert lut.shape == (3, 256) and lut.dtype == np.float32
    img = synthetic(0, 480, 640)
    small = P.resize_linear_u8(np.ascontiguousarray(img[:, :, ::-1]), 518, 518)
    via_lut = np.stack([lut[c][small[:, :, c]] for c in range(3)])[None]
    assert np.array_equal(via_lut, P.preprocess_stretch_imagenet(img, 518, 518))


def test_im2col_layout_is_conv_layout():
    torch = pytest.importorskip("torch")
    x = np.random.default_rng(0).standard_normal((2, 3, 28, 42)).astype(np.float32)
    w = np.random.default_rng(1).standard_normal((5, 3, 14, 14)).astype(np.float32)
    cols = P.im2col(x, 14, 640)
    assert cols.shape == (2 * 2 * 3, 640) and np.all(cols[:, 588:] == 0)
    ref = torch.nn.functional.conv2d(torch.from_numpy(x), torch.from_numpy(w), stride=14).flatten(2).transpose(1, 2).reshape(-1, 5)
    got = cols[:, :588] @ w.reshape(5, -1).T
    assert np.allclose(got, ref.numpy(), atol=1e-3)
    # row = (b, gy, gx), column = c*196 + ky*14 + kx
    assert cols[1 * 6 + 1 * 3 + 2, 2 * 196 + 5 * 14 + 7] == x[1, 2, 14 + 5, 28 + 7]


@pytest.mark.skipif(not os.path.isdir("/root/reference/core"), reason="reference checkout not present on this box")
@pytest.mark.parametrize("model", ["depth_anything_v2", "depth_anything_v3", "distill_any_depth"])
def test_against_live_reference_module(model):
    sys.path.insert(0, "/root/reference")
    try:
        from core import preprocess as ref
    finally:
        sys.path.remove("/root/reference")
    for i, (h, w) in enumerate(SOURCES[:3]):
        img = synthetic(i, h, w)
        r, geom = ref.preprocess_for(img, model, (518, 518))
        assert np.array_equal(r, P.preprocess_stretch_imagenet(img, 518, 518))
        assert (geom.src_h, geom.src_w, geom.dst_h, geom.dst_w) == (h, w, 518, 518)


# ------------------------------------------------------------------ Metric3D V2: keep-ratio + centre pad, no normalisation
def test_metric3d_variant_against_reference_golden_vectors():
    g = np.load(GOLDEN)
    for i, (h, w) in enumerate(SOURCES):
        img = synthetic(i, h, w)
        for th, tw in [(70, 98), (56, 56)]:
            ref = g[f"m3d_full_seed{i}_{h}x{w}_to_{th}x{tw}"]
            got = P.preprocess_pad_none(img, th, tw)
            assert got.dtype == np.float32 and np.array_equal(got, ref)        # byte-exact
        got = P.preprocess_pad_none(img, 616, 1064)
        assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).digest() == g[f"m3d_sha_seed{i}_{h}x{w}_to_616x1064"].tobytes()
        assert list(P.pad_geometry(h, w, 616, 1064)) == list(g[f"m3d_geom_seed{i}_{h}x{w}_to_616x1064"])


def test_metric3d_pad_uses_truncation_and_the_rounded_mean_colour():
    # 300 x 777 -> 616 x 1064: scale = min(616/300, 1064/777) = 1.3694; int() truncates 410.81 -> 410 (round would give 411)
    assert P.pad_geometry(300, 777, 616, 1064) == (410, 1064, 103, 0)
    out = P.preprocess_pad_none(synthetic(4, 300, 777), 616, 1064)
    assert out[0, :, 0, 0].tolist() == [124.0, 116.0, 104.0]                    # saturate_cast of (123.675, 116.28, 103.53)
    assert out[0, :, 102, 5].tolist() == [124.0, 116.0, 104.0] and out.max() <= 255.0 and out.min() >= 0.0


@pytest.mark.skipif(not os.path.isdir("/root/reference/core"), reason="reference checkout not present on this box")
def test_metric3d_variant_against_live_reference_module():
    sys.path.insert(0, "/root/reference")
    try:
        from core import preprocess as ref
    finally:
        sys.path.remove("/root/reference")
    for i, (h, w) in enumerate(SOURCES):
        img = synthetic(i, h, w)
        for size in [(616, 1064), (518, 518)]:
            r, geom = ref.preprocess_for(img, "metric3d_v2", size)
            assert np.array_equal(r, P.preprocess_pad_none(img, *size))
            assert (geom.inner_h, geom.inner_w, geom.pad_top, geom.pad_left) == P.pad_geometry(h, w, *size)


# ------------------------------------------------------------------ VGGT / StreamVGGT: white square pad + cubic resize
VGGT_SOURCES = SOURCES + [(501, 500), (33, 57)]


def test_vggt_preprocessing_against_reference_golden_vectors():
    """Goldens made by the reference module (core.preprocess.preprocess_for(img, 'vggt', size)) with IPP switched off."""
    g = np.load(GOLDEN)
    n = 0
    for i, (h, w) in enumerate(VGGT_SOURCES):
        img = synthetic(i, h, w)
        for th, tw in [(70, 70), (56, 84)]:
            ref = g[f"vggt_full_seed{i}_{h}x{w}_to_{th}x{tw}"]
            got = P.preprocess_square_pad_cubic(img, th, tw)
            assert got.dtype == np.float32 and got.shape == ref.shape == (1, 1, 3, th, tw)
            assert np.array_equal(got, ref)
            n += 1
        got = P.preprocess_square_pad_cubic(img, 518, 518)
        assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).digest() == g[f"vggt_sha_seed{i}_{h}x{w}_to_518x518"].tobytes()
        # the four box floats the post-processing slices with (core/preprocess.py:254-265): dst / max_dim scaling of the source frame
        top, left, _, _ = P.square_pad_geometry(h, w)
        s = 518 / max(h, w)
        assert np.allclose(g[f"vggt_box_seed{i}_{h}x{w}_to_518x518"], [left * s, top * s, (left + w) * s, (top + h) * s], rtol=0, atol=1e-12)
    assert n == 14


@pytest.mark.parametrize("src,dst", [((48, 64), (37, 50)), ((60, 60), (100, 90)), ((480, 640), (518, 518)), ((33, 57), (70, 70)),
                                     ((300, 777), (56, 56)), ((2, 3), (9, 5)), ((720, 1280), (518, 518)), ((64, 64), (64, 33))])
def test_cubic_restatement_bit_exact_vs_cv2_own_path(src, dst):
    """OpenCV's own INTER_CUBIC for 8-bit images (IPP off).  With IPP on -- the wheel's default -- the same call differs by one
    level in a few per cent of the pixels: Intel's closed implementation, not restatable; the difference is bounded here."""
    cv2 = pytest.importorskip("cv2")
    img = synthetic(11, *src)
    was = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        ref = cv2.resize(img, (dst[1], dst[0]), interpolation=cv2.INTER_CUBIC)
        cv2.ipp.setUseIPP(True)
        ipp = cv2.resize(img, (dst[1], dst[0]), interpolation=cv2.INTER_CUBIC)
    finally:
        cv2.ipp.setUseIPP(was)
    got = P.resize_cubic_u8(img, *dst)
    assert np.array_equal(got, ref)
    assert int(np.abs(got.astype(int) - ipp.astype(int)).max()) <= 1


@pytest.mark.skipif(not os.path.exists("/root/reference/core/preprocess.py"), reason="reference checkout not mounted")
def test_vggt_preprocessing_against_live_reference_module():
    cv2 = pytest.importorskip("cv2")
    sys.path.insert(0, "/root/reference")
    try:
        from core import preprocess as ref
        was = cv2.ipp.useIPP()
        cv2.ipp.setUseIPP(False)
        try:
            for seed, (h, w) in enumerate([(480, 640), (769, 1025), (500, 500)]):
                img = synthetic(seed, h, w)
                t, geom = ref.preprocess_for(img, "vggt", (518, 518))
                assert np.array_equal(t, P.preprocess_square_pad_cubic(img, 518, 518))
                t2, _ = ref.preprocess_for(img, "streamvggt", (518, 518))
                assert np.array_equal(t, t2)
        finally:
            cv2.ipp.setUseIPP(was)
    finally:
        sys.path.remove("/root/reference")


# ------------------------------------------------------------------------------------------------ Depth-Anything-AC
def test_depth_anything_ac_contract_against_reference_golden_vectors():
    """core/preprocess.py:470-476 `da_ac`: depth_anything_v2's stretch with the division by 255 in float32 (the last bits
    differ), and the "ceil" keep-ratio rule for the network size (4:3 -> 518 x 700 where depth_anything_v2 gets 518 x 686)."""
    import hashlib
    g = np.load(GOLDEN)
    keys = [k for k in g.files if k.startswith("ac_full_")]
    assert len(keys) == 5
    for k in keys:
        seed, src = int(k.split("seed")[1].split("_")[0]), tuple(int(v) for v in k.split("_")[3].split("x"))
        img = np.random.default_rng(seed).integers(0, 256, (*src, 3), dtype=np.uint8)
        assert np.array_equal(P.preprocess_stretch_imagenet(img, 56, 56, "float32"), g[k]), k
        big = P.preprocess_stretch_imagenet(img, 518, 518, "float32")
        assert hashlib.sha256(big.tobytes()).digest() == g[f"ac_sha_seed{seed}_{src[0]}x{src[1]}_to_518x518"].tobytes()
        assert not np.array_equal(big, P.preprocess_stretch_imagenet(img, 518, 518))          # float64 division: other last bits
        want = g[f"ac_keep_ratio_seed{seed}_{src[0]}x{src[1]}"]
        assert P.keep_ratio_size(*src, 518, 14, "ceil") == tuple(want[:2]) and P.keep_ratio_size(*src, 518, 14, "constrain") == tuple(want[2:])
        from monocular_depth_estimation_trt_b200 import weights as W
        assert W.keep_ratio_size(*src, 518, 14, "ceil") == tuple(want[:2]) and W.keep_ratio_size(*src, 518, 14, "constrain") == tuple(want[2:])
    assert P.keep_ratio_size(480, 640) == (518, 700) and P.keep_ratio_size(480, 640, rounding="constrain") == (518, 686)


def test_depth_anything_ac_native_profile_against_reference_golden_vectors():
    """models/depth_anything_ac/onnx2trt.py:50-75 `profile = 'native'` (core/preprocess.py:470-476 `da_ac(h, w, stretch=False)`):
    no uint8 resize, float32 / 255, cv2 float INTER_CUBIC to the keep-ratio "ceil" size, ImageNet statistics in float64 --
    byte-exact against tensors the reference module produced (IPP off, see `resize_cubic_f32`)."""
    import hashlib
    g = np.load(GOLDEN)
    keys = [k for k in g.files if k.startswith("acn_full_")]
    assert len(keys) == 6
    for k in keys:
        seed, src = int(k.split("seed")[1].split("_")[0]), tuple(int(v) for v in k.split("_")[3].split("x"))
        img = np.random.default_rng(seed).integers(0, 256, (*src, 3), dtype=np.uint8)
        small = P.preprocess_keep_ratio_cubic_f32(img, 56)
        assert small.shape == g[k].shape and np.array_equal(small, g[k]), k
        big = P.preprocess_keep_ratio_cubic_f32(img, 518)
        assert tuple(g[f"acn_size_seed{seed}_{src[0]}x{src[1]}_target518"]) == big.shape[2:]
        assert hashlib.sha256(big.tobytes()).digest() == g[f"acn_sha_seed{seed}_{src[0]}x{src[1]}_target518"].tobytes()
    assert P.preprocess_keep_ratio_cubic_f32(synthetic(0, 480, 640)).shape == (1, 3, 518, 700)


@pytest.mark.parametrize("src", [(480, 640), (769, 1025), (500, 500), (37, 53), (1036, 720), (518, 518), (123, 457), (5, 4)])
@pytest.mark.parametrize("dst", [(518, 700), (518, 518), (70, 98), (56, 57)])
def test_float_cubic_restatement_bit_exact_vs_cv2_own_path(src, dst):
    """`resize_cubic_f32` against cv2's own float INTER_CUBIC (IPP off): taps left to right horizontally, right to left in the
    vertical vector loop, left to right in the scalar loop that finishes a row -- bit for bit, up- and down-scaling, borders.
    With IPP on (the wheel's default) cv2 differs from its own path by up to ~5e-5 on 0..1 data."""
    cv2 = pytest.importorskip("cv2")
    x = np.ascontiguousarray(synthetic(11, *src).astype(np.float32) / np.float32(255.0))
    was = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)
        ref = cv2.resize(x, (dst[1], dst[0]), interpolation=cv2.INTER_CUBIC)
        assert np.array_equal(P.resize_cubic_f32(x, *dst), ref)
        if was and src != dst:
            cv2.ipp.setUseIPP(True)
            ipp = cv2.resize(x, (dst[1], dst[0]), interpolation=cv2.INTER_CUBIC)
            assert float(np.abs(ipp - ref).max()) < 2e-4
    finally:
        cv2.ipp.setUseIPP(was)
