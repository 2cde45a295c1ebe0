"""Host-side logic of the Depth Pro path (no GPU): tap selection, weight packing, the folded transposed convolution."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from monocular_depth_estimation_trt_b200 import depth_pro as DPE


def test_tap_plan_keeps_hooks_and_last_block():
    taps, (a, b) = DPE.tap_plan(24, (11, 5))
    assert len(taps) == 4 and taps == sorted(taps) and taps[3] == 23 and taps[a] == 11 and taps[b] == 5
    taps, (a, b) = DPE.tap_plan(12, (8, 5))
    assert taps[3] == 11 and taps[a] == 8 and taps[b] == 5
    with pytest.raises(ValueError):
        DPE.tap_plan(12, (11, 5))                                      # the last block is not a hook
    with pytest.raises(ValueError):
        DPE.tap_plan(12, (5, 5))


def test_deconv_packing_is_a_gemm_with_pixel_shuffle():
    """ConvTranspose2d(k=2, s=2) == rows (y, x) x packed weight, column (ky*2+kx)*cout + o -> pixel (2y+ky, 2x+kx); with a
    1x1 convolution folded in, the same GEMM gives conv1x1(conv_transpose(x))."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 16, 5, 7, generator=g)
    w = torch.randn(16, 24, 2, 2, generator=g)
    w1 = torch.randn(8, 24, generator=g)
    for fold, ref in ((None, F.conv_transpose2d(x, w, stride=2)),
                      (w1, F.conv2d(F.conv_transpose2d(x, w, stride=2), w1[:, :, None, None]))):
        cout = ref.shape[1]
        p = DPE.pack_deconv2(w, torch.float32, then_1x1=fold)           # [4*cout, cin]
        rows = x[0].permute(1, 2, 0).reshape(35, 16) @ p.t()           # [(y, x), 4*cout]
        got = rows.reshape(5, 7, 2, 2, cout).permute(4, 0, 2, 1, 3).reshape(1, cout, 10, 14)
        assert torch.allclose(got, ref, atol=1e-4)


def test_conv_packings():
    g = torch.Generator().manual_seed(1)
    w = torch.randn(8, 24, 3, 3, generator=g)
    p = DPE.pack_conv3x3(w, torch.float32)
    assert p.shape == (8, 9 * 64) and torch.equal(p.reshape(8, 9, 64)[:, 4, :24], w[:, :, 1, 1]) and float(p.reshape(8, 9, 64)[:, :, 24:].abs().max()) == 0
    q = DPE.pack_conv3x3_s2(w, torch.float32)
    assert q.shape == (8, 9 * 24) and torch.equal(q.reshape(8, 9, 24)[:, 2], w[:, :, 0, 2])


def test_engine_refuses_other_sizes_and_precisions():
    with pytest.raises(ValueError):
        DPE.DepthProEngine({}, precision="fp32")
    with pytest.raises(ValueError):
        DPE.DepthProEngine({}, image_size=1024)


def test_depth_pro_description_round_trips(tmp_path):
    from monocular_depth_estimation_trt_b200 import weights as W
    meta = W.describe_depth_pro("vitl")
    assert meta["family"] == "depth_pro" and meta["hook_blocks"] == [11, 5] and meta["input_h"] == 1536 and meta["features"] == 256
    path = str(tmp_path / "dp.mdew")
    W.save(path, {"head.4.bias": torch.zeros(1)}, meta)
    assert W.read_meta(path) == meta


def test_vggt_aggregator_refuses_bad_configurations():
    from monocular_depth_estimation_trt_b200 import vggt as V
    with pytest.raises(ValueError):
        V.Aggregator({}, 1024, 1, 16, 37, 37, frames_total=16, precision="fp32")
    with pytest.raises(ValueError):
        V.Aggregator({}, 1000, 1, 16, 37, 37, frames_total=16)          # head dimension must be 64
    with pytest.raises(ValueError):
        V.Aggregator({}, 1024, 1, 16, 37, 37, frames_total=15, world=2)
    with pytest.raises(ValueError):
        V.Aggregator({}, 1024, 1, 16, 37, 37, frames_total=16, gather="ring")


def test_metric3d_geometry_matches_the_reference_helper():
    """tools/evaluate_gt.py:133-139 `_metric3d_geometry`, restated; checked live when the checkout is mounted."""
    import importlib.util, os
    from monocular_depth_estimation_trt_b200 import postprocess as PP
    cases = [(480, 640), (768, 1024), (1064, 616), (2268, 3024), (500, 500)]
    assert PP.metric3d_geometry(480, 640) == (616 / 480, (616, 821), (0, 0, 121, 122))
    ref = "/root/reference/tools/evaluate_gt.py"
    if os.path.exists(ref):
        src = open(ref).read()
        start = src.index("def _metric3d_geometry"); end = src.index("\ndef ", start + 10)
        ns = {"METRIC3D_SIZE": (616, 1064)}
        exec(compile(src[start:end], ref, "exec"), ns)                  # the helper alone (the module imports cv2 / tensorrt tooling)
        for h, w in cases:
            assert PP.metric3d_geometry(h, w) == ns["_metric3d_geometry"](h, w)


def test_required_tensors_are_exactly_what_the_oracle_model_holds(tmp_path):
    """The engine's checkpoint contract against the oracle's parameter list (mask_token aside, which inference never reads),
    and the export stage: a state dict saved as a checkpoint comes back as an .mdew file; a wrong file is refused by name."""
    from oracle import depth_pro_torch as DP
    from monocular_depth_estimation_trt_b200 import weights as W
    for enc, feat in (("vits", 64), ("vitl", 256)):
        need = DPE.required_tensors(enc, feat)
        have = {k: v for k, v in DP.full_param_shapes(enc, feat).items() if not k.endswith("mask_token")}
        assert need == have
    sd = DP.init_full_state_dict("vits", features=64, seed=1)
    ckpt = str(tmp_path / "depth_pro.pt")
    torch.save({"module." + k: v for k, v in sd.items()}, ckpt)
    meta = DPE.export_checkpoint(ckpt, str(tmp_path / "dp.mdew"), encoder="vits", features=64, hook_blocks=(8, 5))
    back, meta2 = W.load(str(tmp_path / "dp.mdew"))
    assert meta2 == meta and meta["family"] == "depth_pro" and len(meta["source_checkpoint_sha256"]) == 64
    assert set(back) == set(DPE.required_tensors("vits", 64)) and np.array_equal(back["head.4.bias"], sd["head.4.bias"].numpy())
    torch.save({k: v for k, v in sd.items() if not k.startswith("fov.")}, ckpt)
    with pytest.raises(ValueError, match="tensors missing"):
        DPE.export_checkpoint(ckpt, str(tmp_path / "x.mdew"), encoder="vits", features=64)


def test_vggt_required_tensors_and_export_stage(tmp_path):
    """The VGGT / StreamVGGT checkpoint contract against the oracle's parameter list, and the export stage: a state dict
    saved the way upstream's checkpoints are (extra heads the wrappers never run, a `module.` prefix) comes back as an .mdew
    file with the family in its description; a file without the depth head is refused by name."""
    from oracle import vggt_torch as V
    from monocular_depth_estimation_trt_b200 import vggt as P, weights as W
    for enc, depth, feat, oc in (("vits", 2, 64, (48, 96, 192, 384)), ("vitl", 24, 256, (256, 512, 1024, 1024))):
        need = P.required_tensors(enc, depth, feat, oc)
        if enc == "vits":
            sd = V.init_vggt(enc, depth=depth, features=feat, out_channels=oc, seed=3)
            have = {k: tuple(v.shape) for k, v in sd.items() if not k.endswith("mask_token")}
            assert need == have, (sorted(set(need) ^ set(have))[:6])
        else:
            # trunk 7 + 24 x 14, two token sets, 48 aggregator blocks x 18, head: norm 2, per level 5, 7 residual units x 4, 6 + 6
            assert len(need) == 7 + 24 * 14 + 2 + 48 * 18 + 2 + 4 * 5 + 7 * 4 + 6 + 6
    ckpt = str(tmp_path / "model.pt")
    extra = {"camera_head.trunk.0.weight": torch.zeros(3, 3), "point_head.norm.weight": torch.zeros(8)}
    torch.save({"module." + k: v for k, v in {**sd, **extra}.items()}, ckpt)
    kw = dict(encoder="vits", depth=2, features=64, out_channels=(48, 96, 192, 384), taps=(0, 0, 1, 1), frames=3)
    meta = P.export_checkpoint(ckpt, str(tmp_path / "v.mdew"), family="streamvggt", **kw)
    back, meta2 = W.load(str(tmp_path / "v.mdew"))
    assert meta2 == meta and meta["family"] == "streamvggt" and meta["frames"] == 3 and len(meta["source_checkpoint_sha256"]) == 64
    assert set(back) == set(P.required_tensors("vits", 2, 64, (48, 96, 192, 384)))
    assert np.array_equal(back["aggregator.camera_token"], sd["aggregator.camera_token"].numpy())
    torch.save({k: v for k, v in sd.items() if not k.startswith("depth_head.")}, ckpt)
    with pytest.raises(ValueError, match="tensors missing"):
        P.export_checkpoint(ckpt, str(tmp_path / "x.mdew"), **kw)
