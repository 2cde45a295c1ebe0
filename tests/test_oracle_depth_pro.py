"""Pins oracle/depth_pro_torch.py (crop pyramid, hooked trunk, patch merge) against transformers' independent
DepthProPatchEncoder on copied weights."""
import pytest
import torch

from oracle import dav2_torch as O, depth_pro_torch as DP


def test_merge_geometry():
    t = torch.arange(25 * 576, dtype=torch.float32).reshape(25, 576, 1)
    m = DP.merge_crops(t, 5, 3)
    assert m.shape == (96, 96, 1)                                  # 5 * 24 - 2 * 3 * 4
    assert m[0, 0, 0] == 0 and m[20, 20, 0] == 20 * 24 + 20        # crop 0 keeps rows / columns 0..20
    assert m[21, 21, 0] == 6 * 576 + 3 * 24 + 3                    # then crop (1, 1) starts at its token (3, 3)
    assert DP.merge_crops(t[:9], 3, 6).shape == (48, 48, 1) and DP.merge_crops(t[:1], 1, 12).shape == (24, 24, 1)


def test_patch_encoder_matches_transformers_depth_pro():
    pytest.importorskip("transformers")
    from transformers import DepthProConfig, Dinov2Config
    from transformers.models.depth_pro.modeling_depth_pro import DepthProPatchEncoder
    import hf_bridge as H
    torch.manual_seed(3)
    c = O.MODEL_CONFIGS["vits"]
    sd = O.init_state_dict("vits", seed=8, patch=16, pos_grid=24)
    vit = Dinov2Config(hidden_size=c["embed_dim"], num_hidden_layers=c["depth"], num_attention_heads=c["num_heads"], image_size=384, patch_size=16)
    # hooks (8, 5) = taps 2 and 1 of the ViT-S tap list [2, 5, 8, 11]; the released ViT-L model hooks (11, 5) = taps 1 and 0
    cfg = DepthProConfig(patch_model_config=vit, image_model_config=vit, intermediate_hook_ids=[8, 5], use_fov_model=False)
    enc = DepthProPatchEncoder(cfg).eval()
    hf = {k[len("backbone."):]: v for k, v in H.to_hf({**O.init_state_dict("vits", seed=8), **sd}, "vits").items() if k.startswith("backbone.")}
    enc.model.load_state_dict(hf, strict=True)
    image = torch.randn(3, 1536, 1536)
    with torch.no_grad():
        ref = enc(image[None])                                       # [f24, f48, f96, hook8 96, hook5 96], each [1, D, S, S]
    got = DP.patch_encoder_features(sd, image, "vits", hook_taps=(2, 1))
    assert [tuple(r.shape[-2:]) for r in ref] == [(24, 24), (48, 48), (96, 96), (96, 96), (96, 96)]
    for r, g in zip(ref, got):
        r = r[0].permute(1, 2, 0)
        assert r.shape == g.shape
        assert float((r - g).abs().max()) < 3e-4 * float(r.abs().max())


def test_full_model_matches_transformers_depth_pro():
    """The whole restated model (three trunks, upsampling neck, fusion decoder, depth head, field-of-view head) against
    transformers' independent DepthProForDepthEstimation on copied weights, at the exported 1536 x 1536 size."""
    pytest.importorskip("transformers")
    import hf_bridge as H
    sd = DP.init_full_state_dict("vits", features=64, seed=11)
    model = H.depth_pro_hf_model("vits", 64, hook_ids=(8, 5))
    missing, unexpected = model.load_state_dict(H.depth_pro_to_hf(sd, "vits"), strict=False)
    # transformers allocates a residual_layer1 for the first fusion layer that its forward never uses
    assert not unexpected and all(k.startswith("fusion_stage.intermediate.0.residual_layer1.") for k in missing), (missing, unexpected)
    torch.manual_seed(5)
    x = torch.randn(1, 3, 1536, 1536)
    with torch.no_grad():
        ref = model(pixel_values=x)
    inv, fov = DP.full_forward(sd, x, "vits", hook_taps=(2, 1))
    assert inv.shape == (1, 1, 1536, 1536) and fov.shape == (1,)
    r = ref.predicted_depth[:, None]
    assert float(r.max()) > 0 and float((r > 0).float().mean()) > 0.2                 # not a dead ReLU map
    assert float((inv - r).abs().max()) < 3e-4 * float(r.abs().max())
    assert abs(float(fov) - float(ref.field_of_view)) < 3e-4 * max(1.0, abs(float(ref.field_of_view)))


def test_postprocess_follows_the_reference_script():
    """models/depth_pro/onnx2trt.py:118-134 on a constant map: f_px = 0.5 W / tan(fov / 2), depth = f_px / (W * inv)."""
    inv = torch.full((1, 1, 8, 8), 0.5)
    depth, f_px = DP.postprocess(inv, torch.tensor([90.0]), 4, 6)
    assert depth.shape == (1, 1, 4, 6) and abs(float(f_px) - 3.0) < 1e-5
    assert torch.allclose(depth, torch.full_like(depth, 1.0), atol=1e-6)              # 1 / (0.5 * 6 / 3)
