"""Pins oracle/dav2_torch.py: (1) against transformers' independent DepthAnything implementation with the
same weights, (2) against the committed golden sample of its own ViT-S 518x518 forward (BASELINE.json
configs[0]), (3) structural facts the reference states (parameter count, state-dict keys, I/O shape)."""
import os

import numpy as np
import pytest
import torch

import refsetup as R
from oracle import dav2_torch as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "dav2_vits_golden.npz")


def test_parameter_counts_match_upstream_models():
    n = {e: sum(int(np.prod(s)) for k, s in O.param_shapes(e).items() if "mask_token" not in k) for e in O.MODEL_CONFIGS}
    assert abs(n["vits"] / 1e6 - 24.785) < 0.01       # SURVEY section 8 c: 24.785 M = upstream Small
    assert abs(n["vitl"] / 1e6 - 335.32) < 0.05       # 335.32 M = upstream Large
    keys = O.param_shapes("vitl")
    for k in ("pretrained.blocks.23.attn.qkv.weight", "depth_head.scratch.refinenet4.resConfUnit2.conv2.bias",
              "depth_head.resize_layers.3.weight", "depth_head.scratch.output_conv2.2.bias"):
        assert k in keys


def test_vits_forward_matches_golden_sample():
    sd, x, depth, _ = R.reference("vits")
    g = np.load(GOLDEN)
    d = depth[0].numpy()
    assert d.shape == (518, 518) and d.dtype == np.float32
    # the golden was produced on another host: fp32 summation order may differ, values may not
    assert np.allclose(d[::7, ::7], g["depth_stride7"], rtol=2e-4, atol=2e-4)
    mn, mx, mean, std = g["stats"][:4]
    assert abs(d.min() - mn) < 1e-3 and abs(d.max() - mx) < 1e-3 and abs(d.mean() - mean) < 1e-3
    assert std > 3.0          # the calibrated init spans the metric range: the parity gate is not vacuous


def test_oracle_matches_transformers_depth_anything():
    pytest.importorskip("transformers")
    import hf_bridge as H
    sd, x, depth, _ = R.reference("vits")
    m = H.hf_model("vits", 20.0)
    missing = m.load_state_dict(H.to_hf(sd, "vits"), strict=True)
    with torch.no_grad():
        dh = m(pixel_values=x).predicted_depth
    rel = ((depth - dh).abs() / depth).max().item()
    assert rel < 5e-5, rel       # fp32 round-off between two independent implementations


def test_relative_head_and_batch_and_nonsquare():
    sd = O.init_state_dict("vits", seed=1)
    x = torch.randn(2, 3, 56, 84)
    d = O.forward(sd, x, "vits", max_depth=None)
    assert d.shape == (2, 56, 84) and float(d.min()) >= 0.0
    # batch entries are independent
    d0 = O.forward(sd, x[:1], "vits", max_depth=None)
    assert torch.allclose(d[:1], d0, atol=1e-5)


def test_calibration_makes_the_gate_meaningful():
    """Default-style init would give a constant 10.0 map (hard part 3); ours must not."""
    sd, x, depth, _ = R.reference("vits")
    d = depth.numpy()
    assert d.max() - d.min() > 15.0 and (d < 5).mean() > 0.05 and (d > 15).mean() > 0.05


def test_trunk_taps_match_transformers_dinov2_patch16():
    """The trunk-only oracle (ViT/16 at 384 x 384, raw hooked block outputs + normalised final tokens: the form Depth
    Pro's patch encoder uses) against transformers' independent Dinov2Model with the same weights."""
    pytest.importorskip("transformers")
    from transformers import Dinov2Config, Dinov2Model
    import hf_bridge as H
    torch.manual_seed(2)
    c = O.MODEL_CONFIGS["vits"]
    sd = O.init_state_dict("vits", seed=6, patch=16, pos_grid=24)
    x = torch.randn(2, 3, 384, 384)
    cfg = Dinov2Config(hidden_size=c["embed_dim"], num_hidden_layers=c["depth"], num_attention_heads=c["num_heads"],
                       image_size=384, patch_size=16)
    m = Dinov2Model(cfg).eval()
    hf = {k[len("backbone."):]: v for k, v in H.to_hf({**O.init_state_dict("vits", seed=6), **sd}, "vits").items() if k.startswith("backbone.")}
    m.load_state_dict(hf, strict=True)
    with torch.no_grad():
        out = m(pixel_values=x, output_hidden_states=True)
        taps = O.encoder_taps(sd, x, c, norm_mask=0x8)
    hs = out.hidden_states                      # hs[i + 1] = output of block i (cls included)
    for t, blk in zip(taps[:3], c["taps"][:3]):
        assert float((t - hs[blk + 1][:, 1:]).abs().max()) < 2e-4 * float(t.abs().max())
    assert float((taps[3] - out.last_hidden_state[:, 1:]).abs().max()) < 2e-4 * float(taps[3].abs().max())
