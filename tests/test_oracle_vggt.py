"""Anchors for oracle/vggt_torch.py (parity unpinned against VGGT itself: see its header).  What can be checked:
the position table against the reference's own re-implementation (core/export_compat.py:84-93, when the checkout is
mounted), the rotary embedding against its complex-number form, the block's attention against torch's SDPA, and the
product-side tables (monocular_depth_estimation_trt_b200/vggt.py) against the oracle's."""
import importlib.util
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import vggt_torch as V

REF = "/root/reference/core/export_compat.py"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not mounted")
def test_patch_positions_match_the_reference_export_patch():
    spec = importlib.util.spec_from_file_location("ref_export_compat", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)

    class PositionGetter:                       # the two members the reference's patch touches
        def __init__(self):
            self.position_cache = {}

    with mod.no_cartesian_prod(PositionGetter):
        ref = PositionGetter()(2, 37, 37, torch.device("cpu"))          # [2, 1369, 2] rows of (y, x)
    ours = V.positions(37, 37)
    assert ours.shape == (5 + 1369, 2) and int(ours[:5].abs().max()) == 0
    assert torch.equal(ours[5:] - 1, ref[0]) and torch.equal(ref[0], ref[1])


def test_rope_is_a_rotation_by_position_times_frequency():
    torch.manual_seed(0)
    t = torch.randn(2, 3, 7, 64)
    pos = torch.tensor([[0, 0], [1, 1], [1, 5], [2, 3], [7, 1], [4, 4], [9, 2]])
    out = V.rope_2d(t, pos)
    assert torch.equal(out[:, :, 0], t[:, :, 0])                         # position (0, 0): identity (special tokens)
    f = 1.0 / (100.0 ** (torch.arange(16).float() / 16))                # frequency j of a 32-feature half
    for half, axis in ((slice(0, 32), 0), (slice(32, 64), 1)):
        x = t[..., half]
        z = torch.complex(x[..., :16], x[..., 16:])                     # feature i pairs with i + 16
        rot = z * torch.polar(torch.ones(7, 16), pos[:, axis, None].float() * f)
        assert torch.allclose(out[..., half], torch.cat([rot.real, rot.imag], dim=-1), atol=1e-5)
    assert torch.allclose(out.norm(dim=-1), t.norm(dim=-1), rtol=1e-5)  # rotations preserve length


def test_block_attention_agrees_with_sdpa_and_global_mixes_frames():
    torch.manual_seed(1)
    D, H, S, gh, gw = 128, 2, 3, 2, 3
    sd = V.init_aggregator(D, 1, seed=3)
    N = 5 + gh * gw
    tok = torch.randn(S, N, D)
    pos = V.positions(gh, gw)
    pre = "aggregator.frame_blocks.0."
    out = V.block(sd, pre, tok, pos, H)
    # the same block with torch's fused attention
    y = F.layer_norm(tok, (D,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], 1e-6)
    qkv = F.linear(y, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"]).reshape(S, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    q = V.rope_2d(F.layer_norm(qkv[0], (64,), sd[pre + "attn.q_norm.weight"], sd[pre + "attn.q_norm.bias"]), pos)
    k = V.rope_2d(F.layer_norm(qkv[1], (64,), sd[pre + "attn.k_norm.weight"], sd[pre + "attn.k_norm.bias"]), pos)
    a = F.scaled_dot_product_attention(q, k, qkv[2]).transpose(1, 2).reshape(S, N, D)
    t1 = tok + sd[pre + "ls1.gamma"] * F.linear(a, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
    y = F.layer_norm(t1, (D,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], 1e-6)
    t2 = t1 + sd[pre + "ls2.gamma"] * F.linear(F.gelu(F.linear(y, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])),
                                                sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
    assert torch.allclose(out, t2, atol=2e-5)
    # frame blocks keep frames independent, global blocks do not
    layers = V.aggregate(sd, tok, gh, gw, H, 1)
    tok2 = tok.clone(); tok2[2] += 1.0
    layers2 = V.aggregate(sd, tok2, gh, gw, H, 1)
    assert layers[0].shape == (S, N, 2 * D)
    assert torch.equal(layers[0][:2, :, :D], layers2[0][:2, :, :D])      # frame halves of frames 0, 1 unchanged
    assert not torch.allclose(layers[0][:2, :, D:], layers2[0][:2, :, D:])


def test_causal_aggregator_is_the_streaming_model():
    """StreamVGGT (models/streamvggt/onnx_export.py:35-53): with temporal causal attention a frame does not see its successors
    -- frame i of the causal forward over S frames equals frame i of the forward over the first i + 1 frames (what a key / value
    cache computes frame by frame); with one frame it IS the VGGT aggregator; the last frame sees everything in both."""
    torch.manual_seed(2)
    D, H, S, gh, gw, depth = 128, 2, 3, 2, 3, 2
    sd = V.init_aggregator(D, depth, seed=5)
    tok = torch.randn(S, 5 + gh * gw, D)
    full = V.aggregate(sd, tok, gh, gw, H, depth)
    causal = V.aggregate(sd, tok, gh, gw, H, depth, causal=True)
    for i in range(S):
        prefix = V.aggregate(sd, tok[:i + 1], gh, gw, H, depth, causal=True)
        for layer in range(depth):
            assert torch.allclose(causal[layer][i], prefix[layer][i], atol=1e-5)
    one = V.aggregate(sd, tok[:1], gh, gw, H, depth)
    assert torch.allclose(causal[0][0], one[0][0], atol=1e-5)                      # frame 0 never sees the others
    assert not torch.allclose(causal[1][0], full[1][0], atol=1e-3)                 # ... which it does without the mask
    tok2 = tok.clone(); tok2[2] += 1.0                                              # a later frame changes nothing before it
    causal2 = V.aggregate(sd, tok2, gh, gw, H, depth, causal=True)
    assert torch.equal(causal[1][:2], causal2[1][:2]) and not torch.allclose(causal[1][2], causal2[1][2])


def test_product_tables_equal_the_oracle_tables():
    from monocular_depth_estimation_trt_b200 import vggt as P
    assert torch.equal(torch.from_numpy(P.token_positions(37, 37)).long(), V.positions(37, 37))
    cos, sin = V.rope_tables(39)
    t = P.cos_sin_table(39)
    assert torch.equal(t[:, :16], cos[:, :16]) and torch.equal(t[:, 16:], sin[:, :16]) and torch.equal(cos[:, 16:], cos[:, :16])


# ------------------------------------------------------------------------------------------------ the whole model's pieces
def test_trunk_with_registers_matches_transformers_dinov2_with_registers():
    """VGGT's (and Metric3D V2's) trunk is DINOv2 *with registers*: four learned tokens between cls and the patches, without a
    position embedding, dropped at the output.  transformers' Dinov2WithRegistersModel is an independent implementation."""
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel
    import hf_bridge as HB
    from oracle import dav2_torch as O
    enc = "vits"
    c = O.MODEL_CONFIGS[enc]
    sd = O.init_state_dict(enc, seed=3, registers=4)
    cfg = Dinov2WithRegistersConfig(hidden_size=c["embed_dim"], num_hidden_layers=c["depth"], num_attention_heads=c["num_heads"],
                                    image_size=518, patch_size=14, num_register_tokens=4)
    model = Dinov2WithRegistersModel(cfg).eval()
    hf = {k[len("backbone."):]: v for k, v in HB.to_hf({**HB.init_heads_placeholder(enc), **sd}, enc).items() if k.startswith("backbone.")}
    hf["embeddings.register_tokens"] = sd["pretrained.register_tokens"]
    missing, unexpected = model.load_state_dict(hf, strict=False)
    assert not unexpected and all("mask_token" in m or "pooler" in m for m in missing), (missing, unexpected)
    torch.manual_seed(1)
    x = torch.randn(2, 3, 518, 518)
    with torch.no_grad():
        want = model(pixel_values=x).last_hidden_state            # [2, 1 + 4 + 1369, D], final LayerNorm applied
        cfg1 = dict(c); cfg1["taps"] = [c["depth"] - 1]
        got = O.encoder_taps(sd, x, cfg1, norm_mask=0x1)[0]
    assert want.shape == (2, 1374, c["embed_dim"]) and got.shape == (2, 1369, c["embed_dim"])
    assert float((got - want[:, 5:]).abs().max() / want.abs().max()) < 5e-5


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not mounted")
def test_head_position_embedding_matches_the_reference_export_patch():
    """core/export_compat.py:145-152 is the reference's own float32 restatement of upstream's make_sincos_pos_embed."""
    import sys
    import types
    spec = importlib.util.spec_from_file_location("ref_export_compat2", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    holder = types.ModuleType("fake.heads.utils")
    holder.make_sincos_pos_embed = lambda *a, **k: None
    with mod.float32_sincos_pos_embed(holder):
        ref_fn = holder.make_sincos_pos_embed
        pos = torch.linspace(-0.7, 0.7, 37)
        for dim in (128, 512):
            assert torch.equal(ref_fn(dim, pos), V.make_sincos_pos_embed(dim, pos))
    pe = V.head_pos_embed(256, 37, 37, 518, 518)
    assert pe.shape == (256, 37, 37) and float(pe.abs().max()) <= 0.1 + 1e-6
    # u varies along x only and feeds the first half of the channels, v along y and the second half
    assert torch.equal(pe[:128, 0], pe[:128, 5]) and torch.equal(pe[128:, :, 0], pe[128:, :, 7])
    assert not torch.equal(pe[:128, :, 0], pe[:128, :, 1])


def test_whole_model_shapes_and_special_tokens():
    sd = V.init_vggt("vits", depth=2, features=64, out_channels=(48, 96, 192, 384), seed=0)
    sp = V.special_tokens(sd, 3)
    assert sp.shape == (3, 5, 384) and torch.equal(sp[1], sp[2]) and not torch.equal(sp[0], sp[1])
    torch.manual_seed(0)
    img = torch.rand(3, 3, 70, 84)
    tr = {}
    d = V.vggt_depth(sd, img, "vits", 2, (0, 0, 1, 1), trace=tr)
    assert d.shape == (3, 70, 84) and bool((d > 0).all()) and tr["tokens"].shape == (3, 5 + 30, 384)
    assert tr["aggregated"][0].shape == (3, 35, 768) and tr["logits"].shape == (3, 2, 70, 84)
    # frames interact through the global blocks only: changing frame 2 changes frame 0's depth
    img2 = img.clone(); img2[2] = torch.rand(3, 70, 84)
    assert not torch.allclose(V.vggt_depth(sd, img2, "vits", 2, (0, 0, 1, 1))[0], d[0])
