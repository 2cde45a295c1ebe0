"""Build transformers' DepthAnythingForDepthEstimation with the oracle's weights (test helper).
transformers is an independent implementation that loads the real upstream checkpoints, so agreement
with it is what pins oracle/dav2_torch.py's restatement of the un-vendored upstream module."""
from oracle import dav2_torch as O


def hf_model(encoder, max_depth=20.0):
    from transformers import DepthAnythingConfig, DepthAnythingForDepthEstimation, Dinov2Config
    c = O.MODEL_CONFIGS[encoder]
    bc = Dinov2Config(hidden_size=c["embed_dim"], num_hidden_layers=c["depth"], num_attention_heads=c["num_heads"],
                      image_size=518, patch_size=14, out_indices=[t + 1 for t in c["taps"]],
                      apply_layernorm=True, reshape_hidden_states=False)
    cfg = DepthAnythingConfig(backbone_config=bc, reassemble_hidden_size=c["embed_dim"],
                              neck_hidden_sizes=c["out_channels"], fusion_hidden_size=c["features"],
                              head_hidden_size=32, depth_estimation_type="metric" if max_depth else "relative",
                              max_depth=int(max_depth) if max_depth else 1)
    return DepthAnythingForDepthEstimation(cfg).eval()


def to_hf(sd, encoder):
    c = O.MODEL_CONFIGS[encoder]
    D = c["embed_dim"]
    o = {}
    e = "backbone.embeddings."
    o[e + "cls_token"] = sd["pretrained.cls_token"]
    o[e + "position_embeddings"] = sd["pretrained.pos_embed"]
    o[e + "mask_token"] = sd["pretrained.mask_token"]
    o[e + "patch_embeddings.projection.weight"] = sd["pretrained.patch_embed.proj.weight"]
    o[e + "patch_embeddings.projection.bias"] = sd["pretrained.patch_embed.proj.bias"]
    for i in range(c["depth"]):
        p, q = f"pretrained.blocks.{i}.", f"backbone.encoder.layer.{i}."
        for n in ("norm1", "norm2"):
            o[q + n + ".weight"], o[q + n + ".bias"] = sd[p + n + ".weight"], sd[p + n + ".bias"]
        W, b = sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]
        for j, n in enumerate(("query", "key", "value")):
            o[q + f"attention.attention.{n}.weight"] = W[j * D:(j + 1) * D]
            o[q + f"attention.attention.{n}.bias"] = b[j * D:(j + 1) * D]
        o[q + "attention.output.dense.weight"], o[q + "attention.output.dense.bias"] = sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"]
        o[q + "layer_scale1.lambda1"], o[q + "layer_scale2.lambda1"] = sd[p + "ls1.gamma"], sd[p + "ls2.gamma"]
        for n in ("fc1", "fc2"):
            o[q + f"mlp.{n}.weight"], o[q + f"mlp.{n}.bias"] = sd[p + f"mlp.{n}.weight"], sd[p + f"mlp.{n}.bias"]
    o["backbone.layernorm.weight"], o["backbone.layernorm.bias"] = sd["pretrained.norm.weight"], sd["pretrained.norm.bias"]
    h = "depth_head."
    for i in range(4):
        o[f"neck.reassemble_stage.layers.{i}.projection.weight"] = sd[h + f"projects.{i}.weight"]
        o[f"neck.reassemble_stage.layers.{i}.projection.bias"] = sd[h + f"projects.{i}.bias"]
        if i != 2:
            o[f"neck.reassemble_stage.layers.{i}.resize.weight"] = sd[h + f"resize_layers.{i}.weight"]
            o[f"neck.reassemble_stage.layers.{i}.resize.bias"] = sd[h + f"resize_layers.{i}.bias"]
        o[f"neck.convs.{i}.weight"] = sd[h + f"scratch.layer{i + 1}_rn.weight"]
        r, f = h + f"scratch.refinenet{i + 1}.", f"neck.fusion_stage.layers.{3 - i}."
        o[f + "projection.weight"], o[f + "projection.bias"] = sd[r + "out_conv.weight"], sd[r + "out_conv.bias"]
        for u, hu in (("resConfUnit1", "residual_layer1"), ("resConfUnit2", "residual_layer2")):
            for cv, hc in (("conv1", "convolution1"), ("conv2", "convolution2")):
                o[f + f"{hu}.{hc}.weight"], o[f + f"{hu}.{hc}.bias"] = sd[r + f"{u}.{cv}.weight"], sd[r + f"{u}.{cv}.bias"]
    o["head.conv1.weight"], o["head.conv1.bias"] = sd[h + "scratch.output_conv1.weight"], sd[h + "scratch.output_conv1.bias"]
    o["head.conv2.weight"], o["head.conv2.bias"] = sd[h + "scratch.output_conv2.0.weight"], sd[h + "scratch.output_conv2.0.bias"]
    o["head.conv3.weight"], o["head.conv3.bias"] = sd[h + "scratch.output_conv2.2.weight"], sd[h + "scratch.output_conv2.2.bias"]
    return o
