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


# ---------------------------------------------------------------------------------------------- Depth Pro (whole model)

def depth_pro_hf_model(encoder, features, hook_ids):
    from transformers import DepthProConfig, DepthProForDepthEstimation, Dinov2Config
    c = O.MODEL_CONFIGS[encoder]
    D = c["embed_dim"]
    vit = Dinov2Config(hidden_size=D, num_hidden_layers=c["depth"], num_attention_heads=c["num_heads"], image_size=384, patch_size=16)
    cfg = DepthProConfig(patch_model_config=vit, image_model_config=vit, fov_model_config=vit, intermediate_hook_ids=list(hook_ids),
                         use_fov_model=True, fusion_hidden_size=features, scaled_images_feature_dims=[D, D, D // 2],
                         intermediate_feature_dims=[features, features])
    return DepthProForDepthEstimation(cfg).eval()


def depth_pro_to_hf(sd, encoder):
    """oracle/depth_pro_torch.py `init_full_state_dict` keys (upstream's module tree) -> transformers' DepthProForDepthEstimation."""
    from oracle import depth_pro_torch as DP
    o = {}
    for pre, hf_pre in zip(DP.TRUNKS, ("depth_pro.encoder.patch_encoder.model.", "depth_pro.encoder.image_encoder.model.",
                                       "fov_model.fov_encoder.model.")):
        trunk = DP.trunk_state_dict(sd, pre)
        full = {**init_heads_placeholder(encoder), **trunk}
        for k, v in to_hf(full, encoder).items():
            if k.startswith("backbone."):
                o[hf_pre + k[len("backbone."):]] = v
    up = "depth_pro.neck.feature_upsample."
    for name, hf in (("upsample_latent0", "intermediate.1"), ("upsample_latent1", "intermediate.0"), ("upsample0", "scaled_images.2"),
                     ("upsample1", "scaled_images.1"), ("upsample2", "scaled_images.0")):
        j = 0
        while f"encoder.{name}.{j}.weight" in sd:
            o[up + f"{hf}.layers.{j}.weight"] = sd[f"encoder.{name}.{j}.weight"]
            j += 1
    o[up + "image_block.layers.0.weight"], o[up + "image_block.layers.0.bias"] = sd["encoder.upsample_lowres.weight"], sd["encoder.upsample_lowres.bias"]
    o["depth_pro.neck.fuse_image_with_low_res.weight"] = sd["encoder.fuse_lowres.weight"]
    o["depth_pro.neck.fuse_image_with_low_res.bias"] = sd["encoder.fuse_lowres.bias"]
    for i in range(1, 5):
        o[f"depth_pro.neck.feature_projection.projections.{4 - i}.weight"] = sd[f"decoder.convs.{i}.weight"]
    for i in range(5):
        f, hf = f"decoder.fusions.{i}.", (f"fusion_stage.intermediate.{4 - i}." if i > 0 else "fusion_stage.final.")
        for r, hr in (("resnet1", "residual_layer1"), ("resnet2", "residual_layer2")):
            for j, hc in ((1, "convolution1"), (3, "convolution2")):
                for wb in ("weight", "bias"):
                    if f + f"{r}.residual.{j}.{wb}" in sd:
                        o[hf + f"{hr}.{hc}.{wb}"] = sd[f + f"{r}.residual.{j}.{wb}"]
        if i > 0:
            o[hf + "deconv.weight"] = sd[f + "deconv.weight"]
        o[hf + "projection.weight"], o[hf + "projection.bias"] = sd[f + "out_conv.weight"], sd[f + "out_conv.bias"]
    for j in (0, 1, 2, 4):
        o[f"head.layers.{j}.weight"], o[f"head.layers.{j}.bias"] = sd[f"head.{j}.weight"], sd[f"head.{j}.bias"]
    o["fov_model.fov_encoder.neck.weight"], o["fov_model.fov_encoder.neck.bias"] = sd["fov.encoder.1.weight"], sd["fov.encoder.1.bias"]
    o["fov_model.conv.weight"], o["fov_model.conv.bias"] = sd["fov.downsample.0.weight"], sd["fov.downsample.0.bias"]
    for j in (0, 2, 4):
        o[f"fov_model.head.layers.{j}.weight"], o[f"fov_model.head.layers.{j}.bias"] = sd[f"fov.head.{j}.weight"], sd[f"fov.head.{j}.bias"]
    return o


_PLACEHOLDER = {}


def init_heads_placeholder(encoder):
    """to_hf() walks the DPT head keys too; give it any tensors of the right names (they are dropped afterwards)."""
    if encoder not in _PLACEHOLDER:
        _PLACEHOLDER[encoder] = {k: v for k, v in O.init_state_dict(encoder, seed=0).items() if k.startswith("depth_head.")}
    return _PLACEHOLDER[encoder]
