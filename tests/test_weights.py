"""The export artefact (.mdew) and the pos-embed rule (host side of `run.py export` / `build`)."""
import numpy as np
import pytest
import torch

from monocular_depth_estimation_trt_b200 import engine as E, weights as W
from oracle import dav2_torch as O


def test_mdew_roundtrip_and_native_loader(tmp_path, lib):
    sd = O.init_state_dict("vits", seed=3)
    meta = W.describe("vits", 518, 518, 20.0)
    path = str(tmp_path / "m.mdew")
    sha = W.save(path, sd, meta)
    assert sha == W.file_sha256(path)
    back, meta2 = W.load(path)
    assert meta2 == meta and set(back) == set(sd)
    for k in sd:
        assert np.array_equal(back[k], sd[k].numpy())
    assert W.read_meta(path)["encoder"] == "vits"
    eng = E.Engine(E.make_desc(meta), meta)
    eng.load_weights_file(path)               # the C loader parses the same file (host only; no GPU needed)
    eng.close()
    bad = tmp_path / "bad.mdew"
    bad.write_bytes(b"NOTMDEW0" + b"\0" * 64)
    eng = E.Engine(E.make_desc(meta), meta)
    with pytest.raises(RuntimeError, match="MDEW0001"):
        eng.load_weights_file(str(bad))
    with pytest.raises(RuntimeError):
        eng.load_weights_file(str(tmp_path / "missing.mdew"))
    trunc = tmp_path / "trunc.mdew"
    trunc.write_bytes(open(path, "rb").read()[:4000])
    with pytest.raises(RuntimeError):
        eng.load_weights_file(str(trunc))
    eng.close()


def test_describe_validates():
    with pytest.raises(KeyError):
        W.describe("vitg")
    with pytest.raises(ValueError):
        W.describe("vits", 520, 518)
    m = W.describe("vitl", 616, 1064, None)
    assert m["embed_dim"] == 1024 and m["taps"] == [4, 11, 17, 23] and m["max_depth"] is None


def test_pos_embed_rule_matches_oracle():
    pe = torch.randn(1, 1370, 64)
    assert np.array_equal(W.resize_pos_embed(pe.numpy(), 37, 37), pe.numpy())          # trained grid: identity
    for gh, gw in [(44, 76), (20, 30), (37, 50)]:
        got = W.resize_pos_embed(pe.numpy(), gh, gw)
        ref = O.interpolate_pos_embed(pe, gh, gw).numpy()
        assert got.shape == (1, 1 + gh * gw, 64) and np.allclose(got, ref, atol=1e-6)
        assert np.array_equal(got[:, 0], pe[:, 0].numpy())                              # cls part untouched


def test_export_from_an_upstream_style_checkpoint(tmp_path, lib):
    """A .pth with upstream's key names (here: the oracle's seeded init saved with torch.save, plus the kind of
    wrapper keys checkpoints carry) -> .mdew -> engine description, without naming the encoder."""
    sd = O.init_state_dict("vits", seed=4)
    ck = tmp_path / "depth_anything_v2_metric_hypersim_vits.pth"
    torch.save({"module." + k: v for k, v in sd.items()}, ck)
    out = str(tmp_path / "m.mdew")
    meta = W.export_checkpoint(str(ck), out, 518, 700, max_depth=20.0)
    assert meta["encoder"] == "vits" and meta["embed_dim"] == 384 and (meta["input_h"], meta["input_w"]) == (518, 700)
    assert meta["source_checkpoint_sha256"] == W.file_sha256(str(ck))
    back, meta2 = W.load(out)
    assert set(back) == set(sd) and meta2["max_depth"] == 20.0
    eng = E.Engine(E.make_desc(meta2), meta2)
    eng.load_weights_file(out)
    eng.close()
    assert W.encoder_of(O.init_state_dict("vitb", seed=1)) == "vitb"
    with pytest.raises(ValueError):
        W.encoder_of({"foo": torch.zeros(1)})


def test_export_from_a_transformers_checkpoint(tmp_path, lib):
    """SURVEY 8 f-2 on a REAL key set: transformers' DepthAnythingForDepthEstimation (what the hub's `*-hf` repositories
    hold) with its own random init, saved the way `save_pretrained` does (model.safetensors) -> `export_checkpoint` ->
    .mdew under upstream's key names -> the oracle's forward of THOSE tensors equals transformers' forward of the model
    they came from (the key mapping, the q|k|v stacking and the refinenet order are pinned by an independent
    implementation), and the engine description / native loader accept the file."""
    pytest.importorskip("transformers")
    from safetensors.torch import save_file
    import hf_bridge as H
    torch.manual_seed(11)
    m = H.hf_model("vits", 20.0)
    with torch.no_grad():                      # a fresh init leaves LayerScale at 1 and biases at 0: make every tensor matter
        for k, v in m.state_dict().items():
            if v.dtype.is_floating_point and ("lambda1" in k or k.endswith(".bias")):
                v.add_(0.05 * torch.randn_like(v))
    hf_sd = {k: v.contiguous() for k, v in m.state_dict().items()}
    ck = tmp_path / "model.safetensors"
    save_file(hf_sd, str(ck))
    out = str(tmp_path / "hf.mdew")
    meta = W.export_checkpoint(str(ck), out, 518, 518, max_depth=20.0)
    assert meta["encoder"] == "vits" and meta["max_depth"] == 20.0
    back, meta2 = W.load(out)
    assert set(back) - {"pretrained.mask_token"} == set(O.init_state_dict("vits", seed=0)) - {"pretrained.mask_token"}
    sd = {k: torch.from_numpy(v) for k, v in back.items()}
    x = torch.randn(1, 3, 518, 518, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        want = m(pixel_values=x).predicted_depth
    got = O.forward(sd, x, "vits", max_depth=20.0)
    rel = ((got - want).abs() / want.abs().clamp_min(1e-6)).max().item()
    assert rel < 5e-5, rel
    # the inverse of the test helper that pins the oracle: to_hf(from_hf(x)) is the identity on the tensors the model uses
    again = H.to_hf(W.from_hf_depth_anything(hf_sd), "vits")
    assert all(torch.equal(again[k], hf_sd[k]) for k in again) and set(again) <= set(hf_sd)
    eng = E.Engine(E.make_desc(meta2), meta2)
    eng.load_weights_file(out)
    eng.close()
    with pytest.raises(ValueError):
        W.from_hf_depth_anything({"backbone.embeddings.cls_token": torch.zeros(1, 1, 384),
                                  "backbone.embeddings.position_embeddings": torch.zeros(1, 1370, 384),
                                  "backbone.embeddings.patch_embeddings.projection.weight": torch.zeros(1),
                                  "backbone.embeddings.patch_embeddings.projection.bias": torch.zeros(1)})
