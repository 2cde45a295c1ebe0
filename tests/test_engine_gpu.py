"""Model-level parity on a B200, through the C-ABI engine and the reference-shaped host API.

Oracle: oracle/dav2_torch.py fp32 forward on the same synthetic input and the same seeded,
calibrated random-init weights (BASELINE.json configs[0]: ViT-S 518x518 batch 1; configs[1]: ViT-L).

ONE gate (north_star): final depth max relative error <= 1e-2 and AbsRel <= 2e-3, for every precision.
  * precision "fp16" (the reference's own build target, models/depth_anything_v2/onnx2trt.py:61, and the default /
    benchmarked precision of this runtime) must meet it.
  * precision "bf16" does not on this oracle, and cannot: rounding only the *weights* to bf16 and running everything else
    in fp32 already gives AbsRel 3.0e-3 / max-rel 2.7e-2 (tests/test_precision_plan.py reproduces this on the CPU).  The
    bf16 cases therefore run the same checks, assert a REGRESSION GUARD (the emulated error of the bf16 plan itself,
    tests/bf16_emulation.py: AbsRel 7.3e-3 / max-rel 7.6e-2, x1.6) so that a broken kernel still fails, and then report
    themselves as xfail against the gate with the measured numbers -- the gate is never widened.
  * intermediate tensors (residual stream after blocks 5 and 11, the four layer_rn maps, path_1)
    are gated in RMS-relative error at 2x what the same emulation measures for each precision.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import refsetup as R
from monocular_depth_estimation_trt_b200 import common, engine as E, weights as W

pytestmark = pytest.mark.gpu

GATE = dict(abs_rel=2e-3, max_rel=1e-2)                      # north_star, every precision
BF16_REGRESSION_GUARD = dict(abs_rel=1.2e-2, max_rel=1.2e-1)   # the bf16 plan's own emulated error x1.6: not a parity claim


def gate(m, prec):
    """Assert north_star's gate; bf16 is reported against the same gate as an expected failure (after its regression guard)."""
    if prec == "bf16":
        assert m["abs_rel"] <= BF16_REGRESSION_GUARD["abs_rel"] and m["max_rel"] <= BF16_REGRESSION_GUARD["max_rel"], m
        if not (m["abs_rel"] <= GATE["abs_rel"] and m["max_rel"] <= GATE["max_rel"]):
            pytest.xfail(f"bf16 operands miss north_star's gate on the fp32 oracle: abs_rel {m['abs_rel']:.2e} (<= 2e-3), "
                         f"max_rel {m['max_rel']:.2e} (<= 1e-2); fp16 is the precision that meets it")
        return
    assert m["abs_rel"] <= GATE["abs_rel"], m
    assert m["max_rel"] <= GATE["max_rel"], m
# RMS-relative budgets: emulation gives 4.3e-4 (fp16) / 3.6e-3 (bf16) on the residual stream and
# 7.5e-4 / 5.9e-3 on the head's maps
INTER = {"fp16": 9e-4, "bf16": 7.5e-3}


def build_engine(encoder, prec, batch=1, input_mode="f32_nchw", max_src_hw=(0, 0), h=518, w=518):
    sd, x, depth, trace = R.reference(encoder, h, w)
    meta = W.describe(encoder, h, w, max_depth=20.0)
    eng = E.Engine(E.make_desc(meta, precision=prec, batch=batch, input_mode=input_mode, max_src_hw=max_src_hw), meta)
    eng.load_state_dict(sd)
    eng.finalize()
    return eng, x, depth, trace


def run(eng, x_dev, out_dev, snapshot=-1):
    ctx = eng.create_execution_context()
    ctx.set_tensor_address("input", x_dev.data_ptr())
    ctx.set_tensor_address("output", out_dev.data_ptr())
    ctx.snapshot_block(snapshot)
    ctx.execute_async_v3(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return ctx


def fetch(ctx, name, shape, prec):
    ptr, nbytes, dt = ctx.get_buffer(name)
    dtype = torch.float32 if dt == 0 else (torch.bfloat16 if prec == "bf16" else torch.float16)
    n = int(np.prod(shape))
    buf = torch.empty(n, dtype=dtype, device="cuda")
    assert n * buf.element_size() <= nbytes
    from cuda.bindings import runtime as cudart
    (err,) = cudart.cudaMemcpy(buf.data_ptr(), ptr, n * buf.element_size(), cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
    assert int(err) == 0
    return buf.reshape(shape).float().cpu()


def rms_rel(got, ref):
    got, ref = got.double(), ref.double()
    return float(((got - ref) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt())


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_vits_518_b1_parity_and_intermediates(lib, prec):
    eng, x, depth, trace = build_engine("vits", prec)
    out = torch.full((1, 518, 518), float("nan"), device="cuda")
    ctx = run(eng, x.cuda(), out, snapshot=5)
    m = R.compare_depth(depth.numpy(), out.cpu().numpy())
    print(prec, m)
    # intermediates first: they localise a failure
    xs = fetch(ctx, "x_snapshot", (1, 1370, 384), prec)
    assert rms_rel(xs, trace["block5"]) < INTER[prec]
    xl = fetch(ctx, "x", (1, 1370, 384), prec)
    assert rms_rel(xl, trace["block11"]) < INTER[prec]
    p1 = fetch(ctx, "path_1", (1, 296, 296, 64), prec).permute(0, 3, 1, 2)
    assert rms_rel(p1, trace["path_1"]) < 1.7 * INTER[prec]
    for i in range(4):
        hh = [148, 74, 37, 19][i]
        r = fetch(ctx, f"r{i}", (1, hh, hh, 64), prec).permute(0, 3, 1, 2)
        assert rms_rel(r, trace[f"layer{i + 1}_rn"]) < 1.7 * INTER[prec]
    assert m["compared"] == 518 * 518
    assert m["corr"] > 0.9995
    gate(m, prec)


def test_embed_tokens_match_oracle(lib):
    """Patch-embed GEMM + cls + pos_embed: token layout and values before any block."""
    eng, x, depth, trace = build_engine("vits", "fp16")
    out = torch.empty(1, 518, 518, device="cuda")
    ctx = eng.create_execution_context()
    ctx.set_tensor_address("input", x.cuda().data_ptr())
    ctx.set_tensor_address("output", out.data_ptr())
    ctx.execute_async_v3(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    cols = fetch(ctx, "cols", (1369, 640), "fp16")
    from oracle import preprocess_np as P
    ref_cols = torch.from_numpy(P.im2col(x.numpy(), 14, 640)).half().float()
    assert torch.equal(cols, ref_cols)          # token/patch layout bit-exact


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_vitl_518_b2_parity(lib, prec):
    """BASELINE configs[1] architecture; batch 2 with two different images."""
    sd, x, depth, _ = R.reference("vitl")
    from oracle import dav2_torch as O, preprocess_np as P
    x2 = torch.from_numpy(P.preprocess_stretch_imagenet(R.synthetic_image(1, 720, 1280), 518, 518))
    xb = torch.cat([x, x2])
    ref = torch.cat([depth, O.forward(sd, x2, "vitl", 20.0)])
    meta = W.describe("vitl", 518, 518, 20.0)
    eng = E.Engine(E.make_desc(meta, precision=prec, batch=2), meta)
    eng.load_state_dict(sd)
    eng.finalize()
    out = torch.full((2, 518, 518), float("nan"), device="cuda")
    run(eng, xb.cuda(), out)
    for b in range(2):
        m = R.compare_depth(ref[b].numpy(), out[b].cpu().numpy())
        print(prec, b, m)
        if prec == "fp16" or b == 1:       # bf16: image 0 runs the regression guard only; the xfail report comes with the last image
            gate(m, prec)
        else:
            assert m["abs_rel"] <= BF16_REGRESSION_GUARD["abs_rel"] and m["max_rel"] <= BF16_REGRESSION_GUARD["max_rel"], m


@pytest.mark.parametrize("h,w", [(616, 1064), (518, 700), (266, 518)])
def test_vits_non_square_token_grids(lib, h, w):
    """Non-square patch grids: 44 x 76 (the Metric3D V2 input size, BASELINE configs[2]; pos_embed resized
    bicubically from 37 x 37 as `interpolate_pos_encoding` does), 37 x 50 (Depth-Anything-AC's keep-ratio sizes)
    and 19 x 37.  Odd grids exercise the ceil'd level-4 map and the `size=` upsampling of refinenet4."""
    eng, x, depth, trace = build_engine("vits", "fp16", h=h, w=w)
    out = torch.full((1, h, w), float("nan"), device="cuda")
    ctx = run(eng, x.cuda(), out)
    gh, gw = h // 14, w // 14
    xl = fetch(ctx, "x", (1, gh * gw + 1, 384), "fp16")
    assert rms_rel(xl, trace["block11"]) < INTER["fp16"]
    m = R.compare_depth(depth.numpy(), out.cpu().numpy())
    print(h, w, m)
    assert m["compared"] == h * w
    assert m["abs_rel"] <= GATE["abs_rel"]
    assert m["max_rel"] <= GATE["max_rel"]


def test_batch_entries_are_independent(lib, bitwise):
    """The same image at batch positions 0 and 2 gives bit-identical maps; idempotent across runs."""
    sd, x, depth, _ = R.reference("vits")
    meta = W.describe("vits", 518, 518, 20.0)
    eng = E.Engine(E.make_desc(meta, precision="fp16", batch=3), meta)
    eng.load_state_dict(sd)
    eng.finalize()
    other = torch.randn(1, 3, 518, 518)
    xb = torch.cat([x, other, x]).cuda()
    out = torch.empty(3, 518, 518, device="cuda")
    ctx = run(eng, xb, out)
    first = out.clone()
    assert torch.equal(out[0], out[2])
    ctx.execute_async_v3(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(out, first)


def test_uint8_input_engine_matches_float_engine(lib, bitwise):
    """Fused resize+normalise+im2col input binding == host preprocessing + float32 binding, bit for bit."""
    sd, x, depth, _ = R.reference("vits")
    img = R.synthetic_image(0)                                 # the image `x` was made from
    meta = W.describe("vits", 518, 518, 20.0)
    e8 = E.Engine(E.make_desc(meta, precision="fp16", batch=1, input_mode="u8_hwc", max_src_hw=(1080, 1920)), meta)
    e8.load_state_dict(sd)
    e8.finalize()
    assert e8.get_tensor_dtype("input") == np.uint8 and e8.get_tensor_shape("input") == (1, 1080, 1920, 3)
    ef, _, _, _ = build_engine("vits", "fp16")
    out8 = torch.empty(1, 518, 518, device="cuda")
    outf = torch.empty(1, 518, 518, device="cuda")
    src = torch.from_numpy(img).cuda()
    c8 = e8.create_execution_context()
    c8.set_input_shape("input", (1, 480, 640, 3))
    c8.set_tensor_address("input", src.data_ptr())
    c8.set_tensor_address("output", out8.data_ptr())
    c8.execute_async_v3(torch.cuda.current_stream().cuda_stream)
    run(ef, x.cuda(), outf)
    torch.cuda.synchronize()
    assert torch.equal(out8, outf)
    with pytest.raises(RuntimeError):
        c8.set_input_shape("input", (1, 4000, 640, 3))


def test_reference_call_shape_do_inference(lib, tmp_path):
    """The four calls every onnx2trt.py makes, end to end with host buffers."""
    sd, x, depth, _ = R.reference("vits")
    path = str(tmp_path / "dav2_vits.mdew")
    W.save(path, sd, W.describe("vits", 518, 518, 20.0))
    with common.get_engine(path, str(tmp_path / "engine" / "dav2_vits_fp16.engine"), "fp16") as engine, \
            engine.create_execution_context() as context:
        inputs, outputs, bindings, stream = common.allocate_buffers(engine, (1, 518, 518), profile_idx=0)
        inputs[0].host = x.numpy()
        timer = common.StageTimer()
        res = common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs,
                                  stream=stream, timer=timer)
        got = res[0].reshape(1, 518, 518).copy()
        assert set(timer.last) == {"h2d_ms", "compute_ms", "d2h_ms"} and timer.last["compute_ms"] > 0
        with pytest.raises(ValueError):
            inputs[0].host = np.zeros(3 * 518 * 518 + 1, np.float32)
        timer.free()
        common.free_buffers(inputs, outputs, stream)
    m = R.compare_depth(depth.numpy(), got)
    assert m["abs_rel"] <= GATE["abs_rel"] and m["max_rel"] <= GATE["max_rel"]
    assert (tmp_path / "engine" / "dav2_vits_fp16.fingerprint").exists()


def test_errors_are_loud(lib):
    meta = W.describe("vits", 518, 518, 20.0)
    eng = E.Engine(E.make_desc(meta, precision="fp16"), meta)
    with pytest.raises(RuntimeError, match="missing weight"):
        eng.finalize()
    with pytest.raises(RuntimeError):
        eng.create_execution_context()
    with pytest.raises(ValueError):
        E.make_desc(meta, precision="fp32")


# ------------------------------------------------------------------ trunk-only engine (Depth Pro's patch-encoder stage)
def _trunk_reference(encoder, batch, seed=5):
    from oracle import dav2_torch as O
    torch.manual_seed(seed)
    x = torch.randn(batch, 3, 384, 384)
    sd = O.init_state_dict(encoder, seed=seed, patch=16, pos_grid=24)
    with torch.no_grad():
        taps = torch.stack(O.encoder_taps(sd, x, O.MODEL_CONFIGS[encoder], norm_mask=0x8))     # [4, B, 576, D]
    return sd, x, taps


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_trunk_only_engine_patch16_taps(lib, prec, bitwise):
    """ViT/16 at 384 x 384 (24 x 24 tokens), raw hooked block outputs + normalised final tokens, 16-bit [4][B][T][D]:
    through the output binding and through the fused-gather path with a single rank (both must agree bit for bit)."""
    from monocular_depth_estimation_trt_b200 import sharding as S
    B = 3
    sd, x, taps = _trunk_reference("vits", B)
    meta = W.describe("vits", 384, 384, max_depth=None, patch_size=16)
    eng = E.Engine(E.make_desc(meta, precision=prec, batch=B, head="encoder_taps", tap_norm_mask=0x8), meta)
    eng.load_state_dict({k: v for k, v in sd.items() if k.startswith("pretrained.")})
    eng.finalize()
    assert eng.get_tensor_shape("output") == (4, B, 576, 384)
    dt = torch.bfloat16 if prec == "bf16" else torch.float16
    out = torch.full((4, B, 576, 384), float("nan"), dtype=dt, device="cuda")
    xd = x.cuda()
    run(eng, xd, out)
    for i in range(4):
        assert rms_rel(out[i].float().cpu(), taps[i]) < INTER[prec], i
    enc = S.ShardedPatchEncoder(eng, n_items=B, world=1, rank=0, mode="fused")
    enc.enqueue(xd.data_ptr(), torch.cuda.current_stream().cuda_stream)
    enc.finish()
    assert torch.equal(enc.gathered(), out)
    enc.close()
    eng.close()


def test_source_grid_output_fuses_the_scripts_postprocessing(lib, bitwise):
    """output="source_grid": the engine's binding is the depth map resized back to the source frame and clamped
    (onnx2trt.py:111-117), equal to doing that on the host from the model-grid output of the same engine."""
    import torch.nn.functional as F
    sd, x, depth, _ = R.reference("vits")
    meta = W.describe("vits", 518, 518, 20.0)
    frames = np.stack([R.synthetic_image(i, 480, 640) for i in range(2)])
    outs = {}
    for mode in ("model_grid", "source_grid"):
        eng = E.Engine(E.make_desc(meta, precision="fp16", batch=2, input_mode="u8_hwc", max_src_hw=(480, 640), output=mode), meta)
        eng.load_state_dict(sd)
        eng.finalize()
        shape = eng.get_tensor_shape("output")
        assert shape == ((2, 518, 518) if mode == "model_grid" else (2, 480, 640))
        out = torch.full(shape, float("nan"), device="cuda")
        run(eng, torch.from_numpy(frames).cuda(), out)
        outs[mode] = out.cpu()
        eng.close()
    ref = torch.clamp(F.interpolate(outs["model_grid"][:, None], (480, 640), mode="bilinear", align_corners=True)[:, 0], 1e-3, 1e3)
    assert float((outs["source_grid"] - ref).abs().max()) <= 2e-5 * float(ref.abs().max())


def test_graph_replay_matches_plain_launches(lib, bitwise):
    """On a capturable stream the launch sequence is recorded into a CUDA graph at the first enqueue and replayed
    afterwards; new bindings re-record.  Every variant must reproduce the plain-launch result bit for bit."""
    eng, x, depth, _ = build_engine("vits", "fp16")
    xd = x.cuda()
    ref = torch.full((1, 518, 518), float("nan"), device="cuda")
    run(eng, xd, ref)                                     # legacy default stream: plain launches
    side = torch.cuda.Stream()
    ctx = eng.create_execution_context()
    out_a = torch.full_like(ref, float("nan"))
    out_b = torch.full_like(ref, float("nan"))
    ctx.set_tensor_address("input", xd.data_ptr())
    side.wait_stream(torch.cuda.current_stream())
    for out in (out_a, out_a, out_b, out_b):              # capture, replay, re-capture for the new binding, replay
        ctx.set_tensor_address("output", out.data_ptr())
        out.fill_(float("nan"))
        side.wait_stream(torch.cuda.current_stream())
        ctx.execute_async_v3(side.cuda_stream)
        side.synchronize()
        assert torch.equal(out, ref)
    ctx.close()
    eng.close()


def test_split_k_changes_only_the_last_bits(lib):
    """Batch 1 with split_k=True (MDE_FLAG_SPLIT_K in the engine description): FC2 / projection split K over the SMs and meet in the L2's fp32 adds (arrival order).  Against the
    unsplit engine the depth map moves the way any re-association of fp32 sums moves a 16-bit pipeline (some 16-bit
    roundings flip downstream); both stay inside the parity gate on the oracle."""
    sd, x, depth, _ = R.reference("vits")
    outs = []
    for split in (True, False):
        meta = W.describe("vits", 518, 518, 20.0)
        eng = E.Engine(E.make_desc(meta, precision="fp16", batch=1, split_k=split), meta)
        eng.load_state_dict(sd)
        eng.finalize()
        out = torch.full((1, 518, 518), float("nan"), device="cuda")
        run(eng, x.cuda(), out)
        outs.append(out.cpu())
        eng.close()
    m = R.compare_depth(outs[1].numpy(), outs[0].numpy())
    print("split-K vs unsplit:", m)
    # fp32 re-association upstream flips 16-bit roundings downstream, and twelve blocks later the two runs carry two
    # independent realisations of the pipeline's rounding noise: they differ from each other by about as much as each
    # differs from the fp32 oracle (measured AbsRel 7.0e-4 / max-rel 9.5e-3), i.e. inside north_star's gate, not bitwise
    assert m["abs_rel"] < GATE["abs_rel"] and m["max_rel"] < 1.5 * GATE["max_rel"]
    for o in outs:
        g = R.compare_depth(depth.numpy(), o.numpy())
        assert g["abs_rel"] <= GATE["abs_rel"] and g["max_rel"] <= GATE["max_rel"]


def test_vitb_relative_head_parity(lib):
    """The middle encoder size (D=768, 12 heads: 2304 / 3072-wide GEMMs, 128-wide DPT maps) with the RELATIVE head
    (models/depth_anything_v2/infer.py: no max_depth -> ReLU instead of sigmoid * max_depth)."""
    from oracle import dav2_torch as O, preprocess_np as P
    x = torch.from_numpy(P.preprocess_stretch_imagenet(R.synthetic_image(0), 518, 518))
    sd = O.init_state_dict("vitb", seed=0)
    O.calibrate_head(sd, x, "vitb")
    sd["depth_head.scratch.output_conv2.2.bias"] = sd["depth_head.scratch.output_conv2.2.bias"] + 6.0     # logits ~ N(6, 1): the whole map sits above the ReLU, relative errors stay meaningful
    depth = O.forward(sd, x, "vitb", max_depth=None)
    meta = W.describe("vitb", 518, 518, max_depth=None)
    eng = E.Engine(E.make_desc(meta, precision="fp16", batch=1), meta)
    eng.load_state_dict(sd)
    eng.finalize()
    out = torch.full((1, 518, 518), float("nan"), device="cuda")
    run(eng, x.cuda(), out)
    got = out.cpu().numpy()
    assert float(got.min()) >= 0.0
    m = R.compare_depth(depth.numpy(), got)
    print("vitb relative", m)
    assert m["abs_rel"] <= GATE["abs_rel"] and m["max_rel"] <= 2 * GATE["max_rel"]
    assert m["corr"] > 0.9995


def test_depth_pro_patch_encoder_stage_end_to_end(lib, bitwise):
    """Depth Pro's patch-encoder stage on one GPU: 35 crops of a 1536 x 1536 image -> trunk-only ViT/16 engine (fused
    gather path, world 1) -> patch merge kernel -> the five feature maps.  Checked against the oracle
    (oracle/depth_pro_torch.py, pinned on transformers' DepthProPatchEncoder); the merge itself is a copy and must be
    bit-exact with the oracle's merge applied to the same taps."""
    from oracle import dav2_torch as O, depth_pro_torch as DP
    from monocular_depth_estimation_trt_b200 import sharding as S
    torch.manual_seed(9)
    image = torch.randn(3, 1536, 1536)
    sd = O.init_state_dict("vits", seed=8, patch=16, pos_grid=24)
    ref = DP.patch_encoder_features(sd, image, "vits", hook_taps=(2, 1))
    crops = S.make_crops(image)
    meta = W.describe("vits", 384, 384, max_depth=None, patch_size=16)
    eng = E.Engine(E.make_desc(meta, precision="fp16", batch=35, head="encoder_taps", tap_norm_mask=0x8), meta)
    eng.load_state_dict({k: v for k, v in sd.items() if k.startswith("pretrained.")})
    eng.finalize()
    enc = S.ShardedPatchEncoder(eng, n_items=35, world=1, rank=0, mode="fused")
    stream = torch.cuda.current_stream().cuda_stream
    xd = crops.cuda()
    enc.enqueue(xd.data_ptr(), stream)
    enc.finish()
    taps = enc.gathered()
    maps = S.merge_features(taps, "fp16", stream, hook_taps=(2, 1))
    torch.cuda.synchronize()
    same = DP.merged_features([t.cpu() for t in taps], hook_taps=(2, 1))
    assert [tuple(m.shape) for m in maps] == [(24, 24, 384), (48, 48, 384), (96, 96, 384), (96, 96, 384), (96, 96, 384)]
    for m, s_, r in zip(maps, same, ref):
        assert torch.equal(m.cpu(), s_)                              # the merge moves tokens, nothing else
        assert rms_rel(m.float().cpu(), r) < INTER["fp16"]
    enc.close()
    eng.close()


def test_profiler_surface_like_core_profile(lib, tmp_path, bitwise):
    """tools/profile_model.py:120-135 attaches `context.profiler = LayerTimer()` and calls do_inference; core/profile.py
    then builds rows {name, ms, calls_per_iter, share} and reads the engine's layer list.  The same flow here."""
    from collections import OrderedDict

    class LayerTimer:                         # core/profile.py:30-58, minus the tensorrt base class
        def __init__(self):
            self.total_ms, self.calls = OrderedDict(), OrderedDict()

        def report_layer_time(self, layer_name, ms):
            self.total_ms[layer_name] = self.total_ms.get(layer_name, 0.0) + ms
            self.calls[layer_name] = self.calls.get(layer_name, 0) + 1

    sd, x, depth, _ = R.reference("vits")
    path = str(tmp_path / "dav2_vits.mdew")
    W.save(path, sd, W.describe("vits", 518, 518, 20.0))
    with common.get_engine(path, "", "fp16") as engine, engine.create_execution_context() as context:
        inputs, outputs, bindings, stream = common.allocate_buffers(engine, (1, 518, 518), profile_idx=0)
        inputs[0].host = x.numpy()
        plain = common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)[0].copy()
        timer = LayerTimer()
        context.profiler = timer
        for _ in range(2):
            prof = common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)[0].copy()
        layers = context.layer_information()
        common.free_buffers(inputs, outputs, stream)
    assert np.array_equal(plain, prof)                                   # profiling changes the schedule, not the result
    assert sum(timer.calls.values()) == 2 * len(layers) and all(v > 0 for v in timer.total_ms.values())
    names = {l["Name"] for l in layers}
    assert set(timer.total_ms) == names and any(n.startswith("attention") for n in names)
    assert {l["LayerType"] for l in layers} >= {"MatrixMultiply", "Convolution", "Attention", "Normalization"}
    total_flops = sum(l["AlgorithmicFlops"] for l in layers)
    assert abs(total_flops / 115.3e9 - 1.0) < 0.12     # SURVEY section 8 d: 115.3 GFLOP per ViT-S image (the plan commutes two 1x1 / 3x3 maps with upsamplings)


def test_depth_anything_v3_exp_and_sky_heads(lib):
    """MDE_HEAD_DPT_EXP_SKY (models/depth_anything_v3/onnx_export.py:31-55: outputs `depth`, `sky`): the DPT head ending in
    exp(), and the parallel sky branch on the same up-sampled map, against the oracle's restatement of the profile's layer list."""
    from oracle import dav2_torch as O
    import torch.nn.functional as F
    sd, x, _, _ = R.reference("vits")
    sd = dict(sd)
    O.add_sky_branch(sd, "vits", seed=0)
    # calibrate the sky logit to straddle zero (median 0, unit spread) so that the final ReLU is exercised on both sides
    tr = {}
    O.forward_da3(sd, x, "vits", tr)
    h = "depth_head.scratch."
    up = F.interpolate(F.conv2d(tr["path_1"], sd[h + "output_conv1.weight"], sd[h + "output_conv1.bias"], padding=1), (518, 518),
                       mode="bilinear", align_corners=True)
    pre = F.conv2d(F.relu(F.conv2d(up, sd[h + "sky_output_conv2.0.weight"], sd[h + "sky_output_conv2.0.bias"], padding=1)),
                   sd[h + "sky_output_conv2.2.weight"], sd[h + "sky_output_conv2.2.bias"])
    sd[h + "sky_output_conv2.2.weight"] = sd[h + "sky_output_conv2.2.weight"] / float(pre.std())
    sd[h + "sky_output_conv2.2.bias"] = (sd[h + "sky_output_conv2.2.bias"] - float(pre.median())) / float(pre.std())
    depth_ref, sky_ref = O.forward_da3(sd, x, "vits")
    assert 0.3 < float((sky_ref > 0).float().mean()) < 0.7
    meta = W.describe("vits", 518, 518, None)
    eng = E.Engine(E.make_desc(meta, precision="fp16", batch=1, head="dpt_exp_sky"), meta)
    eng.load_state_dict(sd)
    eng.finalize()
    assert [eng.get_tensor_name(i) for i in range(eng.num_io_tensors)] == ["image", "depth", "sky"]
    assert eng.get_tensor_shape("depth") == (1, 518, 518) == eng.get_tensor_shape("sky")
    inputs, outputs, bindings, stream = common.allocate_buffers(eng)
    inputs[0].host = x.numpy()
    with eng.create_execution_context() as ctx:
        with pytest.raises(RuntimeError):
            ctx.set_tensor_address("output", bindings[1])                  # this engine's bindings are image / depth / sky
        depth, sky = common.do_inference(ctx, engine=eng, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)
        depth, sky = depth.reshape(518, 518).copy(), sky.reshape(518, 518).copy()
    common.free_buffers(inputs, outputs, stream)
    eng.close()
    m = R.compare_depth(depth_ref[0].numpy(), depth)
    print("da3 depth", m)
    assert m["abs_rel"] <= GATE["abs_rel"] and m["max_rel"] <= GATE["max_rel"], m
    # the sky map is a rectified logit: absolute error against the logit's unit spread (a relative measure is meaningless at 0)
    err = np.abs(sky - sky_ref[0].numpy())
    print("da3 sky max abs err", float(err.max()), "mean", float(err.mean()))
    assert float(err.max()) <= 1e-2 and float(err.mean()) <= 1e-3 and (sky >= 0).all()


def test_depth_anything_ac_input_contract_bit_exact(lib):
    """The fused uint8 input binding with scale_dtype="float32" (MDE_FLAG_SCALE_F32): the patch rows the trunk sees equal
    im2col of core/preprocess.py's depth_anything_ac tensor bit for bit, and differ from depth_anything_v2's in the last bits.
    Also the model's native, non-square size from its keep-ratio rule (4:3 -> 518 x 700) builds and runs."""
    from oracle import preprocess_np as PP
    sd, _, _, _ = R.reference("vits")
    img = R.synthetic_image(3, 480, 640)
    cols = {}
    for sdt in ("float32", "float64"):
        meta = W.describe("vits", 518, 518, 20.0)
        eng = E.Engine(E.make_desc(meta, precision="fp16", batch=1, input_mode="u8_hwc", max_src_hw=(480, 640), scale_dtype=sdt), meta)
        eng.load_state_dict(sd)
        eng.finalize()
        out = torch.empty(1, 518, 518, device="cuda")
        src = torch.from_numpy(img).cuda()
        ctx = eng.create_execution_context()
        ctx.set_input_shape("input", (1, 480, 640, 3))
        ctx.set_tensor_address("input", src.data_ptr())
        ctx.set_tensor_address("output", out.data_ptr())
        ctx.execute_async_v3(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        cols[sdt] = fetch(ctx, "cols", (1369, 640), "fp16")
        ref = torch.from_numpy(PP.im2col(PP.preprocess_stretch_imagenet(img, 518, 518, sdt), 14, 640)).half().float()
        assert torch.equal(cols[sdt], ref), sdt
        ctx.close(); eng.close()
    h, w = W.keep_ratio_size(480, 640, 518, 14, "ceil")
    assert (h, w) == (518, 700)
    meta = W.describe("vits", h, w, None)
    eng = E.Engine(E.make_desc(meta, precision="fp16", batch=1, input_mode="u8_hwc", max_src_hw=(480, 640), scale_dtype="float32"), meta)
    eng.load_state_dict(sd)
    eng.finalize()
    out = torch.full((1, h, w), float("nan"), device="cuda")
    with eng.create_execution_context() as ctx:
        ctx.set_input_shape("input", (1, 480, 640, 3))
        ctx.set_tensor_address("input", torch.from_numpy(img).cuda().data_ptr())
        ctx.set_tensor_address("output", out.data_ptr())
        ctx.execute_async_v3(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    eng.close()


@pytest.mark.parametrize("src", [(480, 640), (769, 1025), (500, 500), (37, 53), (1036, 720)])
def test_depth_anything_ac_native_preprocessing_kernel_is_bit_exact(lib, src):
    """Depth-Anything-AC's `native` profile on the device (`mde_k_preprocess_u8_cubic_f32`): uint8 frames -> float32 / 255 ->
    cv2's float INTER_CUBIC to the keep-ratio "ceil" size -> ImageNet statistics in float64, against the oracle (byte-exact with
    the reference module's tensors and with cv2's own cubic path), two frames per call, every bit; also a small target and the
    plain resized image without statistics."""
    import kutil as K
    from oracle import preprocess_np as PP
    frames = np.stack([np.random.default_rng(s).integers(0, 256, (*src, 3), dtype=np.uint8) for s in (5, 6)])
    d = torch.from_numpy(frames).cuda()
    for target in (518, 56):
        h, w = PP.keep_ratio_size(*src, target, 14, "ceil")
        ref = np.concatenate([PP.preprocess_keep_ratio_cubic_f32(f, target) for f in frames])
        got = K.preprocess_u8_cubic_f32(d, h, w)
        torch.cuda.synchronize()
        assert got.shape == ref.shape and np.array_equal(got.cpu().numpy(), ref), (src, target)
    h, w = PP.keep_ratio_size(*src, 56, 14, "ceil")
    plain = K.preprocess_u8_cubic_f32(d, h, w, mean=None, std=None)
    x0 = np.ascontiguousarray(frames[0][:, :, ::-1].astype(np.float32) / np.float32(255.0))
    assert np.array_equal(plain[0].cpu().numpy(), PP.resize_cubic_f32(x0, h, w).transpose(2, 0, 1))


def test_depth_anything_ac_native_profile_end_to_end(lib):
    """models/depth_anything_ac/onnx2trt.py with `profile = 'native'`: the frame is preprocessed on the device at its own
    aspect (480 x 640 -> 518 x 700), the engine built for that size takes the float32 tensor, and the map goes through the gate
    against the oracle's forward of the reference-exact tensor."""
    import kutil as K
    from oracle import dav2_torch as O, preprocess_np as PP
    sd, _, _, _ = R.reference("vits")
    img = R.synthetic_image(4, 480, 640)
    x_ref = torch.from_numpy(PP.preprocess_keep_ratio_cubic_f32(img))
    h, w = x_ref.shape[2:]
    assert (h, w) == (518, 700)
    want = O.forward(sd, x_ref, "vits", max_depth=20.0)[0].numpy()
    meta = W.describe("vits", h, w, 20.0)
    eng = E.Engine(E.make_desc(meta, precision="fp16", batch=1), meta)
    eng.load_state_dict(sd)
    eng.finalize()
    x_dev = K.preprocess_u8_cubic_f32(torch.from_numpy(img[None]).cuda(), h, w)
    assert torch.equal(x_dev.cpu(), x_ref)
    out = torch.full((1, h, w), float("nan"), device="cuda")
    with eng.create_execution_context() as ctx:
        ctx.set_tensor_address("input", x_dev.data_ptr())
        ctx.set_tensor_address("output", out.data_ptr())
        ctx.execute_async_v3(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    eng.close()
    m = R.compare_depth(want, out[0].cpu().numpy())
    print("depth_anything_ac native 518x700 fp16", m["abs_rel"], m["max_rel"])
    assert m["abs_rel"] <= GATE["abs_rel"] and m["max_rel"] <= GATE["max_rel"], m


def test_metric3d_v2_trunk_with_registers_and_bilinear_pos_embed(lib):
    """Metric3D V2's encoder as the reference exports it (reports/profile/metric3d_v2.json layers 0-131): 0..255 input with the
    in-graph (x - mean) / std applied where the patch rows are formed (MDE_FLAG_NORMALISE_F32), DINOv2 with four registers (1 + 4 + 44 * 76 = 3349 tokens, layer
    10), position embedding resized to 44 x 76 with a half-pixel BILINEAR resize (layer 6), final LayerNorm (layer 131).
    Trunk-only engine, batch 2, against the oracle's trunk (pinned on transformers' Dinov2WithRegistersModel).  The RAFT-style
    decoder behind it is not built."""
    from oracle import dav2_torch as O, preprocess_np as PP
    enc, H, Wd = "vits", 616, 1064
    cfg = dict(O.MODEL_CONFIGS[enc])
    L = cfg["depth"]
    cfg["taps"] = [L - 4, L - 3, L - 2, L - 1]
    sd = {k: v for k, v in O.init_state_dict(enc, seed=9, registers=4).items() if k.startswith("pretrained.")}
    x = torch.cat([torch.from_numpy(PP.preprocess_pad_none(R.synthetic_image(i, 480, 640), H, Wd)) for i in range(2)])   # 0..255, padded
    mean, std = torch.tensor([123.675, 116.28, 103.53]).view(1, 3, 1, 1), torch.tensor([58.395, 57.12, 57.375]).view(1, 3, 1, 1)
    taps = O.encoder_taps(sd, (x - mean) / std, cfg, norm_mask=0x8, pos_interp="bilinear")
    meta = W.describe(enc, H, Wd, None)
    meta.update(taps=cfg["taps"], registers=4, pos_interp="bilinear")
    eng = E.Engine(E.make_desc(meta, precision="fp16", batch=2, head="encoder_taps", tap_norm_mask=0x8, normalise_f32=True,
                               mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375)), meta)
    eng.load_state_dict(sd)
    eng.finalize()
    T = 44 * 76
    out = torch.zeros(4, 2, T, 384, dtype=torch.float16, device="cuda")
    with eng.create_execution_context() as ctx:
        ctx.set_tensor_address("input", x.cuda().data_ptr())
        ctx.set_tensor_address("output", out.data_ptr())
        ctx.execute_async_v3(torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        xres = fetch(ctx, "x", (2, 1 + 4 + T, 384), "fp16")
    eng.close()
    for i in range(4):
        assert rms_rel(out[i].float().cpu(), taps[i]) < INTER["fp16"], i
    assert xres.shape[1] == 3349
