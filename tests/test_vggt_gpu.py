"""VGGT's aggregator blocks on a B200 against oracle/vggt_torch.py: the qk-norm + 2-D RoPE kernel on its own (fp32
reference on the same 16-bit inputs), then frame / global blocks and the alternating stack at the real 37 x 37 token grid.
Budgets as for the Depth Anything residual stream (tests/test_engine_gpu.py INTER), per block pair."""
import functools

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from monocular_depth_estimation_trt_b200 import vggt as P
from oracle import vggt_torch as V

pytestmark = pytest.mark.gpu

DT = {"fp16": torch.float16, "bf16": torch.bfloat16}
ULP = {"fp16": 2 ** -10, "bf16": 2 ** -7}
INTER = {"fp16": 1.2e-3, "bf16": 9e-3}


def rms_rel(got, ref):
    got, ref = got.double(), ref.double()
    return float(((got - ref) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt())


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("heads,with_pos", [(6, True), (16, True), (2, False)])
def test_qknorm_rope_kernel(lib, prec, heads, with_pos):
    from monocular_depth_estimation_trt_b200.depth_pro import _Ops
    import ctypes as C
    torch.manual_seed(heads)
    gh, gw = 5, 9
    pos = V.positions(gh, gw)
    rows, D = 3 * pos.shape[0] + 1, heads * 64                          # a ragged last CTA
    posr = torch.cat([pos.repeat(3, 1), pos[-1:]])
    qkv = (torch.randn(rows, 3 * D) * 2 + 0.3).to(DT[prec])
    w = [1 + 0.1 * torch.randn(64), 0.1 * torch.randn(64), 1 + 0.1 * torch.randn(64), 0.1 * torch.randn(64)]
    # fp32 reference on the same 16-bit inputs
    r = qkv.float().reshape(rows, 3, heads, 64)
    q = F.layer_norm(r[:, 0], (64,), w[0], w[1], 1e-5).permute(1, 0, 2)[None]
    k = F.layer_norm(r[:, 1], (64,), w[2], w[3], 1e-5).permute(1, 0, 2)[None]
    if with_pos:
        q, k = V.rope_2d(q, posr), V.rope_2d(k, posr)
    ref = torch.stack([q[0].permute(1, 0, 2), k[0].permute(1, 0, 2), r[:, 2]], dim=1).reshape(rows, 3 * D)
    ops = _Ops(prec)
    ops.stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    d = qkv.cuda()
    gathered = torch.full((rows + 2, 2 * D + 8), 7.0, dtype=DT[prec], device="cuda")
    table = P.cos_sin_table(int(pos.max()) + 1).cuda()
    ops.qknorm_rope(d, rows, heads, *[t.cuda() for t in w], 1e-5, posr.int().cuda() if with_pos else None, table if with_pos else None,
                    table.shape[0], gather=[gathered.data_ptr()], gather_ld=2 * D + 8)
    torch.cuda.synchronize()
    got = d.float().cpu()
    assert torch.equal(got[:, 2 * D:], qkv.float()[:, 2 * D:])           # V untouched
    assert float((got - ref).abs().max()) <= 1.01 * ULP[prec] * float(ref.abs().max())
    g = gathered.cpu()
    assert torch.equal(g[:rows, :2 * D], d.cpu()[:, D:]) and bool((g[rows:] == 7.0).all()) and bool((g[:, 2 * D:] == 7.0).all())


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_aggregator_against_the_oracle(lib, prec):
    """Two layers of (frame block, global block), 3 frames of 5 + 37*37 tokens, ViT-S width: the taps [frame | global] of
    both layers against the fp32 oracle."""
    torch.manual_seed(5)
    D, H, S, depth, gh, gw = 384, 6, 3, 2, 37, 37
    sd = V.init_aggregator(D, depth, seed=4)
    N = 5 + gh * gw
    tok = torch.randn(S, N, D)
    ref = V.aggregate(sd, tok, gh, gw, H, depth)
    agg = P.Aggregator(sd, D, depth, H, gh, gw, frames_total=S, precision=prec, taps=(0, 1))
    x = tok.cuda()
    stream = torch.cuda.current_stream().cuda_stream
    agg.forward(x.data_ptr(), stream)
    torch.cuda.synchronize()
    first = {t: agg.tap_out[t].clone() for t in (0, 1)}
    agg.forward(x.data_ptr(), stream)
    torch.cuda.synchronize()
    for t in (0, 1):
        got = agg.tap_out[t].cpu().reshape(S, N, 2 * D)
        assert torch.equal(agg.tap_out[t], first[t])                     # reproducible bit for bit
        e_frame, e_global = rms_rel(got[..., :D], ref[t][..., :D]), rms_rel(got[..., D:], ref[t][..., D:])
        print(prec, "layer", t, "frame", e_frame, "global", e_global)
        assert e_frame < INTER[prec] * (t + 1) and e_global < INTER[prec] * (t + 1)
    assert agg.ops.launches == depth * 2 * 8
    agg.close()


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_causal_aggregator_and_its_streaming_form(lib, prec):
    """StreamVGGT: (a) the causal aggregator over S frames against the oracle's masked forward; (b) the same frames fed ONE AT A
    TIME through an aggregator with a key / value cache: frame t of the stream is frame t of (a), bit for bit (same kernels,
    same key order; only where the keys live differs)."""
    torch.manual_seed(7)
    D, H, S, depth, gh, gw = 384, 6, 3, 2, 9, 11
    sd = V.init_aggregator(D, depth, seed=6)
    N = 5 + gh * gw
    tok = torch.randn(S, N, D)
    ref = V.aggregate(sd, tok, gh, gw, H, depth, causal=True)
    x = tok.cuda()
    stream = torch.cuda.current_stream().cuda_stream
    agg = P.Aggregator(sd, D, depth, H, gh, gw, frames_total=S, precision=prec, taps=(0, 1), causal=True)
    agg.forward(x.data_ptr(), stream)
    torch.cuda.synchronize()
    whole = {t: agg.tap_out[t].clone().reshape(S, N, 2 * D) for t in (0, 1)}
    for t in (0, 1):
        got = whole[t].cpu()
        e_frame, e_global = rms_rel(got[..., :D], ref[t][..., :D]), rms_rel(got[..., D:], ref[t][..., D:])
        print(prec, "causal layer", t, "frame", e_frame, "global", e_global)
        assert e_frame < INTER[prec] * (t + 1) and e_global < INTER[prec] * (t + 1)
    assert agg.ops.launches == depth * (8 + 8 + S - 1)                  # a global block: one attention launch per frame
    agg.close()
    full = V.aggregate(sd, tok, gh, gw, H, depth)                        # the unmasked model is a different function
    assert rms_rel(whole[1][0, :, D:].cpu(), full[1][0, :, D:]) > 3 * INTER[prec]
    step = P.Aggregator(sd, D, depth, H, gh, gw, frames_total=1, precision=prec, taps=(0, 1), causal=True, cache_frames=S)
    for rnd in range(2):                                                 # a second stream over the same cache
        for f in range(S):
            step.forward(x[f].data_ptr(), stream, frame_index=f)
            torch.cuda.synchronize()
            for t in (0, 1):
                assert torch.equal(step.tap_out[t].reshape(N, 2 * D), whole[t][f]), (rnd, f, t)
    with pytest.raises(ValueError):
        step.forward(x[0].data_ptr(), stream, frame_index=S)             # beyond the cache
    with pytest.raises(ValueError):
        step.forward(x[0].data_ptr(), stream)                            # a streaming step needs its position
    step.close()
    with pytest.raises(ValueError):
        P.Aggregator(sd, D, depth, H, gh, gw, frames_total=2, precision=prec, causal=True, world=2)


def test_aggregator_graph_replay_equals_eager(lib):
    """The forward captured into a CUDA graph (one host launch per forward) gives the eager launches' result bit for bit."""
    torch.manual_seed(6)
    D, H, S, depth, g = 384, 6, 2, 2, 9
    sd = V.init_aggregator(D, depth, seed=9)
    tok = torch.randn(S, 5 + g * g, D).cuda()
    agg = P.Aggregator(sd, D, depth, H, g, g, frames_total=S, precision="fp16", taps=(1,))
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        agg.forward(tok.data_ptr(), st.cuda_stream)
        st.synchronize()
        eager = agg.tap_out[1].clone()
        agg.capture(tok.data_ptr(), st.cuda_stream)
        agg.tap_out[1].zero_()
        for _ in range(2):
            agg.replay(st.cuda_stream)
        st.synchronize()
        assert torch.equal(agg.tap_out[1], eager)
    with pytest.raises(ValueError):
        agg.capture(tok.data_ptr(), 0)
    agg.close()


@pytest.mark.parametrize("src", [(480, 640), (769, 1025), (500, 500), (33, 57), (1036, 720)])
@pytest.mark.parametrize("dst", [(518, 518), (70, 70), (56, 84)])
def test_vggt_preprocessing_kernel_is_byte_exact(lib, src, dst):
    """uint8 frames -> white square pad -> cubic resize -> / 255: the kernel against the oracle (itself byte-exact with the
    reference module and cv2's own cubic path), two frames per call, every bit."""
    import kutil as K
    from oracle import preprocess_np as Pn
    frames = np.stack([np.random.default_rng(s).integers(0, 256, (*src, 3), dtype=np.uint8) for s in (3, 4)])
    ref = np.concatenate([Pn.preprocess_square_pad_cubic(f, *dst)[0] for f in frames])
    got = K.preprocess_u8_square_pad_cubic(torch.from_numpy(frames).cuda(), *dst)
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy(), ref)


@pytest.mark.parametrize("src", [(480, 640), (1025, 769), (500, 500)])
def test_vggt_postprocessing_matches_the_reference_adapter(lib, src):
    """tools/evaluate_gt.py:240-262 on the device: box crop, bilinear back to the source size, non-depths -> NaN."""
    from oracle import preprocess_np as Pn
    from monocular_depth_estimation_trt_b200 import postprocess as PP
    yy, xx = torch.meshgrid(torch.linspace(0, 1, 518), torch.linspace(0, 1, 518), indexing="ij")
    depth = 1.5 + torch.sin(6 * xx) * torch.cos(5 * yy) * 1.6           # smooth (the two sides compute the sampling coordinate
    #                                                                     in fp32 vs fp64: ~3e-5 pixels apart), partly below the floor
    ref = Pn.vggt_postprocess(depth.numpy(), 518, 518, *src)
    out = torch.zeros(src, device="cuda")
    PP.vggt_postprocess(depth.cuda().data_ptr(), 518, 518, src[0], src[1], out, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    both = np.isfinite(ref) & np.isfinite(got)
    assert np.isnan(ref).any() and (np.isnan(ref) != np.isnan(got)).mean() < 1e-3      # the floor is crossed by the same pixels (fp32 vs fp64 at the edge)
    assert np.abs(got[both] - ref[both]).max() <= 5e-5


# ------------------------------------------------------------------------------------------------ the whole model
@functools.lru_cache(maxsize=1)
def vggt_reference(frames=3, depth=4, seed=0):
    """(state dict, images [frames, 3, 518, 518] in 0..1, oracle depth, trace): ViT-S trunk with four registers, `depth`
    (frame, global) block pairs, DPT head at ViT-S widths; the frames are the reference's synthetic images through its own
    VGGT preprocessing (white square pad + cubic resize + / 255, oracle/preprocess_np.py)."""
    from oracle import preprocess_np as PP
    sd = V.init_vggt("vits", depth=depth, features=64, out_channels=(48, 96, 192, 384), seed=seed)
    imgs = torch.cat([torch.from_numpy(PP.preprocess_square_pad_cubic(
        np.random.default_rng(i).integers(0, 256, (480, 640, 3), dtype=np.uint8), 518, 518))[0] for i in range(frames)])
    taps = tuple(range(depth))[-4:] if depth >= 4 else (0, 0, depth - 1, depth - 1)
    V.calibrate_vggt(sd, imgs, "vits", depth, taps)
    trace = {}
    ref = V.vggt_depth(sd, imgs, "vits", depth, taps, trace)
    return sd, imgs, ref, trace, taps


def test_assemble_tokens_and_layernorm_drop(lib):
    from monocular_depth_estimation_trt_b200.depth_pro import _Ops
    import ctypes as C
    torch.manual_seed(2)
    S, T, D = 3, 11, 384
    patch = torch.randn(S, T, D).half().cuda()
    special = torch.randn(2, 5, D).cuda()
    ops = _Ops("fp16")
    ops.stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for first in (0, 4):
        out = torch.full((S, 5 + T, D), float("nan"), device="cuda")
        ops.assemble_tokens(patch, special, S, T, 5, D, first, out)
        torch.cuda.synchronize()
        want = torch.cat([torch.stack([special[0 if first + s == 0 else 1] for s in range(S)]), patch.float()], dim=1)
        assert torch.equal(out, want)
    # LayerNorm over 2 * 1024 features, dropping the five special tokens of every frame
    x = torch.randn(S * (5 + T), 2048, device="cuda") * 3 + 0.5
    w, b = 1 + 0.1 * torch.randn(2048, device="cuda"), 0.1 * torch.randn(2048, device="cuda")
    y = torch.empty(S * T, 2048, dtype=torch.float16, device="cuda")
    ops.layernorm(x, w, b, y, S * (5 + T), 2048, 1e-5, drop=5, ntok=5 + T)
    torch.cuda.synchronize()
    ref = F.layer_norm(x.reshape(S, 5 + T, 2048)[:, 5:], (2048,), w, b, 1e-5).reshape(S * T, 2048)
    assert float((y.float() - ref).abs().max()) <= 2.0 ** -10 * float(ref.abs().max())


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_whole_model_against_the_oracle(lib, prec):
    """models/vggt/onnx_export.py:38-52 as one engine: images [1, S, 3, 518, 518] -> depth [1, S, 518, 518, 1], through the
    reference-shaped host API, against the oracle's fp32 forward.  One gate (north_star); bf16 reports against it as xfail."""
    from monocular_depth_estimation_trt_b200 import common
    import refsetup as R
    sd, imgs, ref, trace, taps = vggt_reference()
    S = imgs.shape[0]
    with P.VGGTEngine(sd, encoder="vits", depth=4, features=64, out_channels=(48, 96, 192, 384), taps=taps, frames=S, precision=prec) as engine, \
            engine.create_execution_context() as context:
        assert [engine.get_tensor_name(i) for i in range(engine.num_io_tensors)] == ["images", "depth"]
        assert engine.get_tensor_shape("images") == (1, S, 3, 518, 518) and engine.get_tensor_shape("depth") == (1, S, 518, 518, 1)
        inputs, outputs, bindings, stream = common.allocate_buffers(engine)
        inputs[0].host = imgs.numpy()
        out = common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)
        got = out[0].reshape(S, 518, 518).copy()
        again = common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)
        assert np.array_equal(again[0].reshape(S, 518, 518), got)             # graph replay of the aggregator: bit-identical
        # intermediates localise a failure: aggregator input, the four [frame | global] taps, path_1
        tok = context.tokens.cpu()
        assert rms_rel(tok, trace["tokens"]) < INTER[prec]
        for i, layer in enumerate(taps):
            t = engine.agg.tap_out[layer].cpu().reshape(S, 1374, -1)
            assert rms_rel(t, trace["aggregated"][i]) < INTER[prec] * (layer + 2), layer
        p1 = context.get_buffer("path0").float().cpu().reshape(S, 296, 296, 64).permute(0, 3, 1, 2)
        assert rms_rel(p1, trace["path_1"]) < 3 * INTER[prec]
        launches = context.launches_per_enqueue
        common.free_buffers(inputs, outputs, stream)
    worst = {"abs_rel": 0.0, "max_rel": 0.0}
    for s in range(S):
        m = R.compare_depth(ref[s].numpy(), got[s])
        worst = {k: max(worst[k], m[k]) for k in worst}
    print(prec, launches, "launches", worst)
    assert np.isfinite(got).all() and (got > 0).all()
    if prec == "bf16":
        assert worst["abs_rel"] <= 1.2e-2 and worst["max_rel"] <= 1.2e-1, worst             # regression guard, not a parity claim
        if not (worst["abs_rel"] <= 2e-3 and worst["max_rel"] <= 1e-2):
            pytest.xfail(f"bf16 operands miss north_star's gate on the fp32 oracle: {worst}")
    else:
        assert worst["abs_rel"] <= 2e-3 and worst["max_rel"] <= 1e-2, worst


def test_whole_model_at_the_released_width(lib):
    """BASELINE.json configs[4] at the model's own widths -- ViT-L trunk with registers, 24 + 24 aggregator blocks, 256-feature
    DPT head on the (4, 11, 17, 23) taps -- on TWO frames (the oracle's fp32 forward takes ~20 s on the host cores; 16 frames
    would take minutes), seeded init as it is: the log-depth head of the uncalibrated model spreads the depths over 0.07 .. 4."""
    from monocular_depth_estimation_trt_b200 import common
    import refsetup as R
    from oracle import preprocess_np as PP
    taps = (4, 11, 17, 23)
    sd = V.init_vggt("vitl", depth=24, features=256, out_channels=(256, 512, 1024, 1024), seed=0)
    imgs = torch.cat([torch.from_numpy(PP.preprocess_square_pad_cubic(
        np.random.default_rng(i).integers(0, 256, (480, 640, 3), dtype=np.uint8), 518, 518))[0] for i in range(2)])
    ref = V.vggt_depth(sd, imgs, "vitl", 24, taps)
    assert float(ref.min()) > 0.01 and float(ref.std() / ref.mean()) > 0.2          # a non-degenerate map
    with P.VGGTEngine(sd, encoder="vitl", depth=24, features=256, out_channels=(256, 512, 1024, 1024), taps=taps, frames=2,
                      precision="fp16") as engine, engine.create_execution_context() as context:
        inputs, outputs, bindings, stream = common.allocate_buffers(engine)
        inputs[0].host = imgs.numpy()
        out = common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)
        got = out[0].reshape(2, 518, 518).copy()
        common.free_buffers(inputs, outputs, stream)
    worst = {"abs_rel": 0.0, "max_rel": 0.0}
    for s_ in range(2):
        m = R.compare_depth(ref[s_].numpy(), got[s_])
        worst = {k: max(worst[k], m[k]) for k in worst}
    print("vggt ViT-L 24+24 fp16", worst)
    assert worst["abs_rel"] <= 2e-3 and worst["max_rel"] <= 1e-2, worst


def test_streamvggt_whole_model_stream_equals_causal_forward(lib):
    """models/streamvggt as an engine: (a) `causal=True` over the S frames of a scene against the oracle's causal forward
    (north_star's gate); (b) the streaming engine (one frame per execute, cached keys / values): depth map t of the stream is
    depth map t of (a) bit for bit, a reset starts a new stream, and a stream cannot outgrow its cache."""
    from monocular_depth_estimation_trt_b200 import common
    import refsetup as R
    sd, imgs, _, _, taps = vggt_reference()
    S = imgs.shape[0]
    ref = V.vggt_depth(sd, imgs, encoder="vits", depth=4, taps=taps, causal=True)
    kw = dict(encoder="vits", depth=4, features=64, out_channels=(48, 96, 192, 384), taps=taps, precision="fp16")
    with P.VGGTEngine(sd, frames=S, causal=True, **kw) as engine, engine.create_execution_context() as context:
        inputs, outputs, bindings, stream = common.allocate_buffers(engine)
        inputs[0].host = imgs.numpy()
        out = common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)
        whole = out[0].reshape(S, 518, 518).copy()
        common.free_buffers(inputs, outputs, stream)
    worst = {"abs_rel": 0.0, "max_rel": 0.0}
    for s in range(S):
        m = R.compare_depth(ref[s].numpy(), whole[s])
        worst = {k: max(worst[k], m[k]) for k in worst}
    print("streamvggt causal", worst)
    assert worst["abs_rel"] <= 2e-3 and worst["max_rel"] <= 1e-2, worst
    with P.VGGTEngine(sd, frames=1, stream_frames=S, **kw) as engine, engine.create_execution_context() as context:
        assert engine.get_tensor_shape("images") == (1, 1, 3, 518, 518)
        inputs, outputs, bindings, stream = common.allocate_buffers(engine)
        for rnd in range(2):
            for f in range(S):
                inputs[0].host = imgs[f:f + 1].numpy()
                out = common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)
                assert np.array_equal(out[0].reshape(518, 518), whole[f]), (rnd, f)
            with pytest.raises(RuntimeError):
                common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)
            context.reset_stream()
        common.free_buffers(inputs, outputs, stream)
    with pytest.raises(ValueError):
        P.VGGTEngine(sd, frames=2, stream_frames=4, **kw)


def test_get_engine_builds_vggt_and_streamvggt_from_exported_files(lib, tmp_path):
    """models/{vggt,streamvggt}/onnx2trt.py's build call: export file -> get_engine -> allocate_buffers -> do_inference; the
    family in the file picks the attention mask, `stream_frames` the streaming engine; results equal the directly built
    engines' bit for bit and the fingerprint record is written."""
    from monocular_depth_estimation_trt_b200 import common, weights as W
    sd, imgs, _, _, taps = vggt_reference()
    S = imgs.shape[0]
    kw = dict(encoder="vits", depth=4, features=64, out_channels=(48, 96, 192, 384), taps=taps)

    def run(engine, frames):
        with engine, engine.create_execution_context() as context:
            inputs, outputs, bindings, stream = common.allocate_buffers(engine)
            outs = []
            for f in frames:
                inputs[0].host = f.numpy()
                outs.append(common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)[0].copy())
            common.free_buffers(inputs, outputs, stream)
        return outs

    got = {}
    for family in ("vggt", "streamvggt"):
        path = str(tmp_path / f"{family}_518x518.mdew")
        W.save(path, sd, W.describe_vggt(frames=S, family=family, **kw))
        got[family] = run(common.get_engine(path, str(tmp_path / "engine" / f"{family}_fp16.engine"), "fp16"), [imgs])[0]
        direct = run(P.VGGTEngine(sd, frames=S, precision="fp16", causal=family == "streamvggt", **kw), [imgs])[0]
        assert np.array_equal(got[family], direct)
        assert (tmp_path / "engine" / f"{family}_fp16.fingerprint").exists()
    whole = got["streamvggt"].reshape(S, 518, 518)
    assert not np.array_equal(got["vggt"].reshape(S, 518, 518)[0], whole[0])      # frame 0 sees / does not see the other frames
    path = str(tmp_path / "streamvggt_stream.mdew")
    W.save(path, sd, W.describe_vggt(frames=1, family="streamvggt", stream_frames=S, **kw))
    steps = run(common.get_engine(path, "", "fp16"), [imgs[f:f + 1] for f in range(S)])
    for f in range(S):
        assert np.array_equal(steps[f].reshape(518, 518), whole[f]), f
    with pytest.raises(ValueError):
        W.describe_vggt(family="vggt", stream_frames=4, **kw)
    with pytest.raises(ValueError):
        common.get_engine(path, "", "fp16", batch=2)
