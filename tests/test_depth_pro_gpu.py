"""Depth Pro (BASELINE.json configs[3]) on a B200: preprocessing and crop pyramid kernels against torch's interpolate,
the post-processing kernel against the reference script's formulae, and the whole model -- through the reference-shaped
host API (allocate_buffers / do_inference) -- against the oracle (oracle/depth_pro_torch.py, pinned on transformers'
DepthProForDepthEstimation).

Gate for the model outputs: the same single gate as Depth Anything's (tests/test_engine_gpu.py), north_star's numbers (max
relative error <= 1e-2, AbsRel <= 2e-3) for every precision.  fp16 must meet it; bf16 asserts a regression guard (the
precision plan's own error) and reports itself as xfail against the gate with the measured numbers.  Intermediate maps are
gated in RMS-relative error.  The field of view is one fp32 number out of a 16-bit pipeline: 0.05 degrees (fp16) / 0.5 (bf16)."""
import numpy as np
import pytest
import torch

import refsetup as R
from monocular_depth_estimation_trt_b200 import common, depth_pro as DPE, sharding as S

pytestmark = pytest.mark.gpu

GATE = dict(abs_rel=2e-3, max_rel=1e-2)                        # north_star, every precision
BF16_REGRESSION_GUARD = dict(abs_rel=1.2e-2, max_rel=1.2e-1)     # not a parity claim: keeps a broken bf16 kernel from hiding behind the xfail
FOV_DEG = {"fp16": 0.05, "bf16": 0.5}
INTER = {"fp16": 1.2e-3, "bf16": 9e-3}


def rms_rel(got, ref):
    got, ref = got.double(), ref.double()
    return float(((got - ref) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt())


def test_crop_pyramid_matches_torch_interpolate(lib):
    """One launch writes the 35 crops; torch builds the two lower levels with F.interpolate(align_corners=False) and
    slices.  fp32 bilinear with separately rounded steps: within 2 ulp of 1.0 of torch's (FMA-contracted) CPU kernel."""
    torch.manual_seed(4)
    image = torch.randn(3, 1536, 1536)
    ref = S.make_crops(image)
    ops = DPE._Ops("fp16")
    ops.stream = torch.cuda.current_stream().cuda_stream
    import ctypes as C
    ops.stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = torch.full((35, 3, 384, 384), float("nan"), device="cuda")
    plan = [(side, side, y0, x0) for _, side, y0, x0 in S.pyramid_plan(1536)]
    ops.crops(image.cuda().data_ptr(), False, False, 1536, 1536, plan, out)
    torch.cuda.synchronize()
    got = out.cpu()
    assert torch.equal(got[:25], ref[:25])                            # full-resolution crops are copies
    assert float((got - ref).abs().max()) <= 2.4e-7 * float(ref.abs().max())


@pytest.mark.parametrize("hw", [(480, 640), (1536, 1536), (2268, 3024)])
def test_preprocess_u8_matches_the_script_transform(lib, hw):
    """uint8 frame -> ToTensor -> Normalize(0.5, 0.5) -> interpolate(1536): models/depth_pro/onnx2trt.py:56-74."""
    from oracle import depth_pro_torch as DP
    img = R.synthetic_image(3, *hw)
    ref = DP.preprocess(img, 1536)
    out = torch.full((1, 3, 1536, 1536), float("nan"), device="cuda")
    DPE.preprocess_u8(torch.from_numpy(img).cuda(), 1536, out, stream_handle=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert float((out.cpu() - ref).abs().max()) <= 2.4e-7             # values in [-1, 1]: 2 ulp of 1.0


@pytest.mark.parametrize("src", [(480, 640), (1536, 1536)])
def test_postprocess_matches_the_script(lib, src):
    from oracle import depth_pro_torch as DP
    torch.manual_seed(6)
    inv = torch.rand(1, 1, 1536, 1536) * 4 + 1e-5
    inv[0, 0, :4, :4] = 0.0                                            # the clamp's lower end
    fov = torch.tensor([57.3])
    ref, f_px = DP.postprocess(inv, fov, *src)
    depth = torch.full(src, float("nan"), device="cuda")
    fpx = torch.zeros(1, device="cuda")
    DPE.postprocess(inv.cuda().data_ptr(), fov.cuda().data_ptr(), 1536, src[0], src[1], depth, fpx,
                    torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert abs(float(fpx) - float(f_px)) <= 2e-6 * float(f_px)
    rel = ((depth.cpu() - ref[0, 0]).abs() / ref[0, 0]).max()
    assert float(rel) <= 5e-6


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_whole_model_against_the_oracle(lib, prec):
    sd, x, inv, fov, trace = R.depth_pro_reference()
    with DPE.DepthProEngine(sd, encoder="vits", features=64, precision=prec, hook_blocks=(8, 5)) as engine, \
            engine.create_execution_context() as context:
        assert [engine.get_tensor_name(i) for i in range(engine.num_io_tensors)] == ["input", "canonical_inverse_depth", "fov_deg"]
        inputs, outputs, bindings, stream = common.allocate_buffers(engine)
        inputs[0].host = x.numpy()
        outs = common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)
        got_inv = outs[0].reshape(1536, 1536).copy()
        got_fov = float(outs[1][0])
        second = common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)
        assert np.array_equal(second[0].reshape(1536, 1536), got_inv)                     # reproducible bit for bit
        launches = context.launches_per_enqueue
        # intermediates first: they localise a failure
        def nchw(name, side, ch):
            return context.get_buffer(name).float().cpu().reshape(side, side, ch).permute(2, 0, 1)[None]
        assert rms_rel(nchw("proj4", 48, 64), trace["lowres"]) < INTER[prec]
        for i, side in ((3, 96), (2, 192), (1, 384), (0, 768)):
            assert rms_rel(nchw(f"feat{i}", side, 64), trace[f"fusion{i + 1}"]) < INTER[prec], i
        assert rms_rel(nchw("features", 768, 64), trace["fusion0"]) < INTER[prec]
        common.free_buffers(inputs, outputs, stream)
    m = R.compare_depth(inv.numpy(), got_inv)
    print(prec, launches, "launches", m, "fov", got_fov, float(fov))
    assert m["positive"] == 1536 * 1536
    assert abs(got_fov - float(fov)) <= FOV_DEG[prec]
    if prec == "bf16":
        assert m["abs_rel"] <= BF16_REGRESSION_GUARD["abs_rel"] and m["max_rel"] <= BF16_REGRESSION_GUARD["max_rel"], m
        if not (m["abs_rel"] <= GATE["abs_rel"] and m["max_rel"] <= GATE["max_rel"]):
            pytest.xfail(f"bf16 operands miss north_star's gate on the fp32 oracle: abs_rel {m['abs_rel']:.2e}, max_rel {m['max_rel']:.2e}")
    else:
        assert m["abs_rel"] <= GATE["abs_rel"] and m["max_rel"] <= GATE["max_rel"], m


def test_whole_model_at_the_released_width(lib):
    """BASELINE.json configs[3] at the model's own size -- three ViT-L/16 trunks, 256 decoder features, 1536 x 1536 --
    against the oracle's fp32 forward (one forward on the host cores, ~30-60 s).  The weights are the seeded init with the
    last 1x1 convolution rescaled from that one forward's trace (what `calibrate_full` does with two more forwards: only the
    last layer changes, so the calibrated output is recomputed from the traced decoder features).  Hook blocks (11, 4): the
    oracle taps the DINOv2 configuration's blocks (4, 11, 17, 23); the released model hooks (11, 5)."""
    import torch.nn.functional as F
    from oracle import depth_pro_torch as DP
    x = DP.preprocess(R.synthetic_image(0), 1536)
    sd = DP.init_full_state_dict("vitl", features=256, seed=21)
    trace = {}
    _, fov0 = DP.full_forward(sd, x, "vitl", (1, 0), trace)
    with torch.no_grad():
        h = F.conv2d(trace["features"], sd["head.0.weight"], sd["head.0.bias"], padding=1)
        h = F.conv_transpose2d(h, sd["head.1.weight"], sd["head.1.bias"], stride=2)
        h = F.relu(F.conv2d(h, sd["head.2.weight"], sd["head.2.bias"], padding=1))
        z = F.conv2d(h, sd["head.4.weight"], sd["head.4.bias"])
        m0, s0 = float(z.mean()), float(z.std())
        sd["head.4.weight"] = (sd["head.4.weight"] * (0.5 / s0)).contiguous()
        sd["head.4.bias"] = ((sd["head.4.bias"] - m0) * (0.5 / s0) + 3.0).contiguous()
        sd["fov.head.4.bias"] = (sd["fov.head.4.bias"] + (60.0 - float(fov0))).contiguous()
        inv = F.relu(F.conv2d(h, sd["head.4.weight"], sd["head.4.bias"]))[0, 0]
    assert float(inv.min()) > 0.1                                          # ~N(3, 0.5^2), positive everywhere: the relative measures mean something
    with DPE.DepthProEngine(sd, encoder="vitl", features=256, precision="fp16", hook_blocks=(11, 4)) as engine, \
            engine.create_execution_context() as context:
        inputs, outputs, bindings, stream = common.allocate_buffers(engine)
        inputs[0].host = x.numpy()
        outs = common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)
        got_inv, got_fov = outs[0].reshape(1536, 1536).copy(), float(outs[1][0])
        def nchw(name, side, ch):
            return context.get_buffer(name).float().cpu().reshape(side, side, ch).permute(2, 0, 1)[None]
        assert rms_rel(nchw("features", 768, 256), trace["fusion0"]) < INTER["fp16"]
        common.free_buffers(inputs, outputs, stream)
    m = R.compare_depth(inv.numpy(), got_inv)
    print("depth pro ViT-L fp16", m, "fov", got_fov)
    assert m["abs_rel"] <= GATE["abs_rel"] and m["max_rel"] <= GATE["max_rel"], m
    assert abs(got_fov - 60.0) <= FOV_DEG["fp16"]


def test_get_engine_builds_depth_pro_from_an_exported_file(lib, tmp_path):
    """models/depth_pro/onnx2trt.py:99-116 end to end: export file -> get_engine -> allocate_buffers -> do_inference, two
    outputs in spec.json's order; the result equals the directly constructed engine's bit for bit."""
    from monocular_depth_estimation_trt_b200 import weights as W
    sd, x, inv, fov, _ = R.depth_pro_reference()
    path = str(tmp_path / "depth_pro_1536x1536.mdew")
    W.save(path, sd, W.describe_depth_pro("vits", features=64, hook_blocks=(8, 5)))
    results = []
    for make in (lambda: common.get_engine(path, str(tmp_path / "engine" / "depth_pro_fp16.engine"), "fp16"),
                 lambda: DPE.DepthProEngine(sd, encoder="vits", features=64, precision="fp16", hook_blocks=(8, 5))):
        with make() as engine, engine.create_execution_context() as context:
            inputs, outputs, bindings, stream = common.allocate_buffers(engine)
            inputs[0].host = x.numpy()
            outs = common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream)
            results.append((outs[0].copy(), outs[1].copy()))
            common.free_buffers(inputs, outputs, stream)
    assert np.array_equal(results[0][0], results[1][0]) and np.array_equal(results[0][1], results[1][1])
    assert (tmp_path / "engine" / "depth_pro_fp16.fingerprint").exists()
    with pytest.raises(ValueError):
        common.get_engine(path, "", "fp16", batch=2)


@pytest.mark.parametrize("src,focal", [((480, 640), None), ((768, 1024), 886.8), ((1064, 616), None)])
def test_metric3d_postprocess_matches_the_script(lib, src, focal):
    """models/metric3d_v2/onnx2trt.py:148-158 (un-pad, bilinear back to the source size, clamp 0..300) on the device."""
    from oracle import preprocess_np as P
    from monocular_depth_estimation_trt_b200 import postprocess as PP
    torch.manual_seed(2)
    depth = torch.rand(616, 1064) * 400 - 20                          # both ends of the clamp are hit
    ref = P.metric3d_postprocess(depth, src[0], src[1], focal_px=focal)
    out = torch.full(src, float("nan"), device="cuda")
    PP.metric3d_postprocess(depth.cuda().data_ptr(), src[0], src[1], out, focal_px=focal, stream_handle=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert float((out.cpu() - ref).abs().max()) <= 1e-4               # values up to 300: 3 ulp
