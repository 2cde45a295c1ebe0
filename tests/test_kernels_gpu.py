"""Per-kernel parity on a B200: every mde_k_* entry point against a plain PyTorch fp32 statement of
the same op, on the same (already 16-bit-rounded) inputs.  Tolerances are written at each assert.

fp32-output paths are checked tightly (the only differences are accumulation order); 16-bit
outputs get one rounding of slack: 2^-8 relative for bf16, 2^-11 for fp16.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import kutil as K

pytestmark = pytest.mark.gpu

PRECS = ["bf16", "fp16"]
# unit roundoff of the 16-bit store (8 / 11 significand bits) with 25 % head-room for rounding flips
ULP = {"bf16": 1.25 * 2.0 ** -8, "fp16": 1.25 * 2.0 ** -11}


def rnd(shape, dtype, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(dtype)


# ------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (1370, 3072, 1024), (300, 48, 48), (2740, 1152, 384),
                                   (1369, 32, 128), (4110, 1024, 4096), (77, 768, 640)])
def test_gemm_bias_fp32_out(lib, prec, m, n, k):
    dt = K.TORCH_DT[prec]
    a, b = rnd((m, k), dt, seed=1), rnd((n, k), dt, k ** -0.5, seed=2)
    bias = torch.randn(n, device="cuda")
    x = torch.full((m, n), float("nan"), device="cuda")
    out = torch.zeros(m, n, dtype=dt, device="cuda")
    K.gemm(prec, a, b, K.epilogue(bias=bias, x=x, out=out, ld_out=n))
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t() + bias
    assert K.rel_err(x, ref) < 2e-5            # fp32 accumulate, order only
    assert K.rel_err(out, ref) < ULP[prec]     # one 16-bit rounding


@pytest.mark.parametrize("prec", PRECS)
def test_gemm_gelu(lib, prec):
    dt = K.TORCH_DT[prec]
    m, n, k = 1370, 1536, 384
    a, b = rnd((m, k), dt, seed=3), rnd((n, k), dt, k ** -0.5, seed=4)
    bias = torch.randn(n, device="cuda") * 0.1
    x = torch.empty(m, n, device="cuda")
    K.gemm(prec, a, b, K.epilogue(bias=bias, act=1, x=x, ld_out=n))
    torch.cuda.synchronize()
    ref = F.gelu(a.float() @ b.float().t() + bias)      # exact erf GELU
    assert K.rel_err(x, ref) < 2e-5


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("m,n,k,act,pad", [(1370, 3072, 1024, 0, 0), (2745, 1536, 384, 1, 0), (77, 256, 64, 2, 64),
                                           (4110, 1152, 384, 1, 0), (129, 512, 128, 1, 8), (31, 128, 64, 0, 0)])
def test_gemm_bulk_store_epilogue(lib, prec, m, n, k, act, pad):
    """16-bit output only (QKV / FC1 shape class): registers -> swizzled smem box -> TMA store.  Row tails are
    clipped by the tensor map; the pitch may exceed N; nothing outside [0,m) x [0,n) is written."""
    dt = K.TORCH_DT[prec]
    a, b = rnd((m, k), dt, seed=21), rnd((n, k), dt, k ** -0.5, seed=22)
    bias = torch.randn(n, device="cuda") * 0.5
    ld = n + pad
    out = torch.full((m + 3, ld), 7.0, dtype=dt, device="cuda")
    K.gemm(prec, a, b, K.epilogue(bias=bias, act=act, out=out, ld_out=ld))
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t() + bias
    ref = F.gelu(ref) if act == 1 else (F.relu(ref) if act == 2 else ref)
    assert K.rel_err(out[:m, :n], ref) < ULP[prec]
    assert torch.all(out[m:] == 7.0) and torch.all(out[:, n:] == 7.0)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("m,n,k,pad", [(2740, 384, 1536, 0), (4110, 1024, 1024, 0), (87, 768, 64, 32), (1370, 1024, 4096, 0),
                                       (129, 256, 128, 8), (300, 48, 64, 0)])
def test_gemm_layerscale_residual_in_place(lib, prec, m, n, k, pad):
    """x += gamma * (A B^T + bias) in fp32 (attention projection / FC2 epilogue).  Whole-tile shapes take the TMA
    load / in-place / TMA store path, the last one the staged generic path; rows beyond m and columns beyond n
    are never touched."""
    dt = K.TORCH_DT[prec]
    a, b = rnd((m, k), dt, seed=5), rnd((n, k), dt, k ** -0.5, seed=6)
    bias, gamma = torch.randn(n, device="cuda"), torch.rand(n, device="cuda")
    ld = n + pad
    x0 = torch.randn(m + 2, ld, device="cuda")
    x = x0.clone()
    K.gemm(prec, a, b, K.epilogue(bias=bias, gamma=gamma, x=x, accumulate_x=True, ld_out=ld))
    torch.cuda.synchronize()
    ref = x0[:m, :n] + gamma * (a.float() @ b.float().t() + bias)
    assert K.rel_err(x[:m, :n], ref) < 2e-5
    assert torch.equal(x[m:], x0[m:]) and torch.equal(x[:, n:], x0[:, n:])


@pytest.mark.parametrize("prec", PRECS)
def test_gemm_patch_embed_token_remap(lib, prec):
    """rows (b, t) land at token row b*(T+1)+1+t with pos_embed[1+t] added; cls rows untouched."""
    dt = K.TORCH_DT[prec]
    B, T, D, kp = 3, 1369, 384, 640
    a = rnd((B * T, kp), dt, seed=7)
    a[:, 588:] = 0
    w = rnd((D, kp), dt, 588 ** -0.5, seed=8)
    bias, pos = torch.randn(D, device="cuda"), torch.randn(T + 1, D, device="cuda")
    x = torch.full((B * (T + 1), D), -7.0, device="cuda")
    K.gemm(prec, a, w, K.epilogue(bias=bias, x=x, ld_out=D, tokens=T, pos=pos))
    torch.cuda.synchronize()
    ref = (a.float() @ w.float().t() + bias).reshape(B, T, D) + pos[1:]
    got = x.reshape(B, T + 1, D)
    assert K.rel_err(got[:, 1:], ref) < 2e-5
    assert torch.all(got[:, 0] == -7.0)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("s,c", [(4, 48), (2, 96), (4, 256)])
def test_gemm_conv_transpose_pixel_shuffle(lib, prec, s, c):
    dt = K.TORCH_DT[prec]
    B, H, W = 2, 37, 37
    xin = rnd((B, H, W, c), dt, seed=9)
    wt = rnd((c, c, s, s), dt, c ** -0.5, seed=10)                       # ConvTranspose2d layout [in, out, kh, kw]
    bias = torch.randn(c, device="cuda")
    bmat = wt.permute(2, 3, 1, 0).reshape(s * s * c, c).contiguous()     # row (ky*s+kx)*c + o, col i
    out = torch.zeros(B, H * s, W * s, c, dtype=dt, device="cuda")
    K.gemm(prec, xin.reshape(-1, c), bmat, K.epilogue(bias=bias, out=out, ld_out=c, shuffle=(s, c, H, W)))
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(xin.float().permute(0, 3, 1, 2), wt.float(), bias, stride=s).permute(0, 2, 3, 1)
    assert K.rel_err(out, ref) < ULP[prec]


# ------------------------------------------------------------------------------------------ conv
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("B,H,W,cin,cout", [(2, 19, 19, 64, 64), (1, 37, 37, 48, 64), (2, 74, 74, 256, 256),
                                            (1, 148, 148, 96, 64), (3, 37, 41, 384, 64), (1, 70, 84, 64, 32)])
def test_conv3x3_bias_residuals_relu_copy(lib, prec, B, H, W, cin, cout):
    dt = K.TORCH_DT[prec]
    xin = rnd((B, H, W, cin), dt, seed=11)
    w = rnd((cout, cin, 3, 3), dt, (9 * cin) ** -0.5, seed=12)
    bias = torch.randn(cout, device="cuda")
    r1, r2 = rnd((B, H, W, cout), dt, seed=13), rnd((B, H, W, cout), dt, seed=14)
    out = torch.zeros(B, H, W, cout, dtype=dt, device="cuda")
    out_relu = torch.zeros_like(out)
    K.conv3x3(prec, xin, K.pack_conv3x3(w.float(), dt), cout,
              K.epilogue(bias=bias, res1=r1, res2=r2, out=out, out_relu=out_relu, ld_out=cout))
    torch.cuda.synchronize()
    ref = F.conv2d(xin.float().permute(0, 3, 1, 2), w.float(), bias, padding=1).permute(0, 2, 3, 1) + r1.float() + r2.float()
    assert K.rel_err(out, ref) < ULP[prec]
    assert K.rel_err(out_relu, ref.clamp_min(0)) < ULP[prec]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("max_depth", [20.0, 0.0])
def test_conv3x3_fused_depth_head(lib, prec, max_depth):
    dt = K.TORCH_DT[prec]
    B, H, W, cin = 2, 70, 98, 128
    xin = rnd((B, H, W, cin), dt, seed=15)
    w = rnd((32, cin, 3, 3), dt, (9 * cin) ** -0.5, seed=16)
    bias = torch.randn(32, device="cuda") * 0.1
    hw, hb = torch.randn(32, device="cuda") * 0.5, 0.3
    out = torch.full((B, H, W), float("nan"), device="cuda")
    K.conv3x3(prec, xin, K.pack_conv3x3(w.float(), dt), 32,
              K.epilogue(bias=bias, ld_out=32, head_w=hw, head_b=hb, head_scale=max_depth, head_out=out))
    torch.cuda.synchronize()
    y = F.relu(F.conv2d(xin.float().permute(0, 3, 1, 2), w.float(), bias, padding=1))
    z = (y * hw.view(1, 32, 1, 1)).sum(1) + hb
    ref = torch.sigmoid(z) * max_depth if max_depth > 0 else F.relu(z)
    assert float((out - ref).abs().max()) < 2e-4 * max(1.0, max_depth)   # fp32 path; __expf in the sigmoid


# ------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("variant", ["tc", "tc:0", "tc:4", "q3", "q3:0", "q3:4", "mma"])
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("B,ntok,heads", [(2, 1370, 6), (1, 577, 16), (3, 64, 2), (1, 129, 1), (1, 3349, 2),
                                          (2, 256, 3), (1, 257, 2), (2, 128, 1)])
def test_attention(lib, prec, B, ntok, heads, variant):
    dt = K.TORCH_DT[prec]
    D = heads * 64
    qkv = rnd((B * ntok, 3 * D), dt, seed=17)
    out = K.attention(prec, qkv, B, ntok, heads, variant)
    torch.cuda.synchronize()
    q, k, v = qkv.float().reshape(B, ntok, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v
    ref = ref.transpose(1, 2).reshape(B * ntok, D)
    # P is rounded to 16 bits for the PV MMA and the output once more: two roundings of slack (rms)
    assert K.rms_rel(out, ref) < 2 * ULP[prec]
    assert K.rel_err(out, ref) < 8 * ULP[prec]


# ------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("D", [384, 768, 1024])
def test_layernorm(lib, prec, D):
    rows = 2 * 1370
    x = torch.randn(rows, D, device="cuda") * 3 + 0.5
    w, b = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    out = K.layernorm(prec, x, w, b)
    ref = F.layer_norm(x, (D,), w, b, 1e-6)
    assert K.rel_err(out, ref) < ULP[prec]
    tap = K.layernorm(prec, x, w, b, drop_cls=True, ntok=1370)
    assert tap.shape == (2 * 1369, D)
    assert K.rel_err(tap, ref.reshape(2, 1370, D)[:, 1:].reshape(-1, D)) < ULP[prec]


# ------------------------------------------------------------------------------------------ bilinear / gather
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("hi,wi,ho,wo,c", [(19, 19, 37, 37, 64), (37, 37, 74, 74, 256), (296, 296, 518, 518, 32),
                                           (22, 38, 44, 76, 128)])
def test_bilinear_align_corners(lib, prec, hi, wi, ho, wo, c):
    dt = K.TORCH_DT[prec]
    x = rnd((2, hi, wi, c), dt, seed=18)
    out = K.bilinear(prec, x, ho, wo)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), (ho, wo), mode="bilinear", align_corners=True).permute(0, 2, 3, 1)
    assert K.rel_err(out, ref) < ULP[prec]


@pytest.mark.parametrize("prec", PRECS)
def test_im2col_stride2_matches_conv(lib, prec):
    dt = K.TORCH_DT[prec]
    B, H, W, c = 2, 37, 37, 384
    x = rnd((B, H, W, c), dt, seed=19)
    cols = K.im2col_s2(prec, x)
    w = rnd((64, c, 3, 3), dt, seed=20)
    got = (cols.float() @ w.permute(0, 2, 3, 1).reshape(64, 9 * c).float().t()).reshape(B, 19, 19, 64)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, stride=2, padding=1).permute(0, 2, 3, 1)
    assert K.rel_err(got, ref) < 3e-5     # the gather itself is exact; fp32 matmul order only (K = 9 * c products per element)


# ------------------------------------------------------------------------------------------ kernel (1)
@pytest.mark.parametrize("src_hw", [(480, 640), (720, 1280), (500, 500), (1036, 1036), (300, 777), (518, 518)])
def test_preprocess_u8_bit_exact(lib, src_hw):
    """uint8 resize stage bit-exact and float stage bit-exact (north_star asks <= 1 ulp) against the
    oracle's restatement of core/preprocess.py; im2col layout bit-exact."""
    from oracle import preprocess_np as P
    rng = np.random.default_rng(0)
    imgs = np.stack([rng.integers(0, 256, (*src_hw, 3), dtype=np.uint8) for _ in range(2)])
    cols, nchw = K.preprocess_u8("bf16", torch.from_numpy(imgs).cuda(), 518, 518)
    torch.cuda.synchronize()
    ref = np.concatenate([P.preprocess_stretch_imagenet(im, 518, 518) for im in imgs])
    assert np.array_equal(nchw.cpu().numpy(), ref)
    ref_cols = torch.from_numpy(P.im2col(ref, 14, 640)).to(torch.bfloat16)
    assert torch.equal(cols.cpu(), ref_cols)


@pytest.mark.parametrize("prec", PRECS)
def test_im2col_f32_layout_bit_exact(lib, prec):
    from oracle import preprocess_np as P
    x = torch.randn(2, 3, 518, 518, device="cuda")
    cols = K.im2col_f32(prec, x, 14, 640)
    ref = torch.from_numpy(P.im2col(x.cpu().numpy(), 14, 640)).to(K.TORCH_DT[prec])
    assert torch.equal(cols.cpu(), ref)
    # and it is exactly what a conv with kernel == stride == 14 reads
    w = torch.randn(8, 3, 14, 14)
    a = F.conv2d(x.cpu(), w, stride=14).flatten(2).transpose(1, 2).reshape(-1, 8)
    b = torch.from_numpy(P.im2col(x.cpu().numpy(), 14)) @ w.reshape(8, -1).t()
    assert torch.allclose(a, b, atol=1e-3)


def test_bad_arguments_raise(lib):
    a = torch.zeros(8, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError):
        K.gemm("bf16", a, a, K.epilogue(ld_out=7), n=7)
    with pytest.raises(RuntimeError):
        K.bilinear("bf16", torch.zeros(1, 4, 4, 7, dtype=torch.bfloat16, device="cuda"), 8, 8)


# ------------------------------------------------------------------------------------------ upsample + conv + head tail
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("B,hs,ws,ho,wo,scale", [(2, 296, 296, 518, 518, 20.0), (1, 40, 56, 70, 98, 0.0), (3, 19, 23, 33, 40, 80.0),
                                                 (1, 8, 8, 16, 16, 0.0), (1, 352, 608, 616, 1064, 0.0)])
def test_upconv_head_matches_upsample_conv_head(lib, prec, B, hs, ws, ho, wo, scale):
    """interpolate(align_corners=True) -> conv3x3 128->32 + bias -> ReLU -> conv1x1 32->1 -> activation, computed as
    tap-contracted z at the low resolution + interpolation of z (csrc/upconv_head.cuh); the fp32 torch chain on the
    same 16-bit input is the reference."""
    dt = K.TORCH_DT[prec]
    o1 = rnd((B, hs, ws, 128), dt, seed=31)
    w2 = rnd((32, 128, 3, 3), torch.float32, (9 * 128) ** -0.5, seed=32)
    b2 = torch.randn(32, device="cuda") * 0.2
    w3 = torch.randn(32, device="cuda") * 0.3
    b3 = 0.1
    # z[.., t*32+o] = sum_c w2[o, c, ky, kx] * o1[.., c], rounded once to 16 bits (what the GEMM epilogue stores)
    wz = w2.permute(2, 3, 0, 1).reshape(288, 128)
    z = torch.zeros(B, hs, ws, 384, dtype=dt, device="cuda")
    z[..., :288] = (o1.float() @ wz.t()).to(dt)
    z[..., 288:] = 1e4          # the pad channels must never be read
    got = K.upconv_head(prec, z, ho, wo, b2, w3, b3, scale)
    torch.cuda.synchronize()
    up = F.interpolate(o1.float().permute(0, 3, 1, 2), size=(ho, wo), mode="bilinear", align_corners=True)
    h = F.relu(F.conv2d(up, w2, b2, padding=1))
    v = (h * w3.view(1, 32, 1, 1)).sum(1) + b3
    ref = scale * torch.sigmoid(v) if scale > 0 else F.relu(v)
    assert torch.isfinite(got).all()
    assert float((got - ref).abs().max()) < (4.0 if prec == "bf16" else 1.0) * 2.0 ** -8 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("B,h,w,ho,wo", [(2, 518, 518, 480, 640), (1, 518, 518, 2268, 3024), (3, 37, 50, 37, 50), (1, 266, 518, 1, 7)])
def test_resize_depth_matches_the_scripts_postprocessing(lib, B, h, w, ho, wo):
    """models/depth_anything_v2/onnx2trt.py:111-117: interpolate(bilinear, align_corners=True) then clamp(1e-3, 1e3)."""
    g = torch.Generator(device="cuda").manual_seed(41)
    d = torch.rand(B, h, w, generator=g, device="cuda") * 30.0 - 1.0          # some values below the lower clamp
    d[0, 0, 0] = 5000.0                                                        # and one above the upper
    got = K.resize_depth(d, ho, wo)
    torch.cuda.synchronize()
    ref = torch.clamp(F.interpolate(d[:, None], (ho, wo), mode="bilinear", align_corners=True)[:, 0], 1e-3, 1e3)
    assert float((got - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    assert float(got.min()) >= 1e-3 and float(got.max()) <= 1e3


# ------------------------------------------------------------------------------------------ sharded global attention pieces (one GPU)
@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("B,nq,nkv,heads", [(1, 700, 2748, 6), (2, 130, 517, 2), (1, 1374, 1374, 16),
                                            (1, 2748, 21984, 16), (20, 1374, 2748, 16), (24, 1300, 1501, 16)])
def test_attention_queries_and_keys_from_different_row_sets(lib, prec, B, nq, nkv, heads):
    """Queries = a rank's tokens, keys/values = everybody's tokens in a gathered [nkv, 2D] k|v buffer.  Then 2 of VGGT's
    16 frames per rank against all 16 (the 8-GPU shape), and two shapes large enough (>= 8 rounds of work items) for the library
    to pick the persistent three-query-tile kernel with keys from another row set, the second one ragged on both sides."""
    dt = K.TORCH_DT[prec]
    D = heads * 64
    qkv = rnd((B * nq, 3 * D), dt, seed=51)
    kv = rnd((B * nkv, 2 * D), dt, seed=52)
    out = K.attention_kv(prec, qkv, 3 * D, kv, 2 * D, 0, D, B, nq, nkv, heads)
    torch.cuda.synchronize()
    q = qkv[:, :D].float().reshape(B, nq, heads, 64).transpose(1, 2)
    k = kv[:, :D].float().reshape(B, nkv, heads, 64).transpose(1, 2)
    v = kv[:, D:].float().reshape(B, nkv, heads, 64).transpose(1, 2)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * nq, D)
    assert K.rel_err(out, ref) < 4 * ULP[prec]


@pytest.mark.parametrize("prec", PRECS)
def test_gemm_fused_gather_single_rank(lib, prec):
    """QKV projection whose K|V column boxes go to `gather` destinations (here: two local buffers standing in for two
    ranks) instead of the local output; rows beyond m are clipped in every destination."""
    dt = K.TORCH_DT[prec]
    m, D = 1374 + 13, 384
    a, w = rnd((m, D), dt, seed=53), rnd((3 * D, D), dt, D ** -0.5, seed=54)
    bias = torch.randn(3 * D, device="cuda") * 0.3
    out = torch.full((m, 3 * D), 7.0, dtype=dt, device="cuda")
    g = [torch.full((2 * m + 5, 2 * D), 7.0, dtype=dt, device="cuda") for _ in range(2)]
    row0 = m                                                   # we are "rank 1" of 2: our rows start at row m
    ptrs = [t.data_ptr() + row0 * 2 * D * 2 for t in g]
    K.gemm(prec, a, w, K.epilogue(bias=bias, out=out, ld_out=3 * D, gather=(D, 2 * D, ptrs)))
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    assert K.rel_err(out[:, :D], ref[:, :D]) < ULP[prec]
    assert torch.all(out[:, D:] == 7.0)                        # K|V did not go to the local output
    for t in g:
        assert K.rel_err(t[row0:row0 + m], ref[:, D:]) < ULP[prec]
        assert torch.all(t[:row0] == 7.0) and torch.all(t[row0 + m:] == 7.0)


# ------------------------------------------------------------------------------------------ Metric3D V2 input contract
@pytest.mark.parametrize("src", [(480, 640), (720, 1280), (300, 777), (1036, 1036), (1232, 2128), (37, 41)])
@pytest.mark.parametrize("dst", [(616, 1064), (518, 518)])
def test_preprocess_keep_ratio_pad_bit_exact(lib, src, dst):
    """uint8 BGR -> keep-ratio INTER_LINEAR resize (truncated inner size) -> centre pad with the mean colour -> float32
    0..255, no normalisation: byte-exact with the oracle's restatement of core/preprocess.py MODELS['metric3d_v2']
    (which tests/test_oracle_preprocess.py pins on the reference module's own outputs); im2col rows = that tensor's patches."""
    from oracle import preprocess_np as P
    img = np.random.default_rng(61).integers(0, 256, (*src, 3), dtype=np.uint8)
    ref = P.preprocess_pad_none(img, *dst)
    cols, nchw = K.preprocess_u8_pad("fp16", torch.from_numpy(img)[None].cuda(), dst[0], dst[1])
    torch.cuda.synchronize()
    assert np.array_equal(nchw.cpu().numpy(), ref)
    ref_cols = torch.from_numpy(P.im2col(ref, 14, 640)).half()      # integers 0..255 are exact in fp16
    assert torch.equal(cols.cpu(), ref_cols)


# ------------------------------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("variant", ["tc", "tc:0", "tc:3", "q3", "q3:3"])
@pytest.mark.parametrize("B,ntok,heads", [(1, 1, 1), (2, 5, 2), (1, 31, 1), (1, 32, 1), (1, 33, 3), (1, 127, 1), (1, 128, 2), (3, 255, 1), (1, 256, 1), (1, 257, 1)])
def test_attention_ragged_token_counts(lib, variant, B, ntok, heads):
    """One token, tile boundaries and every +-1 around them: masking of the ragged last key tile, clipped query rows,
    images that end inside another image's tile."""
    dt = torch.float16
    D = heads * 64
    qkv = rnd((B * ntok, 3 * D), dt, seed=71)
    out = K.attention("fp16", qkv, B, ntok, heads, variant)
    torch.cuda.synchronize()
    q, k, v = (qkv[:, i * D:(i + 1) * D].float().reshape(B, ntok, heads, 64).transpose(1, 2) for i in range(3))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * ntok, D)
    assert torch.isfinite(out).all()
    assert K.rel_err(out, ref) < 4 * ULP["fp16"]


@pytest.mark.parametrize("m,n,k", [(1, 256, 64), (1, 1024, 1024), (127, 8, 8), (129, 264, 72), (256, 256, 4096), (5, 3072, 384)])
def test_gemm_degenerate_shapes(lib, m, n, k):
    """A single row, N / K that are not tile multiples, one tile with a long K loop."""
    dt = torch.float16
    a, b = rnd((m, k), dt, seed=72), rnd((n, k), dt, k ** -0.5, seed=73)
    bias = torch.randn(n, device="cuda")
    out = torch.full((m + 1, n), 7.0, dtype=dt, device="cuda")
    x0 = torch.randn(m + 1, n, device="cuda")
    x = x0.clone()
    K.gemm("fp16", a, b, K.epilogue(bias=bias, out=out, ld_out=n))
    K.gemm("fp16", a, b, K.epilogue(bias=bias, x=x, accumulate_x=True, ld_out=n))
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t() + bias
    assert K.rel_err(out[:m], ref) < ULP["fp16"] and torch.all(out[m:] == 7.0)
    assert K.rel_err(x[:m], x0[:m] + ref) < 2e-5 and torch.equal(x[m:], x0[m:])


def test_empty_problems_are_refused(lib):
    """No rows / no tokens: an error through the ABI, never a launch."""
    import ctypes as C
    from monocular_depth_estimation_trt_b200 import _lib
    t = torch.zeros(8, 64, dtype=torch.float16, device="cuda")
    ep = K.epilogue(out=t, ld_out=64)
    assert lib.mde_k_gemm(0, K.ptr(t), 0, 64, 64, K.ptr(t), 8, 64, C.byref(ep), None) != 0
    assert "empty" in _lib.last_error()
    assert lib.mde_k_attention(0, K.ptr(t), K.ptr(t), 1, 0, 1, None) != 0
    assert lib.mde_k_layernorm(0, K.ptr(t), K.ptr(t), K.ptr(t), K.ptr(t), 0, 384, 1e-6, 0, 0, None) != 0


def test_peer_signal_and_wait_single_rank(lib):
    """The stream-ordered hand-shake kernels with one rank (the multi-rank run is tests/mgpu/sharded_global_attention.py):
    signal publishes the epoch in our own flag array, wait returns once every slot has reached it; epochs only grow."""
    from monocular_depth_estimation_trt_b200 import sharding as S
    sync = S.PeerSync(1, 0)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        sync.wait_acks(s)
        sync.signal_ready(s)
        sync.wait_ready(s)
        sync.signal_acks(s)
    torch.cuda.synchronize()
    flags = sync.ready.view().view(torch.int32)
    assert int(flags[0]) == 3 and int(sync.acks.view().view(torch.int32)[0]) == 3
    sync.close()


# ------------------------------------------------------------------------------------------ guard bands
# compute-sanitizer is closed on this GPU pool (profiles/r02_sanitizer_refused.log), so out-of-bounds writes are hunted the
# other way: every output lives inside a larger allocation whose margins are filled with a sentinel, ragged shapes on purpose.
def guarded(rows, cols, dtype, margin_rows=64, fill=7.0):
    buf = torch.full((rows + 2 * margin_rows, cols), fill, dtype=dtype, device="cuda")
    view = buf[margin_rows:margin_rows + rows]

    def intact():
        return bool((buf[:margin_rows] == fill).all()) and bool((buf[margin_rows + rows:] == fill).all())
    return buf, view, intact


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("m,n,k", [(300, 256, 128), (129, 512, 64), (1, 128, 192), (257, 1024, 1024)])
def test_gemm_epilogues_stay_inside_their_outputs(lib, prec, m, n, k):
    dt = K.TORCH_DT[prec]
    a, b = rnd((m, k), dt, seed=1), rnd((n, k), dt, 0.05, seed=2)
    bias = rnd((n,), torch.float32, seed=3)
    ref = a.float() @ b.float().t() + bias
    # bulk-store epilogue (TMA store clips the ragged last row block)
    _, out, ok = guarded(m, n, dt)
    K.gemm(prec, a, b, K.epilogue(bias=bias, out=out, ld_out=n))
    torch.cuda.synchronize()
    assert ok() and K.rel_err(out, ref) < ULP[prec]
    # residual-stream reduction (cp.reduce.async.bulk.tensor into fp32 rows)
    _, x, okx = guarded(m, n, torch.float32)
    x.zero_()
    K.gemm(prec, a, b, K.epilogue(bias=bias, x=x, accumulate_x=True, ld_out=n))
    torch.cuda.synchronize()
    assert okx() and K.rel_err(x, ref) < 2e-5
    # staged generic epilogue (16-bit residual + ReLU copy), narrow ragged N
    n2 = 40
    _, o2, ok2 = guarded(m, n2, dt)
    _, o3, ok3 = guarded(m, n2, dt)
    res = rnd((m, n2), dt, seed=4)
    K.gemm(prec, a, b[:n2].contiguous(), K.epilogue(bias=bias[:n2].contiguous(), res1=res, out=o2, out_relu=o3, ld_out=n2))
    torch.cuda.synchronize()
    ref2 = ref[:, :n2] + res.float()
    assert ok2() and ok3() and K.rel_err(o2, ref2) < ULP[prec] and K.rel_err(o3, ref2.clamp_min(0)) < ULP[prec]


@pytest.mark.parametrize("B,H,W_,cin,cout", [(1, 19, 19, 64, 64), (2, 37, 50, 48, 256), (1, 5, 131, 128, 32)])
def test_conv3x3_stays_inside_its_output(lib, B, H, W_, cin, cout):
    dt = torch.float16
    x = rnd((B, H, W_, cin), dt, seed=5)
    w = rnd((cout, cin, 3, 3), torch.float32, (9 * cin) ** -0.5, seed=6)
    _, out, ok = guarded(B * H * W_, cout, dt)
    K.conv3x3("fp16", x, K.pack_conv3x3(w, dt), cout, K.epilogue(out=out, ld_out=cout))
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.half().float(), None, padding=1).permute(0, 2, 3, 1).reshape(-1, cout)
    assert ok() and K.rel_err(out, ref) < ULP["fp16"]


@pytest.mark.parametrize("B,ntok,heads", [(1, 129, 2), (2, 1370, 1), (3, 5, 1)])
def test_attention_and_layernorm_stay_inside_their_outputs(lib, B, ntok, heads):
    import ctypes as C
    from monocular_depth_estimation_trt_b200 import _lib
    dt, D = torch.float16, heads * 64
    qkv = rnd((B * ntok, 3 * D), dt, seed=8)
    _, out, ok = guarded(B * ntok, D, dt)
    _lib.check(lib.mde_k_attention(_lib.PRECISIONS["fp16"], K.ptr(qkv), K.ptr(out), B, ntok, heads, K.stream()), "attention")
    torch.cuda.synchronize()
    q, k, v = (qkv[:, i * D:(i + 1) * D].float().reshape(B, ntok, heads, 64).transpose(1, 2) for i in range(3))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * ntok, D)
    assert ok() and K.rel_err(out, ref) < 4 * ULP["fp16"]
    # LayerNorm dropping the first token of every image: dense [B][ntok - 1] output
    if ntok > 1:
        x = rnd((B * ntok, 384), torch.float32, 2.0, seed=9)
        w, b = 1 + 0.1 * rnd((384,), torch.float32, seed=10), 0.1 * rnd((384,), torch.float32, seed=11)
        _, y, oky = guarded(B * (ntok - 1), 384, dt)
        _lib.check(lib.mde_k_layernorm(_lib.PRECISIONS["fp16"], K.ptr(x), K.ptr(w), K.ptr(b), K.ptr(y), B * ntok, 384, 1e-6, 1, ntok, K.stream()), "layernorm")
        torch.cuda.synchronize()
        refy = F.layer_norm(x.reshape(B, ntok, 384)[:, 1:], (384,), w, b, 1e-6).reshape(-1, 384)
        assert oky() and K.rel_err(y, refy) < ULP["fp16"]
