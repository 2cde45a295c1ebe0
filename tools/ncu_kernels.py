#!/usr/bin/env python
"""Launch each of the step's non-GEMM-mainstream kernels ONCE at its batch-64 ViT-L shape, inside a cudaProfilerStart/Stop
range, so that one `ncu --set full --profile-from-start off` run captures them all:

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_kernels \
        python tools/ncu_kernels.py [--precision fp16] [--batch 64]

Order of the profiled launches (the report's launch index): 0 layernorm, 1 preprocess_u8 (kernel 1), 2 proj GEMM (+LayerScale
+ residual reduction), 3 conv256 (refinenet1 RCU conv at 148 x 148), 4 conv128 (output_conv1 at 296 x 296), 5 convT 4x4 +
pixel shuffle, 6 convT 2x2 + pixel shuffle, 7 output_conv2 taps GEMM (N = 384), 8 upconv_head, 9 bilinear 148 -> 296, 10 attention (the kernel the library
picks at this size: three query tiles per persistent CTA, attention_q3.cuh), 11 attention, one query tile per CTA
(attention_tc.cuh), 12 QKV GEMM, 13 FC1 + GELU GEMM, 14 FC2 (+ LayerScale + residual reduction) GEMM.
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kutil as K

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="fp16"); ap.add_argument("--batch", type=int, default=64)
a = ap.parse_args()
dt, B, pr = K.TORCH_DT[a.precision], a.batch, a.precision
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, scale=1.0: (torch.randn(*s, generator=g, device=dev) * scale)
rows, D = B * 1370, 1024
x = rn(rows, D)
lw, lb = 1 + 0.1 * rn(D), 0.1 * rn(D)
src = torch.randint(0, 256, (B, 480, 640, 3), dtype=torch.uint8, device=dev)
att = rn(rows, D).to(dt); wproj = rn(D, D, scale=D ** -0.5).to(dt); bproj, ls = 0.1 * rn(D), 0.5 + 0 * rn(D)
f148 = rn(B, 148, 148, 256).to(dt); w256 = K.pack_conv3x3(rn(256, 256, 3, 3, scale=1 / 48.0), dt); b256 = 0.1 * rn(256)
o148 = torch.empty(B, 148, 148, 256, dtype=dt, device=dev)
f296 = rn(B, 296, 296, 256).to(dt); w128 = K.pack_conv3x3(rn(128, 256, 3, 3, scale=1 / 48.0), dt); b128 = 0.1 * rn(128)
o296 = torch.empty(B, 296, 296, 128, dtype=dt, device=dev)
T = B * 1369
p0 = rn(T, 256).to(dt); wct0 = rn(16 * 256, 256, scale=1 / 16.0).to(dt); bct0 = 0.1 * rn(256); l0 = torch.empty(B, 148, 148, 256, dtype=dt, device=dev)
p1 = rn(T, 512).to(dt); wct1 = rn(4 * 512, 512, scale=1 / 22.0).to(dt); bct1 = 0.1 * rn(512); l1 = torch.empty(B, 74, 74, 512, dtype=dt, device=dev)
wz = rn(384, 128, scale=1 / 11.0).to(dt); z = torch.empty(B, 296, 296, 384, dtype=dt, device=dev)
hb, hw = 0.1 * rn(32), rn(32, scale=0.2)
qkv = rn(rows, 3 * D).to(dt)
ln16 = rn(rows, D).to(dt); wqkv = rn(3 * D, D, scale=D ** -0.5).to(dt); bqkv = 0.1 * rn(3 * D); qkv_out = torch.empty(rows, 3 * D, dtype=dt, device=dev)
wfc1 = rn(4 * D, D, scale=D ** -0.5).to(dt); bfc1 = 0.1 * rn(4 * D); hid = torch.empty(rows, 4 * D, dtype=dt, device=dev)
wfc2 = rn(D, 4 * D, scale=(4 * D) ** -0.5).to(dt); bfc2 = 0.1 * rn(D)


def run():
    K.layernorm(pr, x, lw, lb)
    K.preprocess_u8(pr, src, 518, 518, want_nchw=False)
    K.gemm(pr, att, wproj, K.epilogue(bias=bproj, gamma=ls, x=x, accumulate_x=True, ld_out=D))
    K.conv3x3(pr, f148, w256, 256, K.epilogue(bias=b256, act=2, out=o148, ld_out=256))
    K.conv3x3(pr, f296, w128, 128, K.epilogue(bias=b128, out=o296, ld_out=128))
    K.gemm(pr, p0, wct0, K.epilogue(bias=bct0, out=l0, ld_out=256, shuffle=(4, 256, 37, 37)))
    K.gemm(pr, p1, wct1, K.epilogue(bias=bct1, out=l1, ld_out=512, shuffle=(2, 512, 37, 37)))
    K.gemm(pr, o296.reshape(-1, 128), wz, K.epilogue(out=z, ld_out=384))
    K.upconv_head(pr, z, 518, 518, hb, hw, 0.1, 20.0)
    K.bilinear(pr, f148, 296, 296)
    K.attention(pr, qkv, B, 1370, 16)
    K.attention(pr, qkv, B, 1370, 16, "tc:2")
    K.gemm(pr, ln16, wqkv, K.epilogue(bias=bqkv, out=qkv_out, ld_out=3 * D))
    K.gemm(pr, ln16, wfc1, K.epilogue(bias=bfc1, act=1, out=hid, ld_out=4 * D))
    K.gemm(pr, hid, wfc2, K.epilogue(bias=bfc2, gamma=ls, x=x, accumulate_x=True, ld_out=D))


run(); run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
# plain timing of the same launches (CUDA events) for the record printed next to the capture
names = ["layernorm", "preprocess_u8", "proj gemm", "conv256 148^2", "conv128 296^2", "convT4+shuffle", "convT2+shuffle", "taps gemm", "upconv_head", "bilinear 148->296", "attention (q3)",
         "attention (tc)", "qkv gemm", "fc1+gelu gemm", "fc2+ls+res gemm"]
evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
fns = [lambda: K.layernorm(pr, x, lw, lb), lambda: K.preprocess_u8(pr, src, 518, 518, want_nchw=False),
       lambda: K.gemm(pr, att, wproj, K.epilogue(bias=bproj, gamma=ls, x=x, accumulate_x=True, ld_out=D)),
       lambda: K.conv3x3(pr, f148, w256, 256, K.epilogue(bias=b256, act=2, out=o148, ld_out=256)),
       lambda: K.conv3x3(pr, f296, w128, 128, K.epilogue(bias=b128, out=o296, ld_out=128)),
       lambda: K.gemm(pr, p0, wct0, K.epilogue(bias=bct0, out=l0, ld_out=256, shuffle=(4, 256, 37, 37))),
       lambda: K.gemm(pr, p1, wct1, K.epilogue(bias=bct1, out=l1, ld_out=512, shuffle=(2, 512, 37, 37))),
       lambda: K.gemm(pr, o296.reshape(-1, 128), wz, K.epilogue(out=z, ld_out=384)),
       lambda: K.upconv_head(pr, z, 518, 518, hb, hw, 0.1, 20.0), lambda: K.bilinear(pr, f148, 296, 296),
       lambda: K.attention(pr, qkv, B, 1370, 16), lambda: K.attention(pr, qkv, B, 1370, 16, "tc:2"),
       lambda: K.gemm(pr, ln16, wqkv, K.epilogue(bias=bqkv, out=qkv_out, ld_out=3 * D)),
       lambda: K.gemm(pr, ln16, wfc1, K.epilogue(bias=bfc1, act=1, out=hid, ld_out=4 * D)),
       lambda: K.gemm(pr, hid, wfc2, K.epilogue(bias=bfc2, gamma=ls, x=x, accumulate_x=True, ld_out=D))]
for i, f in enumerate(fns):
    evs[i].record(); f()
evs[-1].record(); torch.cuda.synchronize()
for i, n in enumerate(names):
    print(f"{n:20s} {evs[i].elapsed_time(evs[i + 1]):8.3f} ms")
