#!/usr/bin/env python
"""Energy per launch of the step's kernels under SUSTAINED operation (each looped for --seconds, NVML total-energy counter):
J per launch, average power, SM clock under load, ms per launch at that clock.  The batch-64 step runs at the board's power
cap, where time = energy / cap: this is the budget that matters there, not stand-alone cycles.
    python tools/energy_probe.py [--seconds 1.5] [--precision fp16] [--out gpurun_out/energy.txt]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import pynvml
import kutil as K

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=1.5); ap.add_argument("--precision", default="fp16")
ap.add_argument("--out", default=""); ap.add_argument("--cases", default="")
a = ap.parse_args()
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
dt = K.TORCH_DT[a.precision]
dev = "cuda"
B, N, H, D = 64, 1370, 16, 1024
M = B * N
lines = []


def measure(name, fn, flops=0.0, n_in_step=0):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    # calibrate the launch count for ~a.seconds
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / 5
    n = max(10, int(a.seconds * 1000 / per))
    # warm the power state for a third of the window, then measure
    for _ in range(n // 3): fn()
    torch.cuda.synchronize()
    j0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    t0 = time.perf_counter()
    e0.record()
    clocks = []
    for i in range(n):
        fn()
        if i % max(1, n // 8) == 0:
            clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
    e1.record(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    j1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    ms = e0.elapsed_time(e1) / n
    joule = (j1 - j0) / 1000.0 / n
    watts = (j1 - j0) / 1000.0 / (t1 - t0)
    clk = sorted(clocks)[len(clocks) // 2]
    tf = flops / ms / 1e9 if flops else 0.0
    step = f"  x{n_in_step} = {joule * n_in_step:6.2f} J, {ms * n_in_step:6.2f} ms per step" if n_in_step else ""
    line = f"{name:34s} {ms:8.4f} ms  {joule:8.4f} J  {watts:6.0f} W  {clk:5d} MHz  {tf:7.1f} TFLOP/s  {joule / flops * 1e12 if flops else 0:6.3f} pJ/FLOP{step}"
    print(line, flush=True)
    lines.append(line)
    time.sleep(0.5)


def mk(r, c): return (torch.randn(r, c, device=dev) * 0.5).to(dt)


cases = a.cases.split(",") if a.cases else ["idle", "matmul", "qkv", "proj", "fc1", "fc2", "attn", "attn_q3", "attn_p0", "ln", "conv256", "conv128", "step", "step_bf16"]
for case in cases:
    if case == "idle":
        torch.cuda.synchronize()
        j0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h); t0 = time.perf_counter(); time.sleep(1.0)
        j1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h); t1 = time.perf_counter()
        print(f"idle: {(j1 - j0) / 1000 / (t1 - t0):.0f} W", flush=True)
        lines.append(f"idle: {(j1 - j0) / 1000 / (t1 - t0):.0f} W")
    elif case == "matmul":
        x, y = mk(8192, 8192), mk(8192, 8192)
        measure("torch.matmul 8192^3 (cuBLAS)", lambda: torch.matmul(x, y), 2.0 * 8192 ** 3)
    elif case in ("qkv", "proj", "fc1", "fc2"):
        if case == "qkv":   n, k, kw = 3 * D, D, dict(bias=True, out=True)
        elif case == "proj": n, k, kw = D, D, dict(bias=True, gamma=True, x=True)
        elif case == "fc1":  n, k, kw = 4 * D, D, dict(bias=True, act=1, out=True)
        else:  n, k, kw = D, 4 * D, dict(bias=True, gamma=True, x=True)
        A, Bm = mk(M, k), mk(n, k)
        bias = torch.randn(n, device=dev)
        gamma = torch.rand(n, device=dev) * 0.01 if kw.get("gamma") else None
        x = torch.randn(M, n, device=dev) if kw.get("x") else None
        out = torch.empty(M, n, dtype=dt, device=dev) if kw.get("out") else None
        ep = K.epilogue(bias=bias, gamma=gamma, act=kw.get("act", 0), x=x, accumulate_x=bool(kw.get("x")), out=out, ld_out=n)
        measure(f"gemm {case} {M}x{n}x{k}", lambda: K.gemm(a.precision, A, Bm, ep), 2.0 * M * n * k, 24)
        del A, Bm, x, out
    elif case.startswith("attn"):
        qkv = torch.randn(M, 3 * D, device=dev).to(dt)
        v = {"attn": "tc:2", "attn_q3": "q3:2", "attn_p0": "tc:0", "attn_p4": "tc:4"}.get(case) or case.split("=")[1]   # attn=q3:1
        measure(f"attention[{v}]", lambda: K.attention(a.precision, qkv, B, N, H, v), 4.0 * B * H * N * N * 64, 24)
        del qkv
    elif case == "ln":
        x = torch.randn(M, D, device=dev); w = torch.randn(D, device=dev); b = torch.randn(D, device=dev)
        measure("layernorm fp32 -> 16 bit", lambda: K.layernorm(a.precision, x, w, b), 0.0, 48)
        del x
    elif case in ("conv256", "conv128"):
        cout = 256 if case == "conv256" else 128
        hw = 148 if case == "conv256" else 296
        x = (torch.randn(B, hw, hw, 256, device=dev) * 0.5).to(dt)
        w = K.pack_conv3x3(torch.randn(cout, 256, 3, 3, device=dev) * 0.02, dt)
        out = torch.empty(B * hw * hw, cout, dtype=dt, device=dev)
        ep = K.epilogue(bias=torch.randn(cout, device=dev), out=out, ld_out=cout)
        measure(f"conv3x3 {B}x{hw}x{hw} 256->{cout}", lambda: K.conv3x3(a.precision, x, w, cout, ep), 2.0 * B * hw * hw * 9 * 256 * cout, 5 if case == "conv256" else 1)
        del x, out
    elif case.startswith("step"):
        from monocular_depth_estimation_trt_b200 import engine as E, weights as W
        from oracle import dav2_torch as O
        prec = "bf16" if "bf16" in case else a.precision
        Bs = int(case.split(":")[1]) if ":" in case else B
        meta = W.describe("vitl", 518, 518, 20.0)
        eng = E.Engine(E.make_desc(meta, precision=prec, batch=Bs, input_mode="u8_hwc", max_src_hw=(480, 640)), meta)
        eng.load_state_dict(O.init_state_dict("vitl", 0)); eng.finalize()
        ctx = eng.create_execution_context()
        ctx.set_input_shape("input", (Bs, 480, 640, 3))
        src = torch.randint(0, 256, (Bs, 480, 640, 3), dtype=torch.uint8, device="cuda")
        out = torch.empty(Bs, 518, 518, device="cuda")
        ctx.set_tensor_address("input", src.data_ptr()); ctx.set_tensor_address("output", out.data_ptr())
        s_ = torch.cuda.current_stream().cuda_stream
        measure(f"whole step B={Bs} ViT-L {prec}", lambda: ctx.execute_async_v3(s_), 1304.2e9 * Bs, 1)
        ctx.close() if hasattr(ctx, "close") else None
        eng.close()
        del src, out
    torch.cuda.empty_cache()
if a.out:
    open(a.out, "w").write("\n".join(lines) + "\n")
