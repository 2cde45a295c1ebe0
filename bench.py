#!/usr/bin/env python
"""Headline benchmark: Depth Anything V2 ViT-L 518x518, batch 64 per GPU, images/s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp16] [--impl reference]

A step is one forward of the hot path over one batch of synthetic images.  One JSON line on stdout:

  value         images/s over all ranks, inputs (uint8 source frames) already resident in HBM, timed
                with CUDA events on the launching stream, max over ranks
  e2e           the same metric through the reference-facing call (`do_inference`: pinned host
                buffers, H2D of the uint8 frames and D2H of the float32 depth maps inside the region)
  roofline      the dominant kernel (the launch label with the largest share of the step): algorithmic
                FLOPs (or bytes) of its launches in one step / their summed CUDA-event durations, against
                MEASURED_PEAKS.json; `top_kernels` lists the same for the eight largest kernels
  cpu_baseline  the oracle (CPU fp32 PyTorch port of the reference's forward) on the host cores, bounded
                sample, rank 0 / N=1 only
  --impl reference   times that CPU forward alone (the reference's own CPU path) and prints the same line

Multi-GPU: one process per GPU under torchrun, images sharded by rank (weak scaling, no collective on
the data path); NCCL is used for the barrier and the max-over-ranks of the timings only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "depth_anything_v2_vitl_518_images_per_sec"
UNIT = "images/s"
FLOPS_PER_IMAGE = {"vitl": 1304.2e9, "vitb": 380.7e9, "vits": 115.3e9}   # SURVEY section 8 d closed form
SRC_HW = (480, 640)     # the reference's synthetic-input convention (tests/test_preprocess.py:47-51)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1400.0))),
                    tflops_burst=float(p.get("bf16_tflops", 1590.0)), hbm=float(p.get("hbm_gbs", 6650.0)),
                    source="measured (MEASURED_PEAKS.json, sustained bf16 GEMM)")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def synthetic_batch(batch: int, rank: int):
    """uint8 BGR frames, seed = global image index."""
    h, w = SRC_HW
    return np.stack([np.random.default_rng(rank * batch + i).integers(0, 256, (h, w, 3), dtype=np.uint8)
                     for i in range(batch)])


def max_over_ranks(value: float, world: int, device="cuda") -> float:
    """Max of a per-rank scalar (the timing rule: a multi-GPU number is the slowest rank's)."""
    if world == 1:
        return value
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_rate(world: int, batch: int, steps: int, ms_total: float) -> float:
    """Whole-job images/s: every rank processed `batch` images per step (weak scaling), in the slowest rank's time."""
    return world * batch * steps / (ms_total / 1000.0)


def oracle_setup(encoder: str):
    import torch
    from oracle import dav2_torch as O, preprocess_np as P
    x = torch.from_numpy(P.preprocess_stretch_imagenet(np.random.default_rng(0).integers(0, 256, (*SRC_HW, 3), dtype=np.uint8), 518, 518))
    sd = O.init_state_dict(encoder, seed=0)
    O.calibrate_head(sd, x, encoder)
    return sd, x


def cpu_forward_rate(sd, x, encoder: str, images: int, warm: int = 1):
    """images/s of the oracle's fp32 forward on all host threads, batch 1 per call."""
    import torch
    from oracle import dav2_torch as O
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(warm):
        O.forward(sd, x, encoder, 20.0)
    t0 = time.perf_counter()
    for _ in range(images):
        O.forward(sd, x, encoder, 20.0)
    return images / (time.perf_counter() - t0), torch.get_num_threads()


# ------------------------------------------------------------------------------------------------ arms
def run_reference(args):
    """The reference's own CPU implementation of the path (PyTorch fp32 forward, restated in oracle/):
    every step is a bounded sample of the batch-64 workload (1 image)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sd, x = oracle_setup(args.encoder)
    per_step = 1
    rate_w, threads = cpu_forward_rate(sd, x, args.encoder, max(1, args.warmup) * per_step, warm=0)
    t0 = time.perf_counter()
    rate, threads = cpu_forward_rate(sd, x, args.encoder, args.steps * per_step, warm=0)
    dt = time.perf_counter() - t0
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"depth_anything_v2 {args.encoder} 518x518, CPU fp32 forward, {per_step} image per step "
                               f"(bounded sample of the batch-{args.batch} step)"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} x {per_step} image, oracle/dav2_torch.py fp32, torch {threads} threads"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from monocular_depth_estimation_trt_b200 import build, common, engine as E, weights as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("[MDET] bench.py needs a B200; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    build.build()

    B, enc = args.batch, args.encoder
    sd, x_ref = oracle_setup(enc)        # seeded calibrated weights (the same on every rank)
    meta = W.describe(enc, 518, 518, 20.0)
    eng = E.Engine(E.make_desc(meta, precision=args.precision, batch=B, input_mode="u8_hwc", max_src_hw=SRC_HW, device=local), meta)
    eng.load_state_dict(sd)
    eng.finalize()
    ctx = eng.create_execution_context()
    frames = synthetic_batch(B, rank)
    ctx.set_input_shape("input", (B, SRC_HW[0], SRC_HW[1], 3))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm
    d_in = torch.from_numpy(frames).cuda()
    d_out = torch.empty(B, 518, 518, dtype=torch.float32, device="cuda")
    ctx.set_tensor_address("input", d_in.data_ptr())
    ctx.set_tensor_address("output", d_out.data_ptr())
    side = torch.cuda.Stream()          # a capturable stream: the engine replays its launch sequence as one CUDA graph
    side.wait_stream(torch.cuda.current_stream())
    torch.cuda.set_stream(side)
    stream = side.cuda_stream
    for _ in range(args.warmup):
        ctx.execute_async_v3(stream)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        ctx.execute_async_v3(stream)
    ev1.record()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1), world)
    clocks = sampler.stop() if rank == 0 else None
    value = aggregate_rate(world, B, args.steps, ms_total)

    # ---------------- per-launch timing of one step (events between launches) -> roofline of the dominant kernel
    ops = ctx.execute_timed(stream)
    ops = ctx.execute_timed(stream)
    pk = peaks()
    groups, kernels = {}, {}
    for label, ms, fl, by in ops:
        for table, key in ((groups, label.split(" ")[0]), (kernels, label)):
            g = table.setdefault(key, [0.0, 0.0, 0.0, 0])
            g[0] += ms; g[1] += fl; g[2] += by; g[3] += 1
    step_ms_timed = sum(o[1] for o in ops)
    # dominant kernel = the (kernel, shape) whose launches take the largest share of the step
    dom_key = max(kernels, key=lambda k: kernels[k][0])
    dg = kernels[dom_key]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")     # dram bytes per launch from `ncu --set full` captures
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(dom_key, {}).get("dram_bytes_per_launch")

    def kernel_roofline(key, g):
        tensor = g[1] > 0
        ach = (g[1] / (g[0] / 1000.0) / 1e12) if tensor else (g[2] / (g[0] / 1000.0) / 1e9)
        peak = pk["tflops"] if tensor else pk["hbm"]
        return {"kernel": key, "launches_per_step": g[3], "ms_per_launch": g[0] / g[3], "bound": "tensor" if tensor else "hbm",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s" if tensor else "GB/s", "frac": ach / peak,
                "share_of_step": g[0] / step_ms_timed}

    roofline = kernel_roofline(dom_key, dg)
    roofline.update({"traffic": traffic, "peak_source": pk["source"],
                     "algorithmic_per_launch": {"flops": dg[1] / dg[3], "bytes": dg[2] / dg[3]},
                     "whole_step": {"achieved": value / world * FLOPS_PER_IMAGE[enc] / 1e12,
                                    "frac": value / world * FLOPS_PER_IMAGE[enc] / 1e12 / pk["tflops"],
                                    "frac_of_burst": value / world * FLOPS_PER_IMAGE[enc] / 1e12 / pk["tflops_burst"]},
                     "top_kernels": [kernel_roofline(k, g) for k, g in sorted(kernels.items(), key=lambda kv: -kv[1][0])[:8]]})
    breakdown = {k: {"ms": round(v[0], 3), "tflops": round(v[1] / max(v[0], 1e-9) / 1e9, 1) if v[1] else None,
                     "gbs": round(v[2] / max(v[0], 1e-9) / 1e6, 1), "launches": v[3]} for k, v in sorted(groups.items(), key=lambda kv: -kv[1][0])}

    # ---------------- end-to-end arm: the reference-facing call with host buffers
    inputs, outputs, bindings, cstream = common.allocate_buffers(eng, (B, 518, 518), profile_idx=0)
    inputs[0].host = frames
    from cuda.bindings import runtime as cudart
    for _ in range(max(1, min(args.warmup, 3))):
        res = common.do_inference(ctx, engine=eng, bindings=bindings, inputs=inputs, outputs=outputs, stream=cstream)
    e0 = common.cuda_call(cudart.cudaEventCreate())
    e1 = common.cuda_call(cudart.cudaEventCreate())
    barrier()
    common.cuda_call(cudart.cudaEventRecord(e0, cstream))
    for _ in range(args.steps):
        res = common.do_inference(ctx, engine=eng, bindings=bindings, inputs=inputs, outputs=outputs, stream=cstream)
    common.cuda_call(cudart.cudaEventRecord(e1, cstream))
    common.cuda_call(cudart.cudaEventSynchronize(e1))
    e2e_ms = max_over_ranks(float(common.cuda_call(cudart.cudaEventElapsedTime(e0, e1))), world)
    checksum = float(np.asarray(res[0][:518 * 518], dtype=np.float64).mean())      # the step's result was read on the host
    e2e = {"value": aggregate_rate(world, B, args.steps, e2e_ms), "unit": UNIT,
           "h2d_bytes_per_step": int(inputs[0].nbytes), "d2h_bytes_per_step": int(outputs[0].nbytes),
           "ms_per_step": e2e_ms / args.steps, "mean_depth_image0": checksum}
    h2d, d2h = inputs[0].nbytes, outputs[0].nbytes
    common.free_buffers(inputs, outputs, cstream)

    # ---------------- batch-1 latency, measured the way the reference measures (core/bench.py:182-210:
    # wall clock of one do_inference incl. both copies, warm-up 20, 100 iterations, nearest-rank p50)
    latency = None
    if rank == 0 and world == 1 and not args.no_latency:
        def b1_latency(split_k=False):
            e1 = E.Engine(E.make_desc(meta, precision=args.precision, batch=1, input_mode="u8_hwc", max_src_hw=SRC_HW, device=local,
                                      split_k=split_k), meta)
            e1.load_state_dict(sd)
            e1.finalize()
            c1 = e1.create_execution_context()
            c1.set_input_shape("input", (1, SRC_HW[0], SRC_HW[1], 3))
            i1, o1, b1, s1 = common.allocate_buffers(e1, (1, 518, 518), profile_idx=0)
            i1[0].host = frames[0]
            samples = []
            for it in range(120):
                t0 = time.perf_counter()
                common.do_inference(c1, engine=e1, bindings=b1, inputs=i1, outputs=o1, stream=s1)
                if it >= 20:
                    samples.append((time.perf_counter() - t0) * 1000.0)
            samples.sort()
            rank_p = lambda q: samples[max(0, min(len(samples) - 1, int(np.ceil(q / 100.0 * len(samples))) - 1))]
            common.free_buffers(i1, o1, s1)
            c1.close(); e1.close()
            return {"p50_ms": rank_p(50), "p90_ms": rank_p(90), "p99_ms": rank_p(99), "mean_ms": float(np.mean(samples))}

        latency = b1_latency()          # default configuration: bitwise reproducible
        latency["what"] = ("batch 1, wall clock of do_inference incl. H2D (uint8 frame) and D2H (float32 map), warm-up 20, "
                           "100 iterations")
        # opt-in engine flag: split-K for the residual GEMMs (fp32 adds in arrival order, not bitwise reproducible)
        latency["with_split_k"] = b1_latency(split_k=True)

    # ---------------- CPU baseline beside it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_img = 3
        rate, threads = cpu_forward_rate(sd, x_ref, enc, n_img, warm=1)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n_img} images of the batch, oracle/dav2_torch.py fp32 forward, batch 1 per call, torch {threads} threads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"depth_anything_v2 {enc} 518x518 metric head, batch {B} per GPU, uint8 {SRC_HW[0]}x{SRC_HW[1]} "
                                   f"BGR source frames resident in HBM -> float32 depth [B,518,518]",
                       "weights": "seeded calibrated random init (oracle/dav2_torch.py)", "parallelism": f"images sharded over {world} GPU(s), no collective",
                       "l2": f"no flush: one step streams {eng.workspace_bytes / 2**30:.1f} GiB of activations, far above the 126 MB L2",
                       "b1_note": "batch-1 latency: python bench.py --batch 1"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": ctx.launches_per_enqueue * args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "latency_b1": latency, "breakdown_ms_per_step": breakdown,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--encoder", default="vitl", choices=["vits", "vitb", "vitl"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3      # timing rule: at least three untimed steps
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
