"""The reference's measurement and parity harness, restated.  TEST INFRASTRUCTURE ONLY (tests/, bench.py's checker legs,
__graft_entry__.smoke()); the product package never imports this file.

  compare()       core/golden.py:101-174 `compare` for one output: pairwise finite mask, max_abs / mean_abs / rel_mean, and
                  the per-pixel depth metrics (abs_rel, abs_rel_median, rmse, delta1) over pixels where both maps exceed 1e-6
                  -- plus `max_rel`, the per-pixel maximum north_star gates on (the reference does not report it)
  measure()       core/bench.py:182-210 `measure`: warm-up, then perf_counter around fn() + sync() per iteration
  pct()/stats()   core/bench.py:113-150 `Bench.pct` / `Bench.stats`: nearest-rank percentiles (rank = ceil(q/100 * N))
  record()        core/bench.py:345-404 `record` + :316-334 `save` + :247-272 `summarize_outputs` + :275-313 `collect_env`:
                  the result file `<model>_<HxW>_<profile>_<variant>_<precision>.json` in the reference's schema 1

Pinned: tests/test_oracle_harness.py checks every function against the live reference modules where /root/reference is
mounted, and against tests/golden/harness_golden.json (produced by the reference's own functions, oracle/make_golden.py)
everywhere else.  Where the reference is mounted `reference_modules()` hands out the real `core.bench` / `core.golden`,
and callers use those instead of the restatement.
"""
from __future__ import annotations

import importlib
import math
import os
import statistics
import sys
import time
from typing import Callable, Dict, List, Optional

import numpy as np

REFERENCE_ROOT = os.environ.get("MDE_REFERENCE_ROOT", "/root/reference")


def reference_modules():
    """(core.bench, core.golden) of the reference checkout, or (None, None) when it is not mounted (the GPU box)."""
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, "core", "golden.py")):
        return None, None
    added = REFERENCE_ROOT not in sys.path
    if added:
        sys.path.insert(0, REFERENCE_ROOT)
    try:
        return importlib.import_module("core.bench"), importlib.import_module("core.golden")
    except Exception:
        return None, None
    finally:
        if added:
            sys.path.remove(REFERENCE_ROOT)


def compare(ref: np.ndarray, got: np.ndarray, valid_mask: Optional[np.ndarray] = None) -> dict:
    """One entry of core/golden.py `compare` (depth_metrics=True), plus max_rel."""
    a, b = np.asarray(ref, np.float64), np.asarray(got, np.float64)
    if a.shape != b.shape:
        return {"status": "shape_mismatch", "ref_shape": list(a.shape), "new_shape": list(b.shape)}
    finite = np.isfinite(a) & np.isfinite(b)
    if valid_mask is not None and valid_mask.shape == a.shape:
        finite &= valid_mask.astype(bool)
    n = int(finite.sum())
    if n == 0:
        return {"status": "no_finite_overlap"}
    av, bv = a[finite], b[finite]
    d = np.abs(av - bv)
    scale = np.abs(av).mean()
    entry = {"status": "ok", "compared": n, "excluded": int(a.size - n), "max_abs": float(d.max()),
             "mean_abs": float(d.mean()), "rel_mean": float(d.mean() / scale) if scale > 0 else None,
             "identical": bool(d.max() == 0)}
    pos = (av > 1e-6) & (bv > 1e-6)
    k = int(pos.sum())
    if k:
        ap, bp = av[pos], bv[pos]
        rel = np.abs(ap - bp) / ap
        ratio = np.maximum(ap / bp, bp / ap)
        entry.update({"abs_rel": float(rel.mean()), "abs_rel_median": float(np.median(rel)),
                      "rmse": float(np.sqrt(((ap - bp) ** 2).mean())), "delta1": float((ratio < 1.25).mean()),
                      "positive": k, "max_rel": float(rel.max())})
    return entry


def compare_depth(ref: np.ndarray, got: np.ndarray) -> dict:
    """Parity record of one depth map: the reference's `compare` entry (its own function where the checkout is mounted)
    with `max_rel` and the Pearson correlation added."""
    _, golden = reference_modules()
    ref, got = np.asarray(ref), np.asarray(got)
    if ref.shape != got.shape and ref.size == got.size:      # [1, H, W] against [H, W]: the callers' reshape, done here
        ref, got = np.squeeze(ref), np.squeeze(got)
    entry = compare(ref, got)
    if golden is not None and entry.get("status") == "ok":
        theirs = golden.compare({"depth": np.asarray(ref)}, {"depth": np.asarray(got)})["depth"]
        for key, val in theirs.items():
            if isinstance(val, float):
                assert math.isclose(val, entry[key], rel_tol=1e-12, abs_tol=0.0), (key, val, entry[key])
            else:
                assert val == entry[key], (key, val, entry[key])
        entry = {**theirs, "max_rel": entry.get("max_rel")}
    a, b = np.asarray(ref, np.float64).ravel(), np.asarray(got, np.float64).ravel()
    ok = np.isfinite(a) & np.isfinite(b)
    entry["corr"] = float(np.corrcoef(a[ok], b[ok])[0, 1]) if ok.sum() > 1 else float("nan")
    return entry


def measure(fn: Callable[[], object], *, warmup: int = 10, iterations: int = 100, sync: Optional[Callable[[], None]] = None):
    """core/bench.py `measure`: (last return value, per-iteration wall-clock milliseconds incl. sync)."""
    bench, _ = reference_modules()
    if bench is not None:
        return bench.measure(fn, warmup=warmup, iterations=iterations, sync=sync if sync is not None else (lambda: None))
    if sync is None:
        sync = lambda: None
    for _ in range(warmup):
        fn()
    sync()
    out, samples = None, []
    for _ in range(iterations):
        t0 = time.perf_counter()
        out = fn()
        sync()
        samples.append((time.perf_counter() - t0) * 1000.0)
    return out, samples


def pct(samples_ms: List[float], q: float) -> float:
    """Nearest-rank percentile, q in [0, 100] (core/bench.py `Bench.pct`)."""
    if not samples_ms:
        return 0.0
    s = sorted(samples_ms)
    return s[min(len(s) - 1, max(0, math.ceil(q / 100.0 * len(s)) - 1))]


def stats(samples_ms: List[float], warmup: int = 0) -> Dict[str, float]:
    """core/bench.py `Bench.stats` without the stage breakdown."""
    bench, _ = reference_modules()
    if bench is not None:
        return bench.Bench(model="_", samples_ms=list(samples_ms), warmup=warmup).stats()
    if not samples_ms:
        return {}
    mean = statistics.fmean(samples_ms)
    return {"iterations": len(samples_ms), "warmup": warmup, "mean_ms": round(mean, 4), "min_ms": round(min(samples_ms), 4),
            "p50_ms": round(pct(samples_ms, 50), 4), "p90_ms": round(pct(samples_ms, 90), 4),
            "p99_ms": round(pct(samples_ms, 99), 4), "max_ms": round(max(samples_ms), 4),
            "stdev_ms": round(statistics.stdev(samples_ms), 4) if len(samples_ms) > 1 else 0.0,
            "fps": round(1000.0 / mean if mean else 0.0, 2)}


SCHEMA = 1      # core/bench.py:43


def summarize_outputs(outputs) -> Dict[str, dict]:
    """core/bench.py:247-272: shape / dtype / finite counts / min / max / mean over the finite values of every output."""
    out = {}
    for name, arr in outputs.items():
        a = np.asarray(arr)
        finite = np.isfinite(a)
        v = a[finite]
        out[name] = {"shape": list(a.shape), "dtype": str(a.dtype), "finite": int(finite.sum()), "nonfinite": int(a.size - finite.sum()),
                     "min": float(v.min()) if v.size else None, "max": float(v.max()) if v.size else None,
                     "mean": float(v.mean()) if v.size else None}
    return out


def collect_env() -> dict:
    """core/bench.py:275-313: versions, GPU name, driver, graphics clock and its maximum."""
    import platform
    env, device, driver, clock, clock_max = {}, "", "", 0, 0
    try:
        import torch
        env["torch"] = torch.__version__
        if torch.cuda.is_available():
            device = torch.cuda.get_device_name(0)
            env["cuda"] = torch.version.cuda or ""
    except ImportError:
        pass
    try:
        import subprocess
        line = subprocess.check_output(["nvidia-smi", "--query-gpu=driver_version,clocks.gr,clocks.max.gr", "--format=csv,noheader,nounits"],
                                       stderr=subprocess.DEVNULL, text=True, timeout=10).strip().splitlines()[0]
        driver, clock, clock_max = (p.strip() for p in line.split(","))
        clock, clock_max = int(clock), int(clock_max)
    except Exception:
        pass
    env["python"] = platform.python_version()
    return {"versions": env, "device": device, "driver": driver, "clock_mhz": clock, "clock_max_mhz": clock_max}


def record(model: str, samples_ms, *, outputs=None, model_input=None, extra_inputs=None, out_dir: Optional[str] = None,
           echo: bool = True, warmup: int = 0, backend: str = "tensorrt", precision: str = "fp32", profile: str = "bench",
           variant: str = "single", encoder: str = "", input_h: int = 0, input_w: int = 0, engine_path: str = "", notes: str = "",
           stage_samples_ms=None, inputs_dir: Optional[str] = None) -> dict:
    """core/bench.py `record`: write the result file and return the record (the reference returns its Bench object; callers
    in the model scripts ignore the return value).  Where the reference checkout is mounted its own function does the work."""
    import json
    import platform
    bench, _ = reference_modules()
    if out_dir is None:
        raise ValueError("record(): out_dir is required here (the reference defaults to its own reports/bench)")
    if bench is not None:
        b = bench.record(model, samples_ms, outputs=outputs, out_dir=out_dir, echo=echo, warmup=warmup, backend=backend,
                         precision=precision, profile=profile, variant=variant, encoder=encoder, input_h=input_h, input_w=input_w,
                         notes=notes, **({"stage_samples_ms": stage_samples_ms} if stage_samples_ms else {}))
        d = b.to_dict()
    else:
        samples = list(samples_ms)
        d = {"model": model, "samples_ms": [round(x, 4) for x in samples], "warmup": warmup,
             "stage_samples_ms": dict(stage_samples_ms or {}), "backend": backend, "precision": precision, "profile": profile,
             "variant": variant, "encoder": encoder, "input_h": input_h, "input_w": input_w, **collect_env(),
             "host": platform.node(), "outputs": summarize_outputs(outputs) if outputs else {}, "engine_path": "",
             "engine_bytes": 0, "engine_mtime": 0, "onnx_sha256": "", "timestamp": time.strftime("%Y-%m-%dT%H:%M:%S"), "notes": notes,
             "schema": SCHEMA, "stats": stats(samples, warmup)}
        parts = [model, profile, variant, precision]
        if input_h and input_w:
            parts.insert(1, f"{input_h}x{input_w}")
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "_".join(str(p) for p in parts if p) + ".json"), "w", encoding="utf-8") as f:
            json.dump(d, f, indent=2, ensure_ascii=False)
        if echo:
            st = d["stats"]
            print(f"[MDET] {st['iterations']} iterations time: {sum(samples) / 1000.0:.4f} [sec]")
            print(f"[MDET] Average FPS: {st['fps']:.2f} [fps]")
            print(f"[MDET] Average inference time: {st['mean_ms']:.2f} [msec]")
            print(f"[MDET] p50 {st['p50_ms']:.2f} / p90 {st['p90_ms']:.2f} / p99 {st['p99_ms']:.2f} [msec], min {st['min_ms']:.2f}, "
                  f"stdev {st['stdev_ms']:.2f}")
    if model_input is not None and inputs_dir:
        os.makedirs(inputs_dir, exist_ok=True)
        np.save(os.path.join(inputs_dir, f"{model}.npy"), np.ascontiguousarray(model_input))
    return d
