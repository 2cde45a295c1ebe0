"""The drop-in, exercised with the reference's own harness.

integration/models/depth_anything_v2/onnx2trt.py is the reference's stage script after the switch INTEGRATION.md describes.
Here it is held to the reference's contract three ways:

  * statically: the reference's AST contract tests (tests/test_bench_wiring.py, imported from the checkout where it is mounted
    and pointed at integration/models; an equivalent local restatement of the same rules runs everywhere);
  * wired against the REAL harness: an overlay tree -- the reference's core/bench.py, core/preprocess.py, core/golden.py, ...
    next to the two replacement files core/common.py / core/common_runtime.py -- in which the script imports and runs up to
    the engine build (which fails loudly without a GPU: no fallback);
  * on the GPU (-m gpu): the script runs end to end (export -> get_engine -> allocate_buffers -> bench.measure(do_inference) ->
    post-processing -> bench.record), its depth map is compared with the oracle by core.golden.compare's rules, and the
    record it wrote is schema-checked.  Records made that way on a B200 are committed under profiles/r02_reports_bench/ and
    loaded here with the reference's core.bench.load_all and rendered with its tools/compare.py.
"""
import ast
import importlib.util
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MDE_REFERENCE_ROOT", "/root/reference")
SCRIPT = os.path.join(ROOT, "integration", "models", "depth_anything_v2", "onnx2trt.py")
HAVE_REF = os.path.isfile(os.path.join(REF, "core", "bench.py"))
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="reference checkout not mounted")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------------ static contract
def test_script_contract_static():
    """The rules of the reference's tests/test_bench_wiring.py:87-222, restated: one timing loop (bench.measure around
    do_inference only), one bench.record with a literal model key and the five required keywords, the shared warm-up /
    iteration counts, no hand-rolled timing, no upstream package imports, preprocessing through preprocess_for."""
    src = open(SCRIPT, encoding="utf-8").read()
    tree = ast.parse(src)
    assert "from core import bench" in src and "bench.measure(" in src and "bench.record(" in src
    assert not re.search(r"\bdur_time\b|\bavg_time\b", src) and "time.time()" not in src
    assert re.search(r"^\s*warmup\s*=\s*20\b", src, re.M) and re.search(r"^\s*iteration\s*=\s*100\b", src, re.M)
    records = [n for n in ast.walk(tree) if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute) and n.func.attr == "record"]
    assert len(records) == 1 and records[0].args[0].value == "depth_anything_v2"
    assert {"warmup", "precision", "profile", "input_h", "input_w"} <= {k.arg for k in records[0].keywords}
    measure = next(n for n in ast.walk(tree) if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute) and n.func.attr == "measure")
    body = ast.unparse(measure.args[0])
    assert body.count("do_inference") == 1 and "interpolate" not in body
    imports = {(n.module if isinstance(n, ast.ImportFrom) else a.name) for n in ast.walk(tree)
               if isinstance(n, (ast.Import, ast.ImportFrom)) for a in (n.names if isinstance(n, ast.Import) else [None])}
    assert not {m for m in imports if m and m.split(".")[0] in ("tensorrt", "depth_anything_v2", "onnx", "onnxruntime")}
    assert re.search(r"pp\.preprocess_for\(raw_img,\s*'depth_anything_v2',\s*\(input_h,\s*input_w\)\)", src)
    # the two replacement files are exactly the import lines INTEGRATION.md lists
    for name in ("common.py", "common_runtime.py"):
        body = [ln for ln in open(os.path.join(ROOT, "integration", "core", name)).read().splitlines() if ln and not ln.startswith("#")]
        assert all(ln.startswith("from monocular_depth_estimation_trt_b200.") for ln in body) and body


@needs_ref
def test_script_passes_the_references_own_ast_tests(monkeypatch, capsys):
    sys.path.insert(0, os.path.join(REF, "tools"))            # test_record_names_the_model imports compare.FOLDER
    try:
        wiring = _load(os.path.join(REF, "tests", "test_bench_wiring.py"), "ref_test_bench_wiring")
    finally:
        sys.path.remove(os.path.join(REF, "tools"))
    monkeypatch.setattr(wiring, "MODELS", os.path.join(ROOT, "integration", "models"))
    ran = 0
    for name in ("test_every_model_is_wired", "test_no_hand_rolled_timing_remains", "test_warmup_and_iterations_agree_across_models",
                 "test_record_names_the_model", "test_record_call_parses", "test_measure_wraps_do_inference_only",
                 "test_onnx2trt_runs_in_the_shared_env", "test_no_unused_upstream_imports"):
        fn = getattr(wiring, name, None)
        if fn is None:
            continue
        sys.path.insert(0, os.path.join(REF, "tools"))
        try:
            fn()
        finally:
            sys.path.remove(os.path.join(REF, "tools"))
        ran += 1
    assert ran >= 6
    # the one rule that cannot hold here: the reference counts its thirteen model folders, this tree carries one
    failures = [f for f in wiring._failures if f != "found the model scripts"]
    assert not failures, failures
    assert "PASS  depth_anything_v2 measure() wraps only do_inference" in capsys.readouterr().out
    # and the shared constants are the reference's own
    ref_src = open(os.path.join(REF, "models", "depth_anything_v2", "onnx2trt.py"), encoding="utf-8").read()
    ours = open(SCRIPT, encoding="utf-8").read()
    for pat in (r"^\s*warmup\s*=\s*(\d+)", r"^\s*iteration\s*=\s*(\d+)", r"input_h\s*=\s*(\d+)"):
        assert re.search(pat, ref_src, re.M).group(1) == re.search(pat, ours, re.M).group(1)


# ------------------------------------------------------------------------------------------------ real harness, no GPU
def _overlay(tmp_path):
    """<tmp>/core = the reference's untouched modules (symlinks) + the two replacement files; <tmp>/models/... = our script;
    <tmp>/data -> the reference's data/."""
    core = tmp_path / "core"
    core.mkdir()
    for name in os.listdir(os.path.join(REF, "core")):
        if name.endswith(".py") and name not in ("common.py", "common_runtime.py"):
            os.symlink(os.path.join(REF, "core", name), core / name)
    for name in ("common.py", "common_runtime.py"):
        (core / name).write_text(open(os.path.join(ROOT, "integration", "core", name)).read())
    mdir = tmp_path / "models" / "depth_anything_v2"
    mdir.mkdir(parents=True)
    (mdir / "onnx2trt.py").write_text(open(SCRIPT).read())
    os.symlink(os.path.join(REF, "data"), tmp_path / "data")
    return mdir / "onnx2trt.py"


@needs_ref
def test_script_runs_against_the_real_harness_up_to_the_engine(tmp_path, lib):
    """Reference's core.bench / core.preprocess / data/example.jpg, our core.common: everything up to the engine build is the
    reference's own code path; the build then needs the B200 and says so (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("the GPU variant of this test runs the whole script")
    script = _overlay(tmp_path)
    sys.path.insert(0, ROOT)
    from oracle import dav2_torch as O
    from monocular_depth_estimation_trt_b200 import weights as W
    sd = O.init_state_dict("vits", seed=0)
    model_dir = tmp_path / "onnx"
    model_dir.mkdir()
    W.save(str(model_dir / "depth_anything_v2_vits_518x518_metric_hypersim.mdew"), sd, W.describe("vits", 518, 518, 20.0))
    code = (f"import sys, runpy; sys.path.insert(0, {ROOT!r}); ns = runpy.run_path({str(script)!r}); "
            f"import core.bench, core.preprocess, core.common; "
            f"print('BENCH', core.bench.__file__); print('PP', core.preprocess.__file__); print('COMMON', core.common.get_engine.__module__); "
            f"ns['main'](out_dir={str(tmp_path / 'reports')!r}, model_dir={str(model_dir)!r}, results_dir={str(tmp_path / 'results')!r})")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert os.path.realpath(os.path.join(REF, "core", "bench.py")) in os.path.realpath(re.search(r"BENCH (\S+)", out).group(1))
    assert "COMMON monocular_depth_estimation_trt_b200.common" in out
    assert "[MDET] original shape : (2268, 3024, 3)" in out and "[MDET] after preprocess shape : (1, 3, 518, 518)" in out
    assert "[MDET] Build engine" in out
    assert r.returncode != 0 and "RuntimeError: [MDET] mde_engine_finalize failed" in out, out[-2000:]


@needs_ref
def test_restated_record_matches_the_references(tmp_path, monkeypatch):
    from oracle import harness_np as H
    bench, _ = H.reference_modules()
    rng = np.random.default_rng(3)
    samples = [float(v) for v in rng.gamma(5.0, 0.7, 100)]
    depth = rng.uniform(0.5, 9.0, (48, 64)).astype(np.float32)
    kw = dict(warmup=20, precision="fp16", profile="bench", input_h=518, input_w=518, backend="mde_b200", encoder="vits", notes="n")
    theirs = bench.record("depth_anything_v2", samples, outputs={"depth": depth}, out_dir=str(tmp_path / "a"), echo=False, **kw).to_dict()
    monkeypatch.setattr(H, "reference_modules", lambda: (None, None))
    ours = H.record("depth_anything_v2", samples, outputs={"depth": depth}, out_dir=str(tmp_path / "b"), echo=False, **kw)
    assert os.listdir(tmp_path / "a") == os.listdir(tmp_path / "b") == ["depth_anything_v2_518x518_bench_single_fp16.json"]
    assert set(theirs) == set(ours)
    for key in theirs:
        if key not in ("timestamp", "host"):
            assert theirs[key] == ours[key], key
    assert bench.load(str(tmp_path / "b" / "depth_anything_v2_518x518_bench_single_fp16.json"))["stats"] == theirs["stats"]


@needs_ref
def test_records_written_on_the_b200_load_in_the_references_tooling():
    """profiles/r02_reports_bench/*.json were written by the integration script on a B200 (the GPU variant below); the
    reference's loader accepts them and its comparison renderer puts the model in its table."""
    d = os.path.join(ROOT, "profiles", "r02_reports_bench")
    if not os.path.isdir(d) or not os.listdir(d):
        pytest.skip("no GPU-made records committed yet")
    from oracle import harness_np as H
    bench, _ = H.reference_modules()
    runs = bench.load_all(d)
    assert runs and all(r["schema"] == 1 and r["stats"]["iterations"] == 100 and r["warmup"] == 20 for r in runs)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "tools"))
    try:
        compare = _load(os.path.join(REF, "tools", "compare.py"), "ref_compare")
    finally:
        sys.path.remove(os.path.join(REF, "tools"))
        sys.path.remove(REF)
    text = compare.render(runs)
    assert "depth_anything_v2" in text and f"{runs[0]['stats']['mean_ms']:.2f}" in text


# ------------------------------------------------------------------------------------------------ GPU: the whole script
@pytest.mark.gpu
@pytest.mark.parametrize("encoder", ["vits"])
def test_script_end_to_end_on_the_gpu(tmp_path, lib, encoder, monkeypatch):
    import torch
    import torch.nn.functional as F
    import refsetup as R
    export = _load(os.path.join(ROOT, "integration", "models", "depth_anything_v2", "onnx_export.py"), "mdet_export")
    model_dir = tmp_path / "onnx"
    path = export.main(["--seeded-init", "--encoder", encoder, "--out-dir", str(model_dir)])
    assert os.path.exists(path)
    monkeypatch.setenv("MDET_ENCODER", encoder)
    if HAVE_REF:
        script = _overlay(tmp_path)             # a GPU host that also has the checkout: the real harness all the way
    else:
        script = SCRIPT
    for m in [k for k in sys.modules if k == "core" or k.startswith("core.")]:
        del sys.modules[m]
    ns = _load(str(script), "mdet_onnx2trt")
    reports = tmp_path / "reports"
    depth, model_grid, x = ns.main(out_dir=str(reports), model_dir=str(model_dir), results_dir=str(tmp_path / "results"))
    # the engine saw what core/preprocess.py produces for this frame, and answers like the fp32 oracle
    sd, x_ref, ref_depth, _ = R.reference(encoder)
    if not HAVE_REF:
        assert np.array_equal(x, x_ref.numpy())             # synthetic frame seed 0 (no data/example.jpg on this host)
        m = R.compare_depth(ref_depth[0].numpy(), model_grid)
        assert m["abs_rel"] <= 2e-3 and m["max_rel"] <= 1e-2, m
        want = torch.clamp(F.interpolate(ref_depth[:, None], (480, 640), mode="bilinear", align_corners=True)[0, 0], 1e-3, 1e3).numpy()
        m2 = R.compare_depth(want, depth)
        assert m2["abs_rel"] <= 2e-3 and m2["max_rel"] <= 1e-2, m2
    files = os.listdir(reports)
    assert files == ["depth_anything_v2_518x518_bench_single_fp16.json"]
    rec = json.load(open(reports / files[0]))
    assert rec["schema"] == 1 and rec["model"] == "depth_anything_v2" and rec["warmup"] == 20 and rec["stats"]["iterations"] == 100
    assert rec["backend"] == "mde_b200" and rec["outputs"]["depth"]["shape"] == list(depth.shape) and rec["outputs"]["depth"]["nonfinite"] == 0
    assert len(rec["samples_ms"]) == 100 and rec["stats"]["p50_ms"] > 0
    keep = os.path.join(ROOT, "gpurun_out", "reports_bench")
    os.makedirs(keep, exist_ok=True)
    with open(os.path.join(keep, files[0]), "w") as f:
        json.dump(rec, f, indent=2)
