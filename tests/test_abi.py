"""The C ABI: every symbol include/mde_b200.h declares is exported and bound, struct layouts agree between
the header (compiled with gcc) and the ctypes mirror, and without a GPU every compute entry point fails
loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from monocular_depth_estimation_trt_b200 import _lib, engine as E, weights as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mde_b200.h")


def has_gpu():
    import torch
    return torch.cuda.is_available()


def test_header_symbols_are_exported_and_bound(lib):
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(mde_[a-z0-9_]+)\s*\(", text))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert lib.mde_abi_version() == _lib.MDE_ABI_VERSION == 2


def test_struct_layout_matches_header(tmp_path):
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "mde_b200.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(mde_engine_desc), offsetof(mde_engine_desc, norm_mean),"
        " offsetof(mde_engine_desc, max_depth), offsetof(mde_engine_desc, device), sizeof(mde_epilogue),"
        " offsetof(mde_epilogue, ld_out), offsetof(mde_epilogue, d_head_w), offsetof(mde_epilogue, d_head_out));return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(_lib.EngineDesc), _lib.EngineDesc.norm_mean.offset, _lib.EngineDesc.max_depth.offset,
            _lib.EngineDesc.device.offset, C.sizeof(_lib.Epilogue), _lib.Epilogue.ld_out.offset,
            _lib.Epilogue.d_head_w.offset, _lib.Epilogue.d_head_out.offset]
    assert got == want


def test_engine_description_and_io_contract(lib):
    """models/depth_anything_v2/spec.json: input 'input' float32 NCHW [1,3,518,518]; output 'output' [1,518,518]."""
    meta = W.describe("vits", 518, 518, 20.0)
    eng = E.Engine(E.make_desc(meta, precision="fp16", batch=1), meta)
    assert eng.num_io_tensors == 2
    assert [eng.get_tensor_name(i) for i in range(2)] == ["input", "output"]
    assert eng.get_tensor_shape("input") == (1, 3, 518, 518) and eng.get_tensor_dtype("input") == np.float32
    assert eng.get_tensor_shape("output") == (1, 518, 518) and eng.get_tensor_dtype("output") == np.float32
    assert eng.get_tensor_mode("input") == E.TensorIOMode.INPUT and eng.get_tensor_mode("output") == E.TensorIOMode.OUTPUT
    assert eng.get_tensor_profile_shape("input", 0)[-1] == (1, 3, 518, 518)
    assert eng.workspace_bytes > 50e6
    with pytest.raises(KeyError):
        eng.get_tensor_shape("depth")
    eng.close()
    big = E.Engine(E.make_desc(W.describe("vitl", 518, 518, 20.0), precision="bf16", batch=64, input_mode="u8_hwc",
                               max_src_hw=(480, 640)), {})
    assert big.get_tensor_shape("input") == (64, 480, 640, 3) and big.get_tensor_dtype("input") == np.uint8
    assert 8e9 < big.workspace_bytes < 15e9           # one arena with scoped reuse (19.8 GiB before the scopes), 180 GB of HBM3e
    big.close()


@pytest.mark.parametrize("bad", [dict(precision="fp32"), dict(input_mode="nhwc")])
def test_bad_descriptions_raise_in_python(bad):
    with pytest.raises(ValueError):
        E.make_desc(W.describe("vits"), **bad)


def test_bad_descriptions_are_refused_by_the_library(lib):
    meta = W.describe("vits")
    for mutate in (lambda d: setattr(d, "num_heads", 5), lambda d: setattr(d, "input_h", 520),
                   lambda d: setattr(d, "batch", 0), lambda d: setattr(d, "struct_size", 8),
                   lambda d: d.taps.__setitem__(3, 99)):
        d = E.make_desc(meta)
        mutate(d)
        with pytest.raises(RuntimeError):
            E.Engine(d)
    d = E.make_desc(meta, input_mode="u8_hwc")        # uint8 input without a maximum source size
    with pytest.raises(RuntimeError, match="max_src"):
        E.Engine(d)


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_fallback(lib):
    meta = W.describe("vits")
    eng = E.Engine(E.make_desc(meta), meta)
    with pytest.raises(RuntimeError):
        eng.finalize()                                   # needs the device; there is no CPU path
    with pytest.raises(RuntimeError):
        eng.create_execution_context()
    ep = _lib.Epilogue()
    buf = (C.c_char * 4096)()
    rc = lib.mde_k_gemm(1, C.addressof(buf) & ~15, 8, 64, 64, C.addressof(buf) & ~15, 8, 64, C.byref(ep), None)
    assert rc != 0 and _lib.last_error()
    eng.close()


def test_product_code_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package, nor bench.py's own arm, may route through it."""
    pkg = os.path.join(ROOT, "monocular_depth_estimation_trt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f"{f} imports the oracle"
    bench = open(os.path.join(ROOT, "bench.py"), encoding="utf-8").read()
    ours = bench[bench.index("def run_ours"):bench.index("def main")]
    # our arm takes weights/images from the oracle's seeded recipe, and uses the oracle as the CHECKER after the timed
    # regions (parity of the benchmarked batch + the cpu_baseline leg, the reference's measure / stats for the latency
    # protocol); the engine calls in between must not touch it
    assert ours.count("oracle_setup(") == 1 and ours.count("oracle_depths(") == 1
    timed = ours[ours.index("# ---------------- device-resident arm"):ours.index("# ---------------- batch-1 latency")]
    assert "oracle" not in timed and "H." not in timed
