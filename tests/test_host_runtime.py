"""Host logic of the reference-shaped API (no GPU): error contracts of core/common_runtime.py and
core/common.py that the B200 mirror keeps."""
import os

import numpy as np
import pytest

from monocular_depth_estimation_trt_b200 import common, common_runtime as CR, weights as W
from oracle import dav2_torch as O


def test_reexports_match_reference_surface():
    for name in ("get_engine", "allocate_buffers", "do_inference", "free_buffers", "HostDeviceMem", "StageTimer",
                 "cuda_call", "check_cuda_err", "memcpy_host_to_device", "memcpy_device_to_host", "GiB",
                 "engine_staleness"):
        assert hasattr(common, name), name
    assert CR.StageTimer.STAGES == ("h2d_ms", "compute_ms", "d2h_ms")
    assert common.GiB(2) == 2 << 30


def test_cuda_call_contract():
    from cuda.bindings import runtime as cudart
    ok = cudart.cudaError_t.cudaSuccess
    assert CR.cuda_call((ok, 7)) == 7
    assert CR.cuda_call((ok, 1, 2)) == (1, 2)
    with pytest.raises(RuntimeError, match="Cuda Runtime Error"):
        CR.cuda_call((cudart.cudaError_t.cudaErrorInvalidValue, None))
    with pytest.raises(RuntimeError, match="Unknown error type"):
        CR.check_cuda_err(42)


def test_shape_override_rules():
    f = CR._shape_override
    assert f("output", (1, 518, 518), None) is None
    assert f("output", (1, 518, 518), {"output": (1, 4, 4)}) == (1, 4, 4)
    assert f("other", (1, 518, 518), {"output": (1, 4, 4)}) is None
    assert f("output", (1, 518, 518), (1, 8, 8)) is None            # engine shape usable: keep it
    assert f("output", (-1, 518, 518), (1, 8, 8)) == (1, 8, 8)      # dynamic
    assert f("output", (1,), (1, 8, 8)) == (1, 8, 8)                # degenerate volume


def test_engine_staleness_table(tmp_path):
    eng, fp = tmp_path / "a.engine", tmp_path / "a.fingerprint"
    assert common.engine_staleness(str(eng), str(fp), "x", True) == "no engine file"
    eng.write_text("{}")
    assert common.engine_staleness(str(eng), str(fp), "x", False) is None          # engine-only deployment
    assert common.engine_staleness(str(eng), str(fp), None, True) is None
    assert common.engine_staleness(str(eng), str(fp), "x", True) == "no fingerprint recorded"
    fp.write_text("y")
    assert common.engine_staleness(str(eng), str(fp), "x", True) == "weights or build options changed"
    fp.write_text("x")
    assert common.engine_staleness(str(eng), str(fp), "x", True) is None


def test_get_engine_argument_errors(tmp_path, lib):
    with pytest.raises(FileNotFoundError):
        common.get_engine(str(tmp_path / "nope.mdew"), str(tmp_path / "e.engine"), "fp16")
    path = str(tmp_path / "m.mdew")
    W.save(path, O.init_state_dict("vits", 0), W.describe("vits"))
    with pytest.raises(ValueError, match="dynamic"):
        common.get_engine(path, "", "fp16", [[1, 3, 518, 518]] * 3)
    with pytest.raises(ValueError, match="precision"):
        common.get_engine(path, "", "fp32")
    fp_a = common._engine_fingerprint(path, "fp16", 2, None, False, None)
    fp_b = common._engine_fingerprint(path, "bf16", 2, None, False, None)
    assert fp_a != fp_b and fp_a.splitlines()[0] == W.file_sha256(path)
