"""The partitioned configurations (SURVEY section 8 e rows 2 and 3) as pytest cases: one process per GPU under torchrun,
every rank's output checked against the UNSHARDED fp32 oracle by the scripts under tests/mgpu/ (exit code 0 and "ok": true
in their JSON line).  Skipped on boxes with fewer GPUs than the case needs; `pytest -m gpu` on a 2+ GPU box runs them:

    gpurun --gpus 2 -- python -m pytest tests/test_mgpu_gpu.py -m gpu -q
"""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_count() -> int:
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def torchrun(script: str, world: int, *args: str, timeout: int = 600) -> dict:
    if gpu_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {gpu_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(ROOT, "tests", "mgpu", script), *args]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, f"{' '.join(cmd)}\nrc={r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}"
    return json.loads(lines[-1])


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("gather", ["fused", "nccl"])
def test_vggt_aggregator_sharded_matches_unsharded_oracle(lib, world, gather):
    """Frames sharded over the ranks; K|V all-gather fused into the qk-norm + RoPE kernel's stores (or NCCL)."""
    out = torchrun("vggt_aggregator.py", world, "--check", "--gather", gather, "--precision", "fp16")
    assert out["ok"] and out["world"] == world and out["worst_rms_rel_per_layer_over_ranks"] < 1.2e-3, out


def test_vggt_aggregator_sharded_graph_replay(lib):
    out = torchrun("vggt_aggregator.py", 2, "--check", "--graph", "--precision", "fp16")
    assert out["ok"] and out["cuda_graph"], out


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("gather", ["fused", "nccl"])
def test_depth_pro_sharded_matches_unsharded_oracle(lib, world, gather):
    """35 crops sharded over the ranks, taps all-gathered by the kernel that produces them (or NCCL), decoder on every rank:
    every rank must hold the same map, inside north_star's gate against the oracle."""
    out = torchrun("depth_pro_full.py", world, "--check", "--gather", gather, "--precision", "fp16", "--reps", "3")
    assert out["ok"] and out["parity"]["abs_rel"] <= 2e-3 and out["parity"]["max_rel"] <= 1e-2, out


def test_sharded_patch_encoder_fused_equals_nccl(lib):
    """Depth Pro's patch-encoder stage alone: the fused gather is bit-identical to NCCL's on every rank and matches the oracle."""
    out = torchrun("sharded_patch_encoder.py", 2, "--check")
    assert out["fused_equals_nccl_on_every_rank"] and max(out["rms_rel_vs_oracle"]) < 1.2e-3, out


def test_sharded_global_attention_fused_equals_nccl(lib):
    """GEMM -> all-gather in one kernel (TMA stores into peer memory), hand-shake on the stream, attention over the gathered
    keys / values: bit-identical to the NCCL pipeline on every rank and within two 16-bit roundings of the fp32 reference."""
    out = torchrun("sharded_global_attention.py", 2, "--frames", "4", "--tokens", "333", "--heads", "6", "--precision", "fp16")
    assert out["fused_equals_nccl_on_every_rank"] and out["flags_equals_nccl_on_every_rank"], out
    assert out["rel_err_vs_fp32_reference"] < 8 * 1.25 * 2.0 ** -11, out
