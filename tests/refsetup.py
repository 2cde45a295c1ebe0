"""Shared reference set-up for the model-level tests: seeded calibrated weights, the synthetic
input, and the oracle's fp32 output.  Test infrastructure only."""
from __future__ import annotations

import functools

import numpy as np
import torch

from oracle import dav2_torch as O
from oracle import preprocess_np as P


def synthetic_image(seed=0, h=480, w=640):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


@functools.lru_cache(maxsize=4)
def reference(encoder: str, h: int = 518, w: int = 518, max_depth: float = 20.0):
    """-> (state_dict, x [1,3,h,w] float32, oracle depth [1,h,w], trace dict)."""
    torch.manual_seed(0)
    x = torch.from_numpy(P.preprocess_stretch_imagenet(synthetic_image(0), h, w))
    sd = O.init_state_dict(encoder, seed=0)
    O.calibrate_head(sd, x, encoder)
    trace = {}
    depth = O.forward(sd, x, encoder, max_depth=max_depth, trace=trace)
    return sd, x, depth, trace


def compare_depth(ref: np.ndarray, got: np.ndarray) -> dict:
    """The reference's parity metrics (core/golden.py:101-174 `compare` -- the reference's own function where its checkout
    is mounted, the pinned restatement oracle/harness_np.py elsewhere), plus the max relative error north_star gates on."""
    from oracle import harness_np as H
    return H.compare_depth(ref, got)


@functools.lru_cache(maxsize=2)
def depth_pro_reference(encoder: str = "vits", features: int = 64, hook_taps=(2, 1)):
    """-> (state_dict, x [1,3,1536,1536], canonical inverse depth, fov_deg, trace) of the oracle's whole Depth Pro model on
    the reference's synthetic frame (seed 0, 480 x 640), weights seeded and calibrated (oracle/depth_pro_torch.py)."""
    from oracle import depth_pro_torch as DP
    x = DP.preprocess(synthetic_image(0), 1536)
    sd = DP.init_full_state_dict(encoder, features=features, seed=21)
    DP.calibrate_full(sd, x, encoder, hook_taps)
    trace = {}
    inv, fov = DP.full_forward(sd, x, encoder, hook_taps, trace)
    return sd, x, inv, fov, trace
