"""Generate tests/golden/harness_golden.json with the REFERENCE's own functions (run where /root/reference is mounted):

    python oracle/make_golden_harness.py

core.golden.compare on seeded depth-like maps (with NaN / Inf / non-positive pixels mixed in) and core.bench.Bench.stats on
seeded sample lists.  tests/test_oracle_harness.py checks oracle/harness_np.py against these records on every host.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "harness_golden.json")


def depth_pair(seed: int):
    """The seeded pair every consumer regenerates: a positive map, a perturbed copy, and a few invalid pixels."""
    rng = np.random.default_rng(seed)
    ref = rng.uniform(0.3, 20.0, (37, 53))
    got = ref * (1.0 + rng.normal(0.0, 3e-3, ref.shape))
    if seed % 2:
        ref[0, :5] = np.nan
        got[1, :4] = np.inf
        ref[2, :3] = 0.0
        got[3, :2] = -1.0
    return ref, got


def samples(seed: int, n: int):
    return [float(v) for v in np.random.default_rng(100 + seed).gamma(4.0, 1.0, n)]


def main():
    sys.path.insert(0, "/root/reference")
    from core import bench, golden
    rec = {"compare": [], "stats": []}
    for seed in range(4):
        ref, got = depth_pair(seed)
        rec["compare"].append({"seed": seed, "entry": golden.compare({"depth": ref}, {"depth": got})["depth"]})
    for seed, n in enumerate([1, 2, 10, 100, 101]):
        rec["stats"].append({"seed": seed, "n": n, "warmup": 3, "stats": bench.Bench(model="_", samples_ms=samples(seed, n), warmup=3).stats()})
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(rec, f, indent=1, sort_keys=True)
    print(OUT)


if __name__ == "__main__":
    main()
