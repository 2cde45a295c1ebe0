"""oracle/harness_np.py (the restated core/golden.compare and core/bench timing statistics) against golden records made
by the reference's own functions, and against the live reference modules where the checkout is mounted."""
import json
import math
import os

import numpy as np
import pytest

from oracle import harness_np as H
from oracle.make_golden_harness import depth_pair, samples

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "harness_golden.json")


def close(a, b):
    if isinstance(a, float) or isinstance(b, float):
        return a == b or math.isclose(a, b, rel_tol=1e-12)
    return a == b


def test_compare_matches_reference_golden_records():
    rec = json.load(open(GOLDEN))
    assert len(rec["compare"]) == 4
    for item in rec["compare"]:
        ref, got = depth_pair(item["seed"])
        ours = H.compare(ref, got)
        for key, val in item["entry"].items():
            assert close(val, ours[key]), (item["seed"], key, val, ours[key])
        assert ours["max_rel"] >= ours["abs_rel"] > 0


def test_stats_match_reference_golden_records(monkeypatch):
    monkeypatch.setattr(H, "reference_modules", lambda: (None, None))      # exercise the restatement, not the live module
    rec = json.load(open(GOLDEN))
    for item in rec["stats"]:
        ours = H.stats(samples(item["seed"], item["n"]), warmup=item["warmup"])
        assert ours == item["stats"], (item["n"], ours, item["stats"])


def test_against_live_reference_modules():
    bench, golden = H.reference_modules()
    if bench is None:
        pytest.skip("reference checkout not mounted")
    rng = np.random.default_rng(5)
    for _ in range(5):
        ref = rng.uniform(-1.0, 30.0, (20, 31))
        got = ref + rng.normal(0, 0.05, ref.shape)
        got[rng.integers(0, 20), rng.integers(0, 31)] = np.nan
        theirs = golden.compare({"d": ref}, {"d": got})["d"]
        ours = H.compare(ref, got)
        for key, val in theirs.items():
            assert close(val, ours[key]), (key, val, ours[key])
        s = [float(v) for v in rng.gamma(3.0, 2.0, int(rng.integers(1, 150)))]
        b = bench.Bench(model="_", samples_ms=s)
        for q in (0, 1, 50, 90, 99, 99.5, 100):
            assert H.pct(s, q) == b.pct(q)
    out, ms = H.measure(lambda: 7, warmup=2, iterations=5)
    assert out == 7 and len(ms) == 5 and all(m >= 0 for m in ms)


def test_compare_depth_reports_what_the_gates_read():
    ref, got = depth_pair(0)
    m = H.compare_depth(ref, got)
    assert {"abs_rel", "max_rel", "rel_mean", "corr", "compared", "positive"} <= set(m)
    assert m["compared"] == ref.size and m["corr"] > 0.999
    assert H.compare(ref, ref)["identical"] and H.compare(ref, got[:, :-1])["status"] == "shape_mismatch"
    assert H.compare(np.full((2, 2), np.nan), np.ones((2, 2)))["status"] == "no_finite_overlap"
