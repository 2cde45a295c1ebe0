"""Stand-in for the reference's `core` package on hosts where its checkout is not mounted (the GPU box).

In a maintainer's tree `core/` is the reference's own package with `common.py` / `common_runtime.py` replaced by the two
import lines INTEGRATION.md section 1 lists; `bench.py`, `preprocess.py`, `golden.py`, `spec.py` stay untouched.  Here the
same four names resolve to: the two replacement files (identical to what the maintainer writes) and, for the untouched
modules, thin shims over the pinned restatements under oracle/ (tests/test_oracle_harness.py, tests/test_oracle_preprocess.py
pin them against the reference's own functions and golden records).  tests/test_integration.py builds the OTHER arrangement
-- the reference's real bench / preprocess / golden next to the replacement common files -- wherever /root/reference exists.
"""
