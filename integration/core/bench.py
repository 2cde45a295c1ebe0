"""`core.bench` where the reference checkout is absent: the pinned restatement (oracle/harness_np.py), same names."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from oracle.harness_np import SCHEMA, collect_env, measure, pct, stats, summarize_outputs  # noqa: E402,F401
from oracle import harness_np as _H  # noqa: E402

REPORTS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reports", "bench")
INPUTS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "reports", "inputs")


def record(model, samples_ms, *, out_dir=None, **kw):
    return _H.record(model, samples_ms, out_dir=out_dir or REPORTS, inputs_dir=INPUTS, **kw)
