"""The precision plan of the CUDA path, emulated on the CPU (tests/bf16_emulation.py), against the fp32 oracle.

Records the finding that decides what the GPU gates are ("precision plan"):
  * fp16 operands + fp32 accumulation / statistics / residual stream meet north_star's gate
    (max-rel <= 1e-2, AbsRel <= 2e-3);
  * bf16 operands cannot: rounding ONLY the weights to bf16 already breaks it, so no kernel could pass."""
import torch

import bf16_emulation as E
import refsetup as R
from oracle import dav2_torch as O


def _round_weights(sd, fn):
    return {k: (fn(v) if v.dim() >= 2 and "pos_embed" not in k and "cls_token" not in k else v) for k, v in sd.items()}


def test_fp16_plan_meets_north_star_gate():
    sd, x, depth, _ = R.reference("vits")
    fn = lambda t: t.half().float()
    E.r = fn
    got = E.forward(sd, x, "vits", 20.0, act_round=fn)
    m = R.compare_depth(depth.numpy(), got.numpy())
    assert m["abs_rel"] <= 2e-3 and m["max_rel"] <= 1e-2, m


def test_bf16_weights_alone_exceed_the_gate():
    sd, x, depth, _ = R.reference("vits")
    bf = lambda t: t.to(torch.bfloat16).float()
    got = O.forward(_round_weights(sd, bf), x, "vits", 20.0)       # activations, accumulation: all fp32
    m = R.compare_depth(depth.numpy(), got.numpy())
    assert m["abs_rel"] > 2e-3 and m["max_rel"] > 1e-2, m           # the gate is unreachable with bf16 weights
    E.r = bf
    full = R.compare_depth(depth.numpy(), E.forward(sd, x, "vits", 20.0, act_round=bf).numpy())
    assert full["abs_rel"] <= 1.2e-2 and full["max_rel"] <= 1.2e-1, full   # the bf16 budget the GPU test uses
