"""CPU emulation of the precision plan of the CUDA path (test helper, not a product path).

Every tensor the kernels store as bf16 is rounded to bf16 here at the same place; every
accumulation, LayerNorm statistic, softmax, GELU and the residual stream stay fp32 --
exactly what DESIGN.md section "precision plan" says the kernels do.  bf16 x bf16
products are exact in fp32, so rounding the operands and running an fp32 matmul models a
tensor-core GEMM with fp32 accumulation up to summation order.

Used by tests/test_precision_plan.py to show the plan can meet north_star's gate
(max-rel <= 1e-2, AbsRel <= 2e-3) against the fp32 oracle before any GPU is involved.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from oracle import dav2_torch as O


def r(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def _lin(x, w, b):
    return F.linear(r(x), r(w), b)


def _conv(x, w, b=None, **kw):
    return F.conv2d(r(x), r(w), b, **kw)


@torch.no_grad()
def forward(sd, x, encoder="vits", max_depth=20.0, act_round=r, trace=None):
    cfg = O.MODEL_CONFIGS[encoder]
    D, H = cfg["embed_dim"], cfg["num_heads"]
    hd = D // H
    B = x.shape[0]
    gh, gw = x.shape[-2] // O.PATCH, x.shape[-1] // O.PATCH
    t = F.conv2d(r(x), r(sd["pretrained.patch_embed.proj.weight"]),
                 sd["pretrained.patch_embed.proj.bias"], stride=O.PATCH).flatten(2).transpose(1, 2)
    t = torch.cat([sd["pretrained.cls_token"].expand(B, -1, -1), t], 1)
    t = t + O.interpolate_pos_embed(sd["pretrained.pos_embed"], gh, gw)          # fp32 residual stream
    taps = []
    for i in range(cfg["depth"]):
        p = f"pretrained.blocks.{i}."
        y = act_round(F.layer_norm(t, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], O.LN_EPS))
        qkv = act_round(_lin(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]))
        N = qkv.shape[1]
        qkv = qkv.reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
        s = (qkv[0] @ qkv[1].transpose(-2, -1)) * hd ** -0.5
        m = s.amax(-1, keepdim=True)
        pexp = torch.exp(s - m)
        a = (act_round(pexp) @ qkv[2]) / pexp.sum(-1, keepdim=True)                # P rounded for the PV MMA
        a = act_round(a.transpose(1, 2).reshape(B, N, D))
        t = t + sd[p + "ls1.gamma"] * _lin(a, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        y = act_round(F.layer_norm(t, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], O.LN_EPS))
        y = act_round(F.gelu(_lin(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])))
        t = t + sd[p + "ls2.gamma"] * _lin(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        if trace is not None:
            trace[f"block{i}"] = t
        if i in cfg["taps"]:
            y = F.layer_norm(t, (D,), sd["pretrained.norm.weight"], sd["pretrained.norm.bias"], O.LN_EPS)
            taps.append(act_round(y[:, 1:]))

    h = "depth_head."
    A = act_round
    l = []
    for i, tk in enumerate(taps):
        f = tk.permute(0, 2, 1).reshape(B, D, gh, gw)
        f = A(_conv(f, sd[h + f"projects.{i}.weight"], sd[h + f"projects.{i}.bias"]))
        if i == 0:
            f = A(F.conv_transpose2d(r(f), r(sd[h + "resize_layers.0.weight"]), sd[h + "resize_layers.0.bias"], stride=4))
        elif i == 1:
            f = A(F.conv_transpose2d(r(f), r(sd[h + "resize_layers.1.weight"]), sd[h + "resize_layers.1.bias"], stride=2))
        elif i == 3:
            f = A(_conv(f, sd[h + "resize_layers.3.weight"], sd[h + "resize_layers.3.bias"], stride=2, padding=1))
        l.append(f)
    rr = [A(_conv(l[i], sd[h + f"scratch.layer{i + 1}_rn.weight"], None, padding=1)) for i in range(4)]

    def rcu(pre, xin, extra=None):
        y = A(F.relu(_conv(F.relu(xin), sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)))
        y = _conv(y, sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1) + xin
        if extra is not None:
            y = y + extra
        return A(y)

    def fusion(i, x0, x1, size):
        rn = h + f"scratch.refinenet{i}."
        out = x0 if x1 is None else rcu(rn + "resConfUnit1.", x1, extra=x0)
        out = rcu(rn + "resConfUnit2.", out)
        # 1x1 out_conv commutes with bilinear interpolation (both linear, weights sum to 1):
        # the CUDA path runs it at the low resolution, then upsamples.
        out = A(_conv(out, sd[rn + "out_conv.weight"], sd[rn + "out_conv.bias"]))
        if size is None:
            out = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
        else:
            out = F.interpolate(out, size=size, mode="bilinear", align_corners=True)
        return A(out)

    if trace is not None:
        for i in range(4):
            trace[f"layer{i + 1}_rn"] = rr[i]
    p4 = fusion(4, rr[3], None, rr[2].shape[2:])
    p3 = fusion(3, p4, rr[2], rr[1].shape[2:])
    p2 = fusion(2, p3, rr[1], rr[0].shape[2:])
    p1 = fusion(1, p2, rr[0], None)
    if trace is not None:
        trace["path_1"] = p1
    out = A(_conv(p1, sd[h + "scratch.output_conv1.weight"], sd[h + "scratch.output_conv1.bias"], padding=1))
    out = A(F.interpolate(out, (gh * O.PATCH, gw * O.PATCH), mode="bilinear", align_corners=True))
    out = F.relu(_conv(out, sd[h + "scratch.output_conv2.0.weight"], sd[h + "scratch.output_conv2.0.bias"], padding=1))
    z = F.conv2d(out, sd[h + "scratch.output_conv2.2.weight"], sd[h + "scratch.output_conv2.2.bias"])   # fp32 in the epilogue
    d = F.relu(z) if max_depth is None else torch.sigmoid(z) * max_depth
    return d.squeeze(1)
