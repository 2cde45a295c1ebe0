#!/usr/bin/env python
"""Offline analysis of gpurun_out/attn_trace.npz (tools/attn_trace_dump.py): how the exponential phase of a softmax warp depends
on what the OTHER CTA resident on the same SM is doing.  For every full key tile of every traced warp the duration of
"S in registers -> exponentials done" is set against the share of that interval during which the same-numbered warp (same SM
sub-partition) of a co-resident CTA was in ITS exponential phase."""
import sys, numpy as np
tr = np.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/attn_trace.npz")["trace"]   # [cta][warp][slot]
ncta = tr.shape[0]
smid = tr[:, :, 63]
ntile = 10                                      # full key tiles of N = 1370
b = 2 + 5 * np.arange(ntile)
t0 = tr[:, :, :][:, :, b + 1]                   # S in registers
t1 = tr[:, :, :][:, :, b + 2]                   # exponentials done
dur = (t1 - t0).astype(np.float64)
ov = np.zeros_like(dur)
for w in range(4):
    for sm in np.unique(smid[:, w]):
        ids = np.nonzero(smid[:, w] == sm)[0]
        a0, a1 = t0[ids, w], t1[ids, w]          # [n][tile]
        for k, c in enumerate(ids):
            o0 = np.delete(a0, k, 0).reshape(-1); o1 = np.delete(a1, k, 0).reshape(-1)
            for j in range(ntile):
                ov[c, w, j] = np.clip(np.minimum(o1, a1[k, j]) - np.maximum(o0, a0[k, j]), 0, None).sum()
d, o = dur.reshape(-1), ov.reshape(-1)
ok = (d > 0) & (d < 20000)
d, o = d[ok], o[ok]
A = np.stack([np.ones_like(o), o], 1)
c = np.linalg.lstsq(A, d, rcond=None)[0]
print(f"{ok.sum()} (warp, key tile) samples on {len(np.unique(smid))} SMs; CTAs per SM {ncta / len(np.unique(smid)):.1f}")
print(f"exponential phase: mean {d.mean():.0f} clk; least squares  dur = {c[0]:.0f} + {c[1]:.3f} * overlap")
fr = o / d
for lo, hi in ((0, .1), (.1, .3), (.3, .5), (.5, .7), (.7, .9), (.9, 1.01)):
    m = (fr >= lo) & (fr < hi)
    print(f"  overlap share {lo:.1f}-{hi:.1f}: n = {m.sum():6d}   mean duration {d[m].mean():.0f} clk")
