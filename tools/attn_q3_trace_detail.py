#!/usr/bin/env python
"""Per-key-tile detail of one CTA from gpurun_out/attn_q3_trace.npz: MMA issue stamps and the phases of all 12 softmax warps."""
import sys
import numpy as np
tr = np.load('gpurun_out/attn_q3_trace.npz')['trace']
cta = int(sys.argv[1]) if len(sys.argv) > 1 else 5
it = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t0 = tr[cta, 15, 0]
nkv = 15
per = 4 * nkv + 4
m = [tr[cta, 12 + t, it * 2 * nkv:(it + 1) * 2 * nkv] - t0 for t in range(3)]
for j in range(3, 6):
    print('j', j, 'S issued', [int(m[t][2 * j]) for t in range(3)], 'PV issued', [int(m[t][2 * j + 1]) for t in range(3)])
    for w in range(12):
        r = tr[cta, w, it * per + 4 * j:it * per + 4 * j + 4] - t0
        print('   w', w, r.tolist(), 'ld', int(r[1] - r[0]), 'exp', int(r[2] - r[1]), 'st', int(r[3] - r[2]))
