import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch, kutil as K
torch.manual_seed(0)
for prec in ("fp16",):
    dt = K.TORCH_DT[prec]
    for ntok in (32, 64, 96, 97, 128, 192, 193, 288, 384, 480, 1370):
        for v in ("k0", "k1", "k2"):
            B, H = 1, 1
            qkv = torch.randn(B * ntok, 3 * H * 64, device="cuda").to(dt)
            out = K.attention(prec, qkv, B, ntok, H, v).float()
            q, k, vv = qkv.float().reshape(B, ntok, 3, H, 64).permute(2, 0, 3, 1, 4)
            ref = (torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ vv).transpose(1, 2).reshape(B * ntok, H * 64)
            err = (out - ref).abs().max(dim=1).values
            bad = (err > 0.02).nonzero().flatten()
            print(prec, ntok, v, "max err %.3e" % float(err.max()), "bad rows", int(bad.numel()), bad[:6].tolist(), flush=True)
