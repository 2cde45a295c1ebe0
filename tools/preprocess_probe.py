import os, sys
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kutil as K
src = torch.randint(0, 256, (64, 480, 640, 3), dtype=torch.uint8, device="cuda")
def timed(fn, reps=50):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps*1000
t = timed(lambda: K.preprocess_u8("fp16", src, 518, 518, want_nchw=False))
by = 64*480*640*3 + 64*1369*640*2
print(f"preprocess_u8 B=64 480x640 -> 518x518 im2col: {t:.1f} us, {by/t/1e3:.0f} GB/s (src bytes + im2col rows)")
