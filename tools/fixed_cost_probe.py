import os, sys, ctypes as C
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kutil as K
from monocular_depth_estimation_trt_b200 import _lib
lib=_lib.load()
dt=torch.float16; dev="cuda"
def timed(fn, reps=50):
    """device time per launch: `reps` launches recorded into one CUDA graph (the host needs ~15 us per ctypes call, more than
    most of these kernels run), replayed and timed with events"""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps*1000
z=torch.zeros(1024,device=dev)
print("torch tiny kernel (z.add_) %.1f us" % timed(lambda: z.add_(1.0)))
for (M,N,Kd) in [(128,256,64),(128,256,1024),(1370,1024,64),(1370,1024,256),(1370,1024,1024),(1370,1024,4096),(1370,3072,1024)]:
    A=(torch.randn(M,Kd,device=dev)*0.5).to(dt); B=(torch.randn(N,Kd,device=dev)*0.5).to(dt)
    out=torch.empty(M,N,dtype=dt,device=dev)
    ep=K.epilogue(bias=torch.randn(N,device=dev), out=out, ld_out=N)
    t=timed(lambda: K.gemm("fp16",A,B,ep))
    x=torch.randn(M,N,device=dev)
    ep2=K.epilogue(bias=torch.randn(N,device=dev), gamma=torch.rand(N,device=dev)*0.01, x=x, accumulate_x=True, ld_out=N)
    t2=timed(lambda: K.gemm("fp16",A,B,ep2))
    print(f"gemm M={M} N={N} K={Kd}: plain-out {t:.1f} us   residual-reduce {t2:.1f} us")
x=torch.randn(1370,1024,device=dev); w=torch.randn(1024,device=dev); b=torch.randn(1024,device=dev)
print("layernorm 1370x1024 %.1f us" % timed(lambda: K.layernorm("fp16",x,w,b)))
qkv=torch.randn(1370,3072,device=dev).to(dt)
print("attention b1 %.1f us" % timed(lambda: K.attention("fp16",qkv,1,1370,16)))
qkv=torch.randn(128,3072,device=dev).to(dt)
print("attention 128 tokens %.1f us" % timed(lambda: K.attention("fp16",qkv,1,128,16)))
