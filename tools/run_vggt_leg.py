import json, os, sys
sys.path.insert(0, "/root/repo")
import torch
torch.cuda.set_device(0)
from monocular_depth_estimation_trt_b200 import build
build.build()
import bench_partitioned as BP
print(json.dumps(BP.vggt_model(1, 0, 0, "fp16"), indent=1))
