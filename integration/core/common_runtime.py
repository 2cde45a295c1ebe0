# core/common_runtime.py after the switch.
from monocular_depth_estimation_trt_b200.common_runtime import *  # noqa: F401,F403  HostDeviceMem, allocate_buffers, do_inference, free_buffers, StageTimer, cuda_call
