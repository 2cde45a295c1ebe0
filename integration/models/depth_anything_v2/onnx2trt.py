# Depth Anything V2 on the B200 runtime: the `build` + `bench` stage script, shaped like the reference's
# models/depth_anything_v2/onnx2trt.py (:51-127) after the switch INTEGRATION.md describes.  What changed against it:
# no `import tensorrt`, the model file is the exported `.mdew` weights container instead of an ONNX graph, and the figures
# (matplotlib) are left to the caller.  Everything the reference's contract tests look at is where it was:
# `pp.preprocess_for(raw_img, 'depth_anything_v2', (input_h, input_w))`, `get_engine(...)` as a context manager,
# `common.allocate_buffers`, `bench.measure(...)` around `common.do_inference` ONLY, the script's post-processing outside the
# timed call, one `bench.record('depth_anything_v2', ...)`.
import os
import sys

# repository root = the directory that holds core/ (walk up, as the reference's scripts do)
_R = os.path.dirname(os.path.abspath(__file__))
while not os.path.isdir(os.path.join(_R, "core")) and os.path.dirname(_R) != _R:
    _R = os.path.dirname(_R)
sys.path.insert(1, _R)

import cv2
import numpy as np
import torch
import torch.nn.functional as F

from core import common
from core.common import *          # noqa: F401,F403
from core import bench
from core import preprocess as pp

CUR_DIR = os.path.dirname(os.path.abspath(__file__))
DEVICE = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
print(f"[MDET] using device: {DEVICE}")


def load_frame(root):
    """data/example.jpg of the checkout, or -- where there is none -- the reference's synthetic-input convention
    (tests/test_preprocess.py:47-51: seeded uint8 noise, 480 x 640)."""
    path = os.path.join(root, "data", "example.jpg")
    img = cv2.imread(path) if os.path.exists(path) else None
    if img is None:
        img = np.random.default_rng(0).integers(0, 256, (480, 640, 3), dtype=np.uint8)
        path = "synthetic seed 0"
    return img, path


def main(out_dir=None, model_dir=None, results_dir=None):
    results_dir = results_dir or os.path.join(CUR_DIR, "results")
    os.makedirs(results_dir, exist_ok=True)

    input_h = 518
    input_w = 518

    raw_img, image_path = load_frame(_R)
    print(f"[MDET] original shape : {raw_img.shape} ({image_path})")
    batch_images, geom = pp.preprocess_for(raw_img, 'depth_anything_v2', (input_h, input_w))
    print(f"[MDET] after preprocess shape : {batch_images.shape}")

    precision = os.environ.get("MDET_PRECISION", "fp16")   # 'fp16' (the reference's build target) or 'bf16'
    encoder = os.environ.get("MDET_ENCODER", "vits")       # 'vits' | 'vitb' | 'vitl'
    metric_model = True
    dataset = "hypersim"
    model_name = f"depth_anything_v2_{encoder}_{input_h}x{input_w}"
    model_name = f"{model_name}_metric_{dataset}" if metric_model else model_name
    model_dir = model_dir or os.path.join(CUR_DIR, "onnx")
    model_path = os.path.join(model_dir, f"{model_name}.mdew")            # written by onnx_export.py (stage `export`)
    engine_file_path = os.path.join(CUR_DIR, "engine", f"{model_name}_{precision}.engine")
    os.makedirs(os.path.dirname(engine_file_path), exist_ok=True)

    input_shape = batch_images.shape
    output_shape = (1, batch_images.shape[2], batch_images.shape[3])
    print(f"[MDET] engine input shape : {input_shape}")
    print(f"[MDET] engine output shape : {output_shape}")

    iteration = 100
    warmup = 20
    with get_engine(model_path, engine_file_path, precision, None) as engine, \
            engine.create_execution_context() as context:
        inputs, outputs, bindings, stream = common.allocate_buffers(engine, output_shape, profile_idx=0)
        inputs[0].host = batch_images

        engine_outputs, samples = bench.measure(
            lambda: common.do_inference(context, engine=engine, bindings=bindings, inputs=inputs, outputs=outputs, stream=stream),
            warmup=warmup, iterations=iteration)

        print("[MDET] Post process")
        model_grid = np.array(engine_outputs[0], dtype=np.float32).reshape(output_shape)
        depth = torch.from_numpy(model_grid)
        depth = F.interpolate(depth[:, None], (geom.src_h, geom.src_w), mode="bilinear", align_corners=True)[0, 0]
        depth = torch.clamp(depth, min=1e-3, max=1e3).numpy()

        bench.record('depth_anything_v2', samples, warmup=warmup, precision=precision, profile='bench',
                     input_h=input_h, input_w=input_w, outputs={'depth': depth}, backend="mde_b200", encoder=encoder,
                     notes=f"encoder={encoder}, weights={'metric_' + dataset if metric_model else 'relative'}",
                     out_dir=out_dir)
        print(f"[MDET] max : {depth.max():0.5f} , min : {depth.min():0.5f}")
        common.free_buffers(inputs, outputs, stream)

    np.savez_compressed(os.path.join(results_dir, f"example_{model_name}_b200"), depth=depth, model_grid=model_grid[0])
    return depth, model_grid[0], batch_images


if __name__ == '__main__':
    main()
