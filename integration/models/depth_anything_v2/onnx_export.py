# Stage `export` for the B200 runtime (reference: models/depth_anything_v2/onnx_export.py:24-65 builds the upstream model,
# loads `checkpoints/depth_anything_v2_metric_hypersim_<encoder>.pth` and writes an ONNX graph).  Here the artefact is the
# `.mdew` weights container `get_engine` reads -- the checkpoint's tensors under their upstream names plus a description of
# the architecture -- so the build stage needs neither the upstream package nor an ONNX toolchain.
#
#   python onnx_export.py --checkpoint checkpoints/depth_anything_v2_metric_hypersim_vits.pth
#   python onnx_export.py --seeded-init            # no checkpoint at hand (tests): the oracle's seeded, calibrated weights
import argparse
import os
import sys

CUR_DIR = os.path.dirname(os.path.abspath(__file__))
_ROOT = CUR_DIR
while not os.path.isdir(os.path.join(_ROOT, "monocular_depth_estimation_trt_b200")) and os.path.dirname(_ROOT) != _ROOT:
    _ROOT = os.path.dirname(_ROOT)
sys.path.insert(1, _ROOT)

from monocular_depth_estimation_trt_b200 import weights  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint")
    ap.add_argument("--seeded-init", action="store_true")
    ap.add_argument("--encoder", default="vits")
    ap.add_argument("--out-dir", default=os.path.join(CUR_DIR, "onnx"))
    a = ap.parse_args(argv)
    input_h = input_w = 518
    metric_model, dataset = True, "hypersim"
    max_depth = 20.0 if dataset == "hypersim" else 80.0        # models/depth_anything_v2/infer_metric.py:61-65
    os.makedirs(a.out_dir, exist_ok=True)
    if a.checkpoint:
        import torch
        encoder = weights.encoder_of(torch.load(a.checkpoint, map_location="cpu", weights_only=True))
        name = f"depth_anything_v2_{encoder}_{input_h}x{input_w}_metric_{dataset}.mdew"
        meta = weights.export_checkpoint(a.checkpoint, os.path.join(a.out_dir, name), input_h, input_w, max_depth if metric_model else None)
    elif a.seeded_init:
        import numpy as np
        import torch
        from oracle import dav2_torch as O, preprocess_np as P       # test tooling: stands in for the missing checkpoint
        x = torch.from_numpy(P.preprocess_stretch_imagenet(np.random.default_rng(0).integers(0, 256, (480, 640, 3), dtype=np.uint8), input_h, input_w))
        sd = O.init_state_dict(a.encoder, seed=0)
        O.calibrate_head(sd, x, a.encoder)
        name = f"depth_anything_v2_{a.encoder}_{input_h}x{input_w}_metric_{dataset}.mdew"
        meta = weights.describe(a.encoder, input_h, input_w, max_depth)
        weights.save(os.path.join(a.out_dir, name), sd, meta)
    else:
        ap.error("--checkpoint or --seeded-init")
    print(f"[MDET] export -> {os.path.join(a.out_dir, name)}  ({meta['encoder']}, {meta['input_h']}x{meta['input_w']})")
    return os.path.join(a.out_dir, name)


if __name__ == "__main__":
    main()
