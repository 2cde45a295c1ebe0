#!/bin/bash
# compute-sanitizer over the kernel unit tests (small shapes): memcheck, synccheck and racecheck logs under gpurun_out/.
# Usage (on the GPU box): bash tools/sanitize.sh
set -u
SEL='gemm_bias_fp32_out or gemm_gelu or gemm_bulk_store or layerscale_residual or token_remap or pixel_shuffle or conv3x3 or layernorm or bilinear or im2col or upconv_head or resize_depth or ragged_token_counts or different_row_sets or degenerate or fused_gather or peer_signal or preprocess'
for tool in memcheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 \
    python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/r02_sanitizer_$tool.log
  tail -4 gpurun_out/r02_sanitizer_$tool.log
done
# racecheck watches shared-memory hazards; it is the slowest tool, so it gets the protocol-heavy kernels only
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 77 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "gemm_bias_fp32_out or ragged_token_counts or layernorm or upconv_head" > gpurun_out/r02_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/r02_sanitizer_racecheck.log
tail -4 gpurun_out/r02_sanitizer_racecheck.log
