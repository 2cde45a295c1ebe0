#!/bin/bash
# Run the GPU test groups in separate processes (a device trap poisons the CUDA context of the
# process that hit it), each under its own timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in "$@"; do
  case "$grp" in
    engine*) file=tests/test_engine_gpu.py; key="${grp#engine}"; key="${key#:}";;
    *) file=tests/test_kernels_gpu.py; key="$grp";;
  esac
  log="gpurun_out/test_$(echo "$grp" | tr ':/ ' '___').log"
  if [ -n "$key" ]; then
    timeout 900 python -m pytest "$file" -q -m gpu -k "$key" -p no:cacheprovider > "$log" 2>&1
  else
    timeout 900 python -m pytest "$file" -q -m gpu -p no:cacheprovider > "$log" 2>&1
  fi
  echo "== $grp exit $? : $(tail -n 1 "$log")"
done
