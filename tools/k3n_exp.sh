#!/bin/bash
# N-pixels-per-lane K3 (image_variant 4): parity, then tools/image_probe.py for the product library and per experimental build
# (tools/build_exp.sh name "-DPVP_K3N_..." pvp_image); usage: tools/k3n_exp.sh <prefix> lib1 lib2 ...
mkdir -p gpurun_out
P=${1:-r2m}; shift
timeout 900 python -m pytest tests/test_gpu_process_image.py -x -q -m gpu > gpurun_out/${P}_k3_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/${P}_k3_parity.log | cut -c1-200
python tools/image_probe.py 3 4 2>&1 | grep -v "^$" | sed "s/^/product: /"
for L in "$@"; do
  PVP_LIB_VARIANT=$L python tools/image_probe.py 4 2>&1 | sed "s/^/$L: /"
done
