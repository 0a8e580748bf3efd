#!/bin/bash
mkdir -p gpurun_out
for L in "$@"; do
  PVP_LIB_VARIANT=$L python tools/image_probe.py 4 2>&1 | sed "s/^/$L: /"
done
