#!/bin/bash
# exact per-launch ncu figures (tools/ncu_case.py: identical launches): --set full of the last launch of K2 dense / K2 scattered / K3
P=${1:-r2j}
mkdir -p gpurun_out
cap() { # name workload kernel-regex
  python tools/ncu_case.py $2 4 > gpurun_out/${P}_$1_case.json 2> gpurun_out/${P}_$1_case.err && \
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 3 -c 1 -o gpurun_out/${P}_$1 python tools/ncu_case.py $2 4 > gpurun_out/${P}_$1_ncu.log 2>&1
  cat gpurun_out/${P}_$1_case.json
}
cap k2_lean4_dense 1080p_dense_512 dda_accumulate_lean4
cap k2_lean_scatter10 1080p_scatter10_512 dda_accumulate_lean_kernel
cap k3_lockstepN image_4096_512_fixed process_image_lockstepN
ls -la gpurun_out | grep ${P}_
