#!/bin/bash
# exact ncu captures (identical launches, tools/ncu_case.py) of the dominant kernel of the secondary workloads
P=${1:-r2q}
mkdir -p gpurun_out
cap() { # name workload kernel-regex [pairs]
  python tools/ncu_case.py $2 4 $4 > gpurun_out/${P}_$1_case.json 2> gpurun_out/${P}_$1_case.err && \
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 3 -c 1 -o gpurun_out/${P}_$1 python tools/ncu_case.py $2 4 $4 > gpurun_out/${P}_$1_ncu.log 2>&1
  cat gpurun_out/${P}_$1_case.json
}
cap k2_lean4_dense256 1080p_dense_256 dda_accumulate_lean4
cap k2_lean4_4k 4k_dense_512 dda_accumulate_lean4 8
cap k2_lean2_object 1080p_object_512 dda_accumulate_lean2
cap k2_lean_1024 1080p_dense_1024 dda_accumulate_lean_kernel
cap k3_lockstepN_f64 image_4096_512_f64 process_image_lockstepN
ls -la gpurun_out | grep ${P}_ | awk '{print $5, $9}'
