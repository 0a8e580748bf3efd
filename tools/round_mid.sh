#!/bin/bash
# mid-round GPU pass after a K2 change: the K2 captures on identical launches + the checked-build runs
P=${1:-r2l}
mkdir -p gpurun_out
cap() { # name workload kernel-regex
  python tools/ncu_case.py $2 4 > gpurun_out/${P}_$1_case.json 2> gpurun_out/${P}_$1_case.err && \
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 3 -c 1 -o gpurun_out/${P}_$1 python tools/ncu_case.py $2 4 > gpurun_out/${P}_$1_ncu.log 2>&1
  cat gpurun_out/${P}_$1_case.json
}
cap k2_lean4_dense 1080p_dense_512 dda_accumulate_lean4
cap k2_lean_scatter10 1080p_scatter10_512 dda_accumulate_lean_kernel
B="python bench.py --workload 1080p_dense_512 --steps 2 --warmup 3 --no-cpu --no-also"
$B > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches_bench_1080p_dense_512.csv $B > gpurun_out/${P}_ncu_launches.log 2>&1
bash tools/checked_run.sh $P
