#!/bin/bash
# Profiling recipe of a round (run on the GPU box through gpurun): ncu launch lists + one --set full capture of K2 and K3 (each after the
# same command has exited 0 without ncu), then the plain benches.  Outputs land in gpurun_out/; the summaries are copied into profiles/.
# usage: bash tools/profile_round.sh <prefix>     e.g. r2a
P=${1:-r2a}
set -x
mkdir -p gpurun_out
B="python bench.py --workload 1080p_dense_512 --steps 2 --warmup 3 --no-cpu --no-also"
I="python bench.py --workload image_4096_512_fixed --steps 2 --warmup 3 --no-cpu"
$B > gpurun_out/plain_k2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches_bench_1080p_dense_512.csv $B > gpurun_out/ncu_l2.log 2>&1
$B > gpurun_out/plain_k2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dda_accumulate_lean4 -s 6 -c 2 -o gpurun_out/${P}_k2_lean4 $B > gpurun_out/ncu_f2.log 2>&1
$I > gpurun_out/plain_k3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/${P}_launches_bench_image_4096_512_fixed.csv $I > gpurun_out/ncu_l3.log 2>&1
$I > gpurun_out/plain_k3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:process_image_lockstepN -s 3 -c 1 -o gpurun_out/${P}_k3_lockstepN $I > gpurun_out/ncu_f3.log 2>&1
ls -la gpurun_out | grep ${P}
# scattered motion (10 % iid): the auto pick is the one-ray-per-lane kernel
S="python bench.py --workload 1080p_scatter10_512 --steps 2 --warmup 3 --no-cpu --no-also"
$S > gpurun_out/plain_sc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dda_accumulate_lean_kernel -s 6 -c 1 -o gpurun_out/${P}_k2_lean_scatter10 $S > gpurun_out/ncu_fs.log 2>&1
ls -la gpurun_out | grep ${P}
