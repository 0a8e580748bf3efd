#!/bin/bash
# K2 on scattered motion: variants x queue tile heights (bench lines under gpurun_out/), then one full ncu capture of the auto pick.
mkdir -p gpurun_out
W=${1:-1080p_scatter10_512}
run() { python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-also "${@:2}" > gpurun_out/sc_$1.json 2> gpurun_out/sc_$1.err; python - <<P
import json
try:
    d=json.load(open("gpurun_out/sc_$1.json")); print("$1", "%.1f G" % (d["value"]/1e9), "atomics/step %.3f" % d["atomics_per_voxel_step"], "steps/ray %.0f" % d["voxel_steps_per_ray"], "ms %.2f" % d["ms_per_step"])
except Exception as e: print("$1", "failed", e)
P
}
run auto
run v4_rg2 --variant 4 --opt k1_row_group=2
run v4_rg4 --variant 4 --opt k1_row_group=4
run v4_rg8 --variant 4 --opt k1_row_group=8
run v5_rg2 --variant 5 --opt k1_row_group=2
run v5_rg8 --variant 5 --opt k1_row_group=8
run v1 --variant 1
run v2 --variant 2
run v3 --variant 3
B="python bench.py --workload $W --steps 2 --warmup 3 --no-cpu --no-also"
$B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dda_accumulate_lean2 -s 6 -c 1 -o gpurun_out/r2a_k2_lean2_scatter10 $B > gpurun_out/ncu_sc.log 2>&1
