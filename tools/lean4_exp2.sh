#!/bin/bash
# lean4 tuning: CTA shape / register budget / run cap / unroll, one short bench line per build; then the other workloads with the product build
mkdir -p gpurun_out
P=${1:-r2g}
run() { # name lib args...
  name=$1; lib=$2; shift 2
  PVP_LIB_VARIANT=$lib python bench.py --steps 5 --warmup 3 --no-cpu --no-also "$@" > gpurun_out/${P}_$name.json 2> gpurun_out/${P}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${P}_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"]/1e9,1), "G steps/s  e2e", round(d["e2e"]["value"]/1e9,1), "ms", round(d["ms_per_step"],2), "k2ms", round(d["roofline"]["avg_launch_ms"],2), "atomics/step", round(d.get("atomics_per_voxel_step",0),3), "pairs/s", round(d["frames_per_s"]))
except Exception as e:
    print("$name FAILED", e, open("gpurun_out/${P}_$name.err").read()[-300:])
PY
}
run v5 "" --variant 5
for L in "" t192r112 t192r104 t128m5 t128m4 t128r88 l4cap4 l4cap16 l4un2 t192un2; do
  run v6_$L "$L" --variant 6
done
run v6_count "" --variant 6 --opt count_atomics=1
run v6_rg8 "" --variant 6 --opt k1_row_group=-8
for W in 1080p_object_512 1080p_dense_256 4k_dense_512 1080p_dense_512_fixed 1080p_scatter60_512 1080p_scatter30_512; do
  run v6_$W "" --variant 6 --workload $W
  run auto_$W "" --workload $W
done
run v6_1024 "" --variant 6 --workload 1080p_dense_1024
run v6_1024rg8 "" --variant 6 --workload 1080p_dense_1024 --opt k1_row_group=-8
run auto_1024 "" --workload 1080p_dense_1024
