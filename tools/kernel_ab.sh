#!/bin/bash
# Kernel A/B on one GPU box: one short bench line per experimental build (tools/build_exp.sh -> libpvp_b200_<name>.so, PVP_LIB_VARIANT).
# usage: tools/kernel_ab.sh <prefix> "<bench.py args>" lib1 lib2 ...        ("" = the product library)
#   e.g. tools/kernel_ab.sh r2k "--variant 6" "" mad0 mad1 mad3
# The lean4 / K3 tuning logs (profiles/r2_lean4_tuning.log, r2_k3n_tuning.log) were produced this way; path B: tools/k3n_exp.sh.
mkdir -p gpurun_out
P=$1; ARGS=$2; shift 2
for L in "$@"; do
  name=${L:-product}
  PVP_LIB_VARIANT=$L python bench.py --steps 5 --warmup 3 --no-cpu --no-also $ARGS > gpurun_out/${P}_$name.json 2> gpurun_out/${P}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${P}_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"]/1e9,1), "G steps/s  e2e", round(d["e2e"]["value"]/1e9,1), "ms", round(d["ms_per_step"],2), "k2ms", round(d["roofline"]["avg_launch_ms"],2), "atomics/step", round(d.get("atomics_per_voxel_step") or 0,3), "pairs/s", round(d["frames_per_s"]), "pick", d["roofline"].get("k2_pick"))
except Exception as e:
    print("$name FAILED", e, open("gpurun_out/${P}_$name.err").read()[-300:])
PY
done
