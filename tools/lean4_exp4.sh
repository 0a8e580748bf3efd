#!/bin/bash
mkdir -p gpurun_out
P=${1:-r2i}
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
for L in "$@"; do
  run v6_$L "$L" --variant 6
done
