#!/bin/bash
# kernel experiments: bench.py with context options; prints one summary line per configuration
mkdir -p gpurun_out
run() { # name, args...
  name=$1; shift
  python bench.py --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/exp_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["value"]/1e9,1), "G steps/s  e2e", round(d["e2e"]["value"]/1e9,1), "ms", round(d["ms_per_step"],2), "k2ms", round(d["roofline"]["avg_launch_ms"],2), "atomics/step", round(d.get("atomics_per_voxel_step",0),3), "k1ms_total", round(d["roofline"]["k1_ms_total"],3), "pairs/s", round(d["frames_per_s"]))
except Exception as e:
    print("$name FAILED", e)
PY
}
