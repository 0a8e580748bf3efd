#!/bin/bash
mkdir -p gpurun_out
run() { python bench.py --workload $1 --steps 3 --warmup 3 --no-cpu --no-also "${@:3}" > gpurun_out/oe_$1_$2.json 2> gpurun_out/oe_$1_$2.err; python - <<P
import json
try:
    d=json.load(open("gpurun_out/oe_$1_$2.json")); print("$1 $2", "%.1f G" % (d["value"]/1e9), "atomics/step %.3f" % d["atomics_per_voxel_step"], "ms %.2f" % d["ms_per_step"], "k1 ms %.2f" % d["roofline"]["k1_ms_total"], "fps %.0f" % d["frames_per_s"], "e2e fps %.0f" % d["e2e"]["frames_per_s"])
except Exception as e: print("$1 $2", "failed", e, open("gpurun_out/oe_$1_$2.err").read()[-300:])
P
}
for W in "$@"; do
run $W auto
run $W z8 --opt k1_row_group=-8
run $W rg4_v5 --variant 5 --opt k1_row_group=4
done
