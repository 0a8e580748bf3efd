#!/bin/bash
mkdir -p gpurun_out
P=${1:-r2h}
PVP_LIB_VARIANT=n3t256m3 timeout 600 python -m pytest tests/test_gpu_motion.py -x -q -m gpu -k "variants or sweep or slab or row" > gpurun_out/${P}_parity_n3.log 2>&1; echo "parity n3 rc=$?"; tail -1 gpurun_out/${P}_parity_n3.log
PVP_LIB_VARIANT=n2t256m4 timeout 600 python -m pytest tests/test_gpu_motion.py -x -q -m gpu -k "variants or sweep or slab or row" > gpurun_out/${P}_parity_n2.log 2>&1; echo "parity n2 rc=$?"; tail -1 gpurun_out/${P}_parity_n2.log
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
for L in n4t128m5 n4t64m10 n4t128m5u2 n4t128m5c16 n3t256m3 n3t128m6 n3t128m5 n2t256m3 n2t256m4 n2t128m7; do
  run v6_$L "$L" --variant 6
done
for L in n4t128m5 n3t256m3 n2t256m4; do
  run v6c_$L "$L" --variant 6 --opt count_atomics=1
  for W in 1080p_object_512 1080p_scatter30_512 1080p_dense_512_fixed 1080p_dense_1024; do
    run v6_${L}_$W "$L" --variant 6 --workload $W
  done
done
