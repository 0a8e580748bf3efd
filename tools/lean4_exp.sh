#!/bin/bash
# lean4 (four rays per lane) and the run-cap / counter experiments: parity first, then one short bench line per build and variant.
mkdir -p gpurun_out
P=${1:-r2f}
timeout 900 python -m pytest tests/test_gpu_motion.py -x -q -m gpu -k "variants or sweep or slab or row or sequence" > gpurun_out/${P}_parity_product.log 2>&1; echo "parity product rc=$?"; tail -2 gpurun_out/${P}_parity_product.log
for L in cap8 cap16; do
  PVP_LIB_VARIANT=$L timeout 900 python -m pytest tests/test_gpu_motion.py -x -q -m gpu -k "variants or sweep" > gpurun_out/${P}_parity_$L.log 2>&1; echo "parity $L rc=$?"; tail -1 gpurun_out/${P}_parity_$L.log
done
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
for L in "" cap16 cap8 nocount cap8nc; do
  run v5_$L "$L" --variant 5
  run v6_$L "$L" --variant 6
  run v6rg8_$L "$L" --variant 6 --opt k1_row_group=-8
done
run v6r80_ "" --variant 6 --opt lean_regs=80
run v6_obj "" --variant 6 --workload 1080p_object_512
run v5_obj "" --variant 5 --workload 1080p_object_512
run v6_256 "" --variant 6 --workload 1080p_dense_256
run v6_4k "" --variant 6 --workload 4k_dense_512
run v6_fixed "" --variant 6 --workload 1080p_dense_512_fixed
