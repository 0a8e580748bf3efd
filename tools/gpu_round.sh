#!/bin/bash
# One gpurun call: GPU parity tests, then K2 variant comparison on the headline workloads.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for wl in 1080p_dense_512 1080p_dense_256; do
  for v in 1 4; do
    python bench.py --workload $wl --variant $v --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_v${v}_${wl}.json 2> gpurun_out/bench_v${v}_${wl}.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/bench_v${v}_${wl}.json").read().strip().splitlines()[-1])
print("$wl v$v", round(d["value"]/1e9,1), "G steps/s  e2e", round(d["e2e"]["value"]/1e9,1), "ms", round(d["ms_per_step"],2), "atomics/step", round(d.get("atomics_per_voxel_step",0),3))
PY
  done
done
python bench.py --workload 1080p_dense_512_fixed --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_fixed.json 2> gpurun_out/bench_fixed.err; tail -c 600 gpurun_out/bench_fixed.json
