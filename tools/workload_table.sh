#!/bin/bash
# one short bench line per secondary workload (the table of DESIGN.md 5): value / e2e / REDs per voxel step / which K2 kernel ran
mkdir -p gpurun_out
P=${1:-r2p}
for W in 1080p_dense_512 1080p_dense_256 1080p_dense_512_fixed 4k_dense_512 1080p_dense_1024 1080p_scatter60_512 1080p_scatter30_512 1080p_scatter10_512 1080p_scatter2_512 1080p_object_512 1080p_sparse_512 files_1080p_1000 files_1080p_1000_dense image_4096_512_fixed image_4096_512_f64 image_1024_256_f64; do
  python bench.py --workload $W --steps 5 --warmup 3 --no-cpu --no-also > gpurun_out/${P}_wl_$W.json 2> gpurun_out/${P}_wl_$W.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${P}_wl_$W.json").read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print("%-26s value %8.1f G %-14s e2e %8.1f G  frames/s %9.0f  e2e frames/s %9.0f  REDs/step %s  pick %s  roofline %s %.3f" % ("$W", d["value"]/1e9, d["unit"], d["e2e"]["value"]/1e9, d.get("frames_per_s",0), d["e2e"].get("frames_per_s",0), ("%.3f" % d["atomics_per_voxel_step"]) if d.get("atomics_per_voxel_step") is not None else "-", r.get("k2_pick"), r.get("bound"), r.get("frac") or 0))
except Exception as e:
    print("$W FAILED", e, open("gpurun_out/${P}_wl_$W.err").read()[-300:])
PY
done
