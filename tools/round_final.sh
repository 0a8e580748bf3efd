#!/bin/bash
# The round's closing GPU run: whole -m gpu suite, the default bench line (what the driver runs), the reference arm, the checked build.
P=${1:-r2p}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${P}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${P}_bench_default.json 2> gpurun_out/${P}_bench_default.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${P}_bench_default.json").read().strip().splitlines()[-1])
print("value %.1f G  e2e %.1f G  ms/step %.2f  pick %s  atomics/step %.3f  roofline %s %.3f md5ok %s" % (d["value"]/1e9, d["e2e"]["value"]/1e9, d["ms_per_step"], d["roofline"].get("k2_pick"), d["atomics_per_voxel_step"], d["roofline"]["bound"], d["roofline"]["frac"], d["roofline"].get("sass_md5_ok")))
print("parity", {k: d["parity"].get(k) for k in ("k2_kernel","counters_equal","support_equal","bit_mismatches_below_2p24","cells_at_or_above_2p24","gpu_max_rel_err_vs_exact","reference_max_rel_err_vs_exact")})
print("also", {k: (round(v.get("value",0)/1e9,1), v.get("unit")) if isinstance(v, dict) else v for k, v in d.get("also", {}).items()})
print("cpu", d["cpu_baseline"]["value"]/1e6, "M/s", "clocks", d["clocks"])
PY
bash tools/checked_run.sh $P
ls gpurun_out | grep ${P}_ | tr '\n' ' '
