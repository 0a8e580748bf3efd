#!/bin/bash
# The round's closing GPU run: whole -m gpu suite, the default bench line (what the driver runs), the exact ncu captures and the launch list.
P=${1:-r2j}
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${P}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${P}_bench_default.json 2> gpurun_out/${P}_bench_default.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${P}_bench_default.json").read().strip().splitlines()[-1])
print("value %.1f G  e2e %.1f G  ms/step %.2f  pick %s  atomics/step %.3f  roofline %s %.3f md5ok %s" % (d["value"]/1e9, d["e2e"]["value"]/1e9, d["ms_per_step"], d["roofline"].get("k2_pick"), d["atomics_per_voxel_step"], d["roofline"]["bound"], d["roofline"]["frac"], d["roofline"].get("sass_md5_ok")))
print("parity", {k: d["parity"].get(k) for k in ("k2_kernel","counters_equal","support_equal","bit_mismatches_below_2p24","cells_at_or_above_2p24","gpu_max_rel_err_vs_exact","reference_max_rel_err_vs_exact")})
print("also", {k: round(v.get("value",0)/1e9,1) if isinstance(v, dict) else v for k, v in d.get("also", {}).items()})
PY
bash tools/ncu_exact.sh $P > gpurun_out/${P}_ncu_exact.log 2>&1; tail -4 gpurun_out/${P}_ncu_exact.log
B="python bench.py --workload 1080p_dense_512 --steps 2 --warmup 3 --no-cpu --no-also"
$B > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches_bench_1080p_dense_512.csv $B > gpurun_out/${P}_ncu_launches.log 2>&1
ls -la gpurun_out | grep ${P}_ | awk '{print $5, $9}'
