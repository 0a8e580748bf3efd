#!/bin/bash
# GPU pass after the K3 change: K3 parity, the path-B bench lines (fixed point and float64, 1 and default steps), the exact K3 capture
P=${1:-r2o}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_process_image.py tests/test_gpu_fullsize_parity.py tests/test_compiled_dropin.py -x -q -m gpu > gpurun_out/${P}_k3_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/${P}_k3_parity.log | cut -c1-200
for W in image_4096_512_fixed image_4096_512_f64 image_1024_256_f64; do
  for V in 0 3; do
    python bench.py --workload $W --steps 5 --warmup 3 --no-cpu --opt image_variant=$V > gpurun_out/${P}_${W}_v$V.json 2> gpurun_out/${P}_${W}_v$V.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${P}_${W}_v$V.json").read().strip().splitlines()[-1])
    print("$W variant $V: %.1f G samples/s  e2e %.1f  ms %.2f  atomics/sample %s" % (d["value"]/1e9, d["e2e"]["value"]/1e9, d["ms_per_step"], d.get("atomics_per_sample")))
except Exception as e:
    print("$W $V FAILED", e, open("gpurun_out/${P}_${W}_v$V.err").read()[-300:])
PY
  done
done
cap() { # name workload kernel-regex
  python tools/ncu_case.py $2 4 > gpurun_out/${P}_$1_case.json 2> gpurun_out/${P}_$1_case.err && \
  ncu --set full --clock-control none --import-source on -k regex:$3 -s 3 -c 1 -o gpurun_out/${P}_$1 python tools/ncu_case.py $2 4 > gpurun_out/${P}_$1_ncu.log 2>&1
  cat gpurun_out/${P}_$1_case.json
}
cap k3_lockstepN image_4096_512_fixed process_image_lockstepN
I="python bench.py --workload image_4096_512_fixed --steps 2 --warmup 3 --no-cpu"
$I > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/${P}_launches_bench_image_4096_512_fixed.csv $I > gpurun_out/${P}_ncu_l3.log 2>&1
