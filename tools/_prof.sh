set -x
mkdir -p gpurun_out
B="python bench.py --workload 1080p_dense_512 --steps 2 --warmup 1 --no-cpu"
I="python bench.py --workload image_4096_512_fixed --steps 2 --warmup 1 --no-cpu"
$B > gpurun_out/plain_k2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches_bench_1080p_dense_512.csv $B > gpurun_out/ncu_l2.log 2>&1
$B > gpurun_out/plain_k2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dda_accumulate_lean -s 1 -c 2 -o gpurun_out/r1b_k2_lean $B > gpurun_out/ncu_f2.log 2>&1
$I > gpurun_out/plain_k3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r1b_launches_bench_image_4096_512_fixed.csv $I > gpurun_out/ncu_l3.log 2>&1
$I > gpurun_out/plain_k3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:process_image_lockstep -s 1 -c 1 -o gpurun_out/r1b_k3_lockstep $I > gpurun_out/ncu_f3.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --workload image_4096_512_fixed --steps 5 --warmup 3 > gpurun_out/bench_image.json 2> gpurun_out/bench_image.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python tools/atomic_roofline.py > gpurun_out/atomic.log 2>&1
ls -la gpurun_out | tail -20
