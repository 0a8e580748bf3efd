for r in peers nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --reduce $r > gpurun_out/bench_n2_$r.json 2> gpurun_out/bench_n2_$r.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n2_$r.json").read().strip().splitlines()[-1])
print("$r", round(d["value"]/1e9,1), "G steps/s", "merge ms", d["grid_merge_ms"], d["config"]["parallelism"])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 1 --no-cpu --workload image_4096_512_fixed > gpurun_out/bench_img_n2.json 2> gpurun_out/bench_img_n2.err; head -c 250 gpurun_out/bench_img_n2.json; tail -3 gpurun_out/bench_img_n2.err
