source tools/gpu_exp.sh
run sparse --workload 1080p_sparse_512
python - <<'PY'
import json
d=json.loads(open("gpurun_out/exp_sparse.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("frames_per_s","rays_per_s","ms_per_step","gpu_launches")}, d["e2e"], d["roofline"]["k1_ms_total"], d["roofline"]["avg_launch_ms"], d["roofline"]["k2_share_of_step"])
PY
