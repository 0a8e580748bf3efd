#!/bin/bash
# The project's stand-in for compute-sanitizer (closed on this GPU pool, profiles/r2_compute_sanitizer_closed.log): run every kernel and
# variant (tools/sanitize_case.py), the exotic-geometry sweep and the whole GPU suite against libpvp_b200_checked.so, the build whose
# kernels test every queue slot, staging slot, queue entry and grid / texture cell index they compute (PVP_ERR_CHECK on a violation).
# usage (GPU box): bash tools/checked_run.sh r2      -> gpurun_out/r2_checked_*.log
P=${1:-r2}
mkdir -p gpurun_out
export PVP_LIB_VARIANT=checked
python - <<'PY' > gpurun_out/${P}_checked_build.log 2>&1
import pixeltovoxelprojector_b200 as pvp
from pixeltovoxelprojector_b200 import _lib
lib = _lib.load()
print("library:", lib._name, "build flags:", lib.pvp_build_flags())
assert lib.pvp_build_flags() & 1, "not the checked build"
PY
cat gpurun_out/${P}_checked_build.log
(python tools/sanitize_case.py; echo "sanitize_case rc=$?") > gpurun_out/${P}_checked_sanitize_case.log 2>&1
(python tests/exotic_pose_check.py --cases 400; echo "exotic rc=$?") > gpurun_out/${P}_checked_exotic.log 2>&1
(python -m pytest tests -m gpu -q -x --deselect tests/test_bench_contract.py; echo "pytest rc=$?") > gpurun_out/${P}_checked_pytest_gpu.log 2>&1
tail -n 2 gpurun_out/${P}_checked_sanitize_case.log gpurun_out/${P}_checked_exotic.log gpurun_out/${P}_checked_pytest_gpu.log
