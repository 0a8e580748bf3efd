"""profiles/traffic.json feeds ncu figures (warp instructions and DRAM bytes per voxel step) into bench.py's `roofline`.  They are only
valid for the SASS they were captured from: this test fails when the kernels of the built libpvp_b200.so differ from the ones the file
names (md5 of the instruction text, tools/sass_md5.py) -- re-capture ncu (tools/profile_round.sh), update the figures, then
`python tools/sass_md5.py --update`."""
import json
import os
import shutil
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.skipif(shutil.which("cuobjdump") is None and not os.path.exists("/usr/local/cuda/bin/cuobjdump"), reason="no cuobjdump")
def test_ncu_figures_belong_to_the_built_kernels():
    import sass_md5
    fresh = sass_md5.check()
    assert fresh and all(fresh.values()), f"stale ncu figures in profiles/traffic.json for {[k for k, ok in fresh.items() if not ok]}"


def test_traffic_entries_are_complete():
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for k, e in t.items():
        if not isinstance(e, dict):
            continue
        assert {"kernel", "bytes_per_launch", "voxel_steps_per_launch", "warp_inst_per_launch", "sass_function", "sass_md5", "source"} <= set(e), k
        assert os.path.exists(os.path.join(ROOT, e["source"].split(" ")[0])), e["source"]
