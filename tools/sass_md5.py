"""md5 of the SASS of one kernel of the built libpvp_b200.so (cuobjdump -sass), instruction text only.
profiles/traffic.json stores the md5 of the kernels its ncu figures (warp instructions and DRAM bytes per voxel step) were captured
from; bench.py and tests/test_profiles_fresh.py compare it with the library that is actually loaded, so a figure taken from an older
build of a kernel cannot silently flow into `roofline`.
    python tools/sass_md5.py                      # md5 of every kernel named in profiles/traffic.json
    python tools/sass_md5.py --update             # rewrite the md5 fields (after re-capturing ncu for the current build)"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pixeltovoxelprojector_b200", "libpvp_b200.so")
TRAFFIC = os.path.join(ROOT, "profiles", "traffic.json")
_cache: dict[str, dict[str, str]] = {}


def _cuobjdump() -> str | None:
    return shutil.which("cuobjdump") or ("/usr/local/cuda/bin/cuobjdump" if os.path.exists("/usr/local/cuda/bin/cuobjdump") else None)


def functions(lib: str = LIB) -> dict[str, str]:
    """{mangled kernel name: md5 of its instruction lines} for every kernel of the library."""
    if lib in _cache:
        return _cache[lib]
    exe = _cuobjdump()
    if exe is None:
        raise FileNotFoundError("cuobjdump not found")
    txt = subprocess.run([exe, "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, name, h = {}, None, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                out[name] = h.hexdigest()
            name, h = m.group(1), hashlib.md5()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", line)   # "/*0010*/  INSTR ;  /* 0x... */"
        if m and h is not None:
            h.update(m.group(1).encode() + b"\n")
    if name:
        out[name] = h.hexdigest()
    _cache[lib] = out
    return out


def md5_of(pattern: str, lib: str = LIB) -> str:
    """md5 of the one kernel whose mangled name matches the regex `pattern`."""
    hits = {k: v for k, v in functions(lib).items() if re.search(pattern, k)}
    if len(hits) != 1:
        raise KeyError(f"{pattern!r} matches {len(hits)} kernels: {sorted(hits)[:6]}")
    return next(iter(hits.values()))


def check(lib: str = LIB) -> dict[str, bool]:
    """{traffic.json entry: its sass_md5 equals the built library's}"""
    t = json.load(open(TRAFFIC))
    return {k: md5_of(e["sass_function"], lib) == e.get("sass_md5") for k, e in t.items() if isinstance(e, dict) and "sass_function" in e}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--update", action="store_true")
    ap.add_argument("--list", action="store_true")
    a = ap.parse_args()
    if a.list:
        for k, v in sorted(functions().items()):
            print(v, k)
        sys.exit(0)
    t = json.load(open(TRAFFIC))
    for k, e in t.items():
        if isinstance(e, dict) and "sass_function" in e:
            cur = md5_of(e["sass_function"])
            print(k, e["sass_function"], cur, "==" if cur == e.get("sass_md5") else "!= " + str(e.get("sass_md5")))
            if a.update:
                e["sass_md5"] = cur
    if a.update:
        json.dump(t, open(TRAFFIC, "w"), indent=1)
        open(TRAFFIC, "a").write("\n")
