#!/usr/bin/env python
"""Recipe for oracle/_ref: the UNMODIFIED Python reference, made able to travel to the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference (lr40/marl-scheduling) is pure Python with no
package metadata (its pyproject.toml only configures black/isort/pyright, so `pip install` has
nothing to install); the modules on the hot path of SURVEY.md section 8(a) import each other by
bare file name (src/SchedulingEnvironment.py:12-18).  This script copies exactly those files,
byte for byte, from /root/reference/src into oracle/_ref/src -- a directory that is listed in
.gitignore (never in history) but not in .gpurunignore (it ships with the snapshot like the built
.so files) -- and writes a manifest with their SHA-256.  __graft_entry__.build() runs it when
/root/reference is mounted (the build container); on the GPU box only the copied files exist.

Consumers: oracle/ref_harness.py (falls back to oracle/_ref/src when /root/reference is absent)
and therefore bench.py's `cpu_baseline` / `--impl reference` legs and the tests that re-generate
fixtures.  Nothing under marl_scheduling_b200/ reads it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src"
DST = os.path.join(HERE, "_ref", "src")
# the import closure of SchedulingEnvironment.py (SURVEY.md section 8(a) + the modules it star-imports)
FILES = ["SchedulingEnvironment.py", "world.py", "Agent.py", "Auctioneer.py", "HardcodedModules.py",
         "Reward.py", "PPOmodules.py", "DQNmodules.py", "Plot.py", "SavingAndLoading.py"]


def main():
    if not os.path.isdir(SRC):
        print("make_ref: %s not mounted, keeping whatever oracle/_ref holds" % SRC)
        return 0
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        with open(os.path.join(DST, f), "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    lic = "/root/reference/LICENSE"
    if os.path.exists(lic):
        shutil.copyfile(lic, os.path.join(HERE, "_ref", "LICENSE"))
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1)
    print("make_ref: %d reference files -> %s" % (len(FILES), DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
