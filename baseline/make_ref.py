#!/usr/bin/env python
"""Copies the UNMODIFIED reference tree into the git-ignored baseline/_ref/ so that it travels to the GPU box with the gpurun
snapshot (the box has no /root/reference).  Used by the overlay tests that drive this backend through the reference's own Runner,
configs and models (tests/test_runner_overlay_gpu.py).  Nothing under baseline/_ref is product source and nothing is edited.

    python baseline/make_ref.py [/root/reference]
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def main(src="/root/reference"):
    if not os.path.isdir(os.path.join(src, "src", "approx")):
        print(f"{src}: no reference tree here; leaving {DST} as it is")
        return 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for sub in ("src", "configs", "scripts", "tests"):
        shutil.copytree(os.path.join(src, sub), os.path.join(DST, sub), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in ("LICENSE", "README.md"):
        if os.path.exists(os.path.join(src, f)):
            shutil.copy(os.path.join(src, f), DST)
    n = sum(len(fs) for _, _, fs in os.walk(DST))
    print(f"copied {n} files of the reference into {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main(*sys.argv[1:]))
