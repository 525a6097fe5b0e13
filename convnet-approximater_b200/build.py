"""Builds libcna_b200.so (sm_100a only) from csrc/*.cu with nvcc, in-tree.

    python convnet-approximater_b200/build.py [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libcna_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "--use_fast_math" if False else "-DCNA_NO_FAST_MATH", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
         "--expt-relaxed-constexpr"]


def _stamp():
    h = hashlib.sha256()
    for f in sorted(glob.glob(os.path.join(CSRC, "*")) + [os.path.join(HERE, "..", "include", "cna_b200.h"), __file__]):
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force=False, verbose=True):
    os.makedirs(LIBDIR, exist_ok=True)
    stamp_file = os.path.join(LIBDIR, ".stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    objs = []
    procs = []
    for src in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    with open(os.path.join(LIBDIR, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    cmd = [NVCC, "-arch=sm_100a", "-shared", "-o", LIB, *objs, "-lcudart"]
    subprocess.check_call(cmd)
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
