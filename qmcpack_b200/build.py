"""Builds libqmcb.so in-tree with nvcc for sm_100a (the only target).  `python -m qmcpack_b200.build`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libqmcb.so")
SOURCES = ["spline.cu", "crowd.cu", "api.cu", "vmc_host.cpp", "dmc_host.cpp"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".hpp"))) + [
    os.path.join("..", "..", "include", "qmcb.h"), os.path.join("..", "..", "include", "qmcb_driver.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-ccbin", "/usr/bin/g++"]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    flags = [f for f in FLAGS if f != "--use_fast_math=false"]
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        cmd = [NVCC] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on " + src)
        if verbose:
            print(out)
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-lcublas", "-ccbin", "/usr/bin/g++",
                                                   "-Xlinker", "-rpath", "-Xlinker", "/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
