"""Builds libmitgcm_b200.so in-tree with nvcc for sm_100a (no JIT cache: the built
library travels to the GPU box with the repo snapshot)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.environ.get("MITGCM_B200_BUILD_OUT") or os.path.join(LIBDIR, "libmitgcm_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CCBIN = ["-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"] + CCBIN
LINK = ["--shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"] + CCBIN
OBJDIR = os.path.join(HERE, "..", "build", "obj")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "mitgcm_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    extra = os.environ.get("MITGCM_B200_EXTRA_FLAGS", "").split()      # tuning variants (-DNAME=value)
    stamp = os.path.join(OBJDIR, ".flags")
    flagtxt = " ".join(FLAGS + extra)
    if not os.path.exists(stamp) or open(stamp).read() != flagtxt:
        force = True
    hdr_t = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if not f.endswith(".cu"))
    hdr_t = max(hdr_t, os.path.getmtime(os.path.join(HERE, "..", "include", "mitgcm_b200.h")))
    # one nvcc per translation unit, all at once (no relocatable device code: no device symbol crosses files)
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
            print(" ".join(cmd), file=sys.stderr)
            jobs.append((cmd, subprocess.Popen(cmd)))
    bad = [cmd for cmd, p in jobs if p.wait() != 0]
    if bad:
        raise subprocess.CalledProcessError(1, bad[0])
    cmd = [NVCC] + LINK + ["-o", LIB] + objs
    print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    open(stamp, "w").write(flagtxt)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
