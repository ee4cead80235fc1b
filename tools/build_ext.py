#!/usr/bin/env python
"""Build libq100.so (and oracle helpers) in-tree with plain nvcc for sm_100a.

    python tools/build_ext.py [--force] [--verbose]

Every .cu under llama3-quantization_b200/csrc is compiled with
``-gencode arch=compute_100a,code=sm_100a -lineinfo`` and linked into
``llama3-quantization_b200/q100/libq100.so`` (git-ignored, travels with gpurun).
No dependency on libcuda at link time: the driver entry points the TMA path needs
are resolved through cudaGetDriverEntryPoint at run time.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "llama3-quantization_b200")
CSRC = os.path.join(PKG, "csrc")
# experiment builds: Q100_VARIANT=<tag> (+ Q100_NVCC_FLAGS) writes libq100_<tag>.so from its own object dir;
# select it at run time with Q100_LIB=<path> (q100/_lib.py)
_VAR = os.environ.get("Q100_VARIANT", "")
OUT = os.path.join(PKG, "q100", f"libq100_{_VAR}.so" if _VAR else "libq100.so")
OBJ = os.path.join(ROOT, "build", f"obj_{_VAR}" if _VAR else "obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = os.environ.get("Q100_NVCC_FLAGS", "").split() + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include")]
# per-file extras: the quantize kernels must not contract mul+add (bit-exact per-op rounding)
EXTRA = {"q100_quant.cu": ["-fmad=false"], "q100_quant_bwd.cu": ["-fmad=false"]}


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libq100.so")
    return exe


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "q100.h"))
    hdrs.append(os.path.abspath(__file__))
    # the object directory does not travel to the GPU box (.gpurunignore): a library newer than every source is current
    if not force and not _stale(OUT, [os.path.join(CSRC, s) for s in srcs] + hdrs):
        return OUT
    jobs = []
    objs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc(), *ARCH, *COMMON, *EXTRA.get(s, []), "-Xptxas", "-v", "-c", src, "-o", obj]
            jobs.append((s, cmd))

    def run(job):
        name, cmd = job
        t0 = time.time()
        p = subprocess.run(cmd, capture_output=True, text=True)
        return name, p.returncode, p.stdout + p.stderr, time.time() - t0

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for name, rc, log, dt in ex.map(run, jobs):
                with open(os.path.join(OBJ, name + ".log"), "w") as f:
                    f.write(log)
                if rc != 0:
                    sys.stderr.write(log)
                    raise RuntimeError(f"nvcc failed on {name}")
                if verbose:
                    print(f"[build] {name}: {dt:.1f}s")
    if force or jobs or _stale(OUT, objs):
        cmd = [nvcc(), *ARCH, "-shared", "-o", OUT, *objs, "-cudart", "static"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            sys.stderr.write(p.stdout + p.stderr)
            raise RuntimeError("link of libq100.so failed")
        if verbose:
            print(f"[build] linked {OUT}")
    return OUT


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose or True))
