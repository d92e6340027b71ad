"""Build recipe for libcti_sm100.so: plain nvcc, sm_100a only, in-tree output.

    python -m cti_b200.build          (or: python iccv19_vqa-cti_b200/build.py)

The shared library has a C ABI (include/cti_sm100.h), links the CUDA runtime statically and
resolves cuTensorMapEncodeTiled through cudaGetDriverEntryPoint, so it needs neither libcuda at
link time nor any torch header.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcti_sm100.so")
SOURCES = ["cti_capi.cu", "gemm_tcgen05.cu", "elementwise.cu", "softmax.cu", "trilinear.cu", "trilinear_tc.cu", "trilinear_bwd_tc.cu", "pool.cu", "bilinear.cu", "bilinear_tc.cu", "optim.cu", "gru.cu", "loss.cu", "glimpse.cu", "rank_proj.cu", "peer.cu"]
HEADERS = ["cti_common.cuh", "cti_kernels.h", "wmma_tiles.cuh", "tc_tiles.cuh", os.path.join("..", "..", "include", "cti_sm100.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
if os.environ.get("CTI_PROF"):                     # debug: per-role cycle counters in the contraction kernel
    NVCC_FLAGS.append("-DCTI_PROF=1")
if os.environ.get("CTI_WATCHDOG"):                 # debug: mbarrier watchdog (see csrc/cti_common.cuh)
    NVCC_FLAGS.append("-DCTI_WATCHDOG=" + os.environ["CTI_WATCHDOG"])


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libcti_sm100.so cannot be built")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    stamp = os.path.join(objdir, "flags.txt")            # a change of flags (debug builds) rebuilds everything
    flags = " ".join(NVCC_FLAGS)
    if not os.path.exists(stamp) or open(stamp).read() != flags:
        force = True
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        r = subprocess.run([nvcc, *NVCC_FLAGS, "-c", s, "-o", o], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{r.stdout}\n{r.stderr}")
        return s, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, log in ex.map(compile_one, jobs):
            if verbose:
                print(f"== {os.path.basename(s)}\n{log}")
    with open(stamp, "w") as f:
        f.write(flags)
    objs = [os.path.join(objdir, src.replace(".cu", ".o")) for src in SOURCES]
    if force or jobs or _stale(LIB, objs):
        r = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-cudart", "static"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
