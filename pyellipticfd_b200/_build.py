"""In-tree build of ``libpde_b200.so`` (sm_100a only) with nvcc.

``python -m pyellipticfd_b200._build`` or ``__graft_entry__.build()``.  The
library is written next to this file so that it travels with the repository
snapshot to the GPU box (``*.so`` is git-ignored, not gpurun-ignored).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpde_b200.so")
SOURCES = ["grid_ctx.cu", "ddg_kernels.cu", "bicgstab.cu", "misc.cu", "ell.cu", "band_builder.cu",
           "cloud_builder.cu", "first_deriv.cu", "ce_sweep.cu", "halo_push.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",                 # parity first: FMA contraction only where written explicitly
    "-Xcompiler", "-fPIC,-O2,-fopenmp,-ffp-contract=off",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "pde_b200.h"))
    objs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [path] + headers):
            cmd = [_nvcc()] + NVCC_FLAGS + ["-c", path, "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            log = os.path.join(HERE, "build", src + ".ptxas.log")
            with open(log, "w") as f:
                f.write(res.stderr)
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError("nvcc failed on %s" % src)
            if verbose:
                sys.stderr.write(res.stderr)
    if force or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-Xcompiler", "-fopenmp", "-lgomp"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
