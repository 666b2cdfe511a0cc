"""In-tree build of the sm_100a CUDA library behind include/lrx.h.

`nvcc` cross-compiles without a GPU, so this runs both in the CPU container
(the driver's "does it build" check) and on the GPU box.  The shared library
is written next to the sources (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = CSRC / "liblrx.so"
SOURCES = ["api.cu", "dense.cu", "bm25.cu", "fuse.cu", "tc_gemm.cu", "encoder.cu", "dense_batched.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def sources():
    return [CSRC / s for s in SOURCES if (CSRC / s).exists()]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    deps.append(HERE.parent / "include" / "lrx.h")
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """LRX_EXTRA_NVCC (extra flags, e.g. -DNAME=value) and LRX_ONLY (comma-separated source
    names to recompile, the other objects being reused) serve kernel tuning runs."""
    extra = os.environ.get("LRX_EXTRA_NVCC", "").split()
    only = [s for s in os.environ.get("LRX_ONLY", "").split(",") if s]
    if not force and not needs_build() and not extra and not only:
        return LIB
    nvcc = _nvcc()
    objs = []
    log = []
    for src in sources():
        obj = src.with_suffix(".o")
        if only and src.name not in only and obj.exists():
            objs.append(str(obj))
            continue
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + log[-1])
        objs.append(str(obj))
    cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + log[-1])
    (CSRC / "build.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose=True)
    print("built", p)
