"""In-tree builds: the CUDA library (nvcc, sm_100a) — nothing else belongs to the product."""
from __future__ import annotations

import os
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libshimmer_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CUDA_SOURCES = ["shim_api.cu", "shim_builder.cpp", "shim_scene.cpp"]
HEADERS = ["shim_types.h", "shim_device.h", "shim_kernels.cuh", "shim_scene.h", "shim_internal.h"]
# -fmad=false / -ffp-contract=off: primitive tests and shading keep the reference's IEEE
# operation order (see csrc/shim_device.h); sm_100a only.
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden,-pthread", "-shared"]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build_cuda(force: bool = False, verbose: bool = False, out: Path | None = None, defines=()) -> Path:
    """``out`` / ``defines``: experiment builds (A/B runs under gpurun select one with SHIMMER_B200_LIB)."""
    srcs = [CSRC / s for s in CUDA_SOURCES]
    deps = srcs + [CSRC / h for h in HEADERS] + [PKG.parent / "include" / "shimmer_b200.h"]
    lib = Path(out) if out else LIB
    if force or _stale(lib, deps):
        lib.parent.mkdir(parents=True, exist_ok=True)
        cmd = [NVCC, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", str(lib), *map(str, srcs)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, check=True)
    return lib
