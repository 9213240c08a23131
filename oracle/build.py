"""Builds the CPU oracle (test infrastructure) into oracle/_build/liborc.so.

The reference is Rust; there is no rustc/cargo in this image, so there is no oracle/_ref:
the oracle is the C++ restatement in shimmer_oracle.cpp ("port").
"""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "liborc.so"
SRC = HERE / "shimmer_oracle.cpp"
# -O2, IEEE (no -ffast-math, no contraction): op-for-op what rustc emits for the reference
FLAGS = ["-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread"]


def build(force: bool = False) -> Path:
    if force or not LIB.exists() or LIB.stat().st_mtime < SRC.stat().st_mtime:
        LIB.parent.mkdir(parents=True, exist_ok=True)
        subprocess.run(["g++", *FLAGS, "-o", str(LIB), str(SRC)], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
