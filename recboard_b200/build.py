"""In-tree build of the CUDA library (nvcc, sm_100a only).  `python -m recboard_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
SRC = PKG / "csrc"
OUT = PKG / "_C" / ("librecboard_b200" + os.environ.get("RB_SO_SUFFIX", "") + ".so")   # suffix: experiment builds
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def needs_build() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    deps = list(SRC.glob("*.cu")) + list(SRC.glob("*.cuh")) + [PKG.parent / "include" / "recboard_b200.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    OUT.parent.mkdir(parents=True, exist_ok=True)
    cmd = [
        NVCC, "-std=c++17", "-O3", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-Xcompiler", "-fPIC", "-shared",
        "-Xptxas", "-v" if verbose else "-O3",
        *os.environ.get("RB_EXTRA_NVCC_FLAGS", "").split(),
        "-o", str(OUT), str(SRC / "abi.cu"),
    ]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building librecboard_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
