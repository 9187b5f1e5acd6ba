"""In-tree build of the CUDA library (nvcc, sm_100a only).  `python -m recboard_b200.build`.

Every ``csrc/*.cu`` is compiled to an object file (in parallel: the tcgen05 kernel templates are instantiated
one epilogue / pass per translation unit) and the objects are linked into ``_C/librecboard_b200.so``; an object
is rebuilt when its source, any header or the flags changed."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
SRC = PKG / "csrc"
SUFFIX = os.environ.get("RB_SO_SUFFIX", "")   # suffix: experiment builds
OUT = PKG / "_C" / ("librecboard_b200" + SUFFIX + ".so")
OBJ = PKG / "_C" / ("obj" + SUFFIX)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _headers():
    return sorted(SRC.glob("*.cuh")) + [PKG.parent / "include" / "recboard_b200.h"]


def _flags(verbose: bool):
    return ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
            "-Xptxas", "-v" if verbose else "-O3", *os.environ.get("RB_EXTRA_NVCC_FLAGS", "").split()]


def needs_build() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    return any(p.stat().st_mtime > t for p in list(SRC.glob("*.cu")) + _headers())


def _compile(src: Path, flags, stamp: str, force: bool, verbose: bool) -> Path:
    obj, tag = OBJ / (src.stem + ".o"), OBJ / (src.stem + ".stamp")
    newest = max(p.stat().st_mtime for p in [src] + _headers())
    if not force and obj.exists() and tag.exists() and tag.read_text() == stamp and obj.stat().st_mtime >= newest:
        return obj
    r = subprocess.run([NVCC, *flags, "-c", "-o", str(obj), str(src)], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed on {src.name}")
    if verbose:
        sys.stderr.write(r.stderr)
    tag.write_text(stamp)
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    OBJ.mkdir(parents=True, exist_ok=True)
    flags = _flags(verbose)
    stamp = hashlib.sha1(" ".join([NVCC] + flags).encode()).hexdigest()
    sources = sorted(SRC.glob("*.cu"))
    with ThreadPoolExecutor(max_workers=min(len(sources), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, flags, stamp, force, verbose), sources))
    r = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(OUT), *map(str, objs)],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking librecboard_b200.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
