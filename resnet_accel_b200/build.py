"""In-tree build of libaccel_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m resnet_accel_b200.build [--force]

The .so lands next to the sources (resnet_accel_b200/libaccel_b200.so): it is git-ignored but
travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libaccel_b200.so")
SOURCES = ["api.cu", "plan.cpp"]
# every file under csrc/ (a header left out of this list once made edits to it silently not rebuild) + the public header
DEPS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".cpp", ".h"))) + [
    os.path.join("..", "..", "include", "accel_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libaccel_b200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False, out: str | None = None, extra: list | None = None) -> str:
    """out / extra: developer builds (another output path, extra nvcc flags such as -DACCEL_DEV=1)."""
    if out is None and not force and not needs_build():
        return LIB
    out = out or LIB
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "--fmad=false",
           "-Xptxas", "-v" if verbose else "-O3",
           "-o", out] + list(extra or []) + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    _out = sys.argv[sys.argv.index("-o") + 1] if "-o" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=_out, extra=[a for a in sys.argv[1:] if a.startswith("-D")]))
