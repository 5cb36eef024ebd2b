"""Build librtmodt_b200.so in-tree with nvcc for sm_100a (no torch involved).

    python -m rtmodt_b200.build            # or: python <package dir>/build.py

The .so lands next to this file so that it travels to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librtmodt_b200.so")
SOURCES = ["api.cu", "letterbox.cu", "nms.cu", "post.cu", "track.cu", "zone.cu"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or add /usr/local/cuda/bin to PATH)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "rtmodt_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str = LIB, defines=()) -> str:
    """Compile every CUDA source for sm_100a into one shared library; returns its path."""
    if out == LIB and not force and not is_stale():
        return LIB
    cmd = [nvcc_path(), "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-O3", "-lineinfo",
           "-gencode", "arch=compute_100a,code=sm_100a",
           # float decisions must reproduce NumPy / torch CPU bit for bit: no FMA contraction,
           # IEEE division and square root, no flush-to-zero (hard parts)
           "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", out] + [f"-D{d}" for d in defines]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
