"""Build libjpp_b200.so (CUDA, sm_100a) in-tree.  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
LIB = PKG / "libjpp_b200.so"
SOURCES = [PKG / "csrc" / "jpp_b200.cu"]
HEADERS = [PKG / "csrc" / "jpp_b200_kernels.h", PKG / "csrc" / "jpp_b200_rnn.h", PKG / "csrc" / "jpp_b200_host.h",
           ROOT / "include" / "jpp_b200.h"]


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and Path(c).exists():
            return c
    raise RuntimeError("nvcc not found")


def build(force: bool = False, verbose: bool = False, defines: list[str] | None = None, out: Path | None = None) -> Path:
    """`defines`/`out` build tuning variants (e.g. JPP_FUSED_MINBLOCKS) beside the product; the default build is the product."""
    newest = max(p.stat().st_mtime for p in SOURCES + HEADERS)
    if out is not None:
        return _compile(verbose, defines or [], Path(out))
    if not force and LIB.exists() and LIB.stat().st_mtime >= newest:
        return LIB
    return _compile(verbose, defines or [], LIB)


def _compile(verbose: bool, defines: list[str], lib: Path) -> Path:
    cmd = [nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v" if verbose else "-O3",
           *[f"-D{d}" for d in defines],
           f"-I{ROOT / 'include'}", f"-I{PKG / 'csrc'}", *map(str, SOURCES), "-o", str(lib), "-lcudart", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout[-3000:] + r.stderr[-6000:])
    if verbose:
        print(r.stderr)
    return lib


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
