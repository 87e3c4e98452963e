#!/usr/bin/env python3
"""Build binding/_build/jumanpp_v2_gpu: the reference CLI with scoring on the B200.

The binding is host C++ against the reference's headers (compiled where /root/reference is
mounted; its objects come from oracle/_ref/libjpp_ref.a, the unmodified reference built by
oracle/build_ref.py) and against include/jpp_b200.h.  It links libjpp_b200.so and nothing else of
this repository; nothing under oracle/*.c / *.cc is compiled into it.  The output directory is
git-ignored and travels to the GPU box like the other built binaries."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
sys.path.insert(0, str(ROOT / "oracle"))
import build_ref  # noqa: E402  (include paths / flags of the reference build)

OUT = HERE / "_build"
WRAP = "_ZN7jumanpp4core8analysis12AnalyzerImpl13computeScoresEPKNS1_9ScorerDefE"


def build(verbose: bool = True) -> Path:
    ref_lib = build_ref.OUT / "libjpp_ref.a"
    if not build_ref.REF.exists() or not ref_lib.exists():
        raise RuntimeError("the binding is compiled against the reference: needs /root/reference and oracle/_ref/libjpp_ref.a")
    OUT.mkdir(exist_ok=True)
    objs = []
    for src in (HERE / "jpp_gpu_batch.cc", HERE / "jumanpp_gpu_main.cc"):
        obj = OUT / (src.stem + ".o")
        newest = max(p.stat().st_mtime for p in [src, HERE / "jpp_gpu_batch.h", ROOT / "include" / "jpp_b200.h"])
        if not obj.exists() or obj.stat().st_mtime < newest:
            cmd = [build_ref.CXX, *build_ref.CXXFLAGS, *build_ref.includes(), f"-I{ROOT / 'include'}", f"-I{HERE}",
                   "-c", str(src), "-o", str(obj)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"compile failed: {src}\n{r.stderr[-6000:]}")
        objs.append(obj)
    exe = OUT / "jumanpp_v2_gpu"
    cmd = [build_ref.CXX, "-pthread", "-o", str(exe), *map(str, objs), str(ref_lib), f"-Wl,--wrap={WRAP}",
           f"-L{ROOT / 'jumanpp_b200'}", "-ljpp_b200", "-Wl,-rpath,$ORIGIN/../../jumanpp_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-6000:]}")
    if verbose:
        print("[build_binding]", exe)
    return exe


if __name__ == "__main__":
    try:
        build()
    except RuntimeError as e:
        print(e, file=sys.stderr)
        sys.exit(1)
