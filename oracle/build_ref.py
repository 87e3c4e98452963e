#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY (oracle/): compile the UNMODIFIED ku-nlp/jumanpp
reference sources where they lie under /root/reference into oracle/_ref/.

Nothing is copied from the reference tree: every translation unit is compiled
in place with g++ (this script replaces the reference's CMake build, which
cannot configure offline -- see DESIGN.md "Oracle").  What this script does
that CMake would have done:

  * writes the three configured headers (util_config.h, core_config.h,
    core_version.h) from the values in the reference's CMakeLists.txt:37-51
    (JPP_MAX_DIC_FIELDS=16, JPP_PREFETCH_FEATURE_WEIGHTS on,
    JPP_TRAIN_VIOLATION_INVALID on, protobuf off);
  * puts oracle/eigen_shim on the include path in place of the Eigen download
    (reference libs/CMakeLists.txt:7-36);
  * runs the reference's own static-feature code generator
    (src/core/codegen/codegen_main.cc.in + src/jumandic/shared/jumandic_spec.cc,
    reference src/CMakeLists.txt:34-67, src/jumandic/CMakeLists.txt:19-26) to
    produce gen/jpp_jumandic_cg.{h,cc}, i.e. the "code-generated
    NgramFeatureApply" the CLI uses, and compiles it.

Outputs (git-ignored, shipped to the GPU box by gpurun):
  oracle/_ref/libjpp_ref.a        all reference objects
  oracle/_ref/jpp_ref_harness     oracle/ref_harness.cc linked against it
  oracle/_ref/jumanpp_v2          the stock reference CLI
  oracle/_ref/jpp_jumandic_bootstrap

Compiler flags: Release `-O3 -DNDEBUG` (reference README.md:119-125), `-march=x86-64-v3`
(AVX2 + FMA; not -march=native: the binaries must also run on the GPU box's
host CPU) and `-ffp-contract=off` so float sums are the plain IEEE sequence the
sources spell (the only fused operations are the explicit fmaf of the Eigen shim).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path(os.environ.get("JPP_REFERENCE_DIR", "/root/reference"))
OUT = HERE / "_ref"
CXX = os.environ.get("CXX", "g++")
CXXFLAGS = [
    "-std=c++14", "-O3", "-DNDEBUG", "-march=x86-64-v3", "-ffp-contract=off",
    "-fPIC", "-w", "-pthread",
]


def includes() -> list[str]:
    return [
        f"-I{HERE / 'eigen_shim'}",
        f"-I{OUT / 'cfg'}",
        f"-I{OUT / 'gen'}",
        f"-I{REF / 'src'}",
        f"-I{REF / 'libs'}",
        f"-I{REF / 'src' / 'jumandic'}",
        f"-I{REF / 'libs' / 'pathie-cpp' / 'include'}",
        "-D_PATHIE_UNIX=1", "-DPATHIE_VERSION_MAJOR=0", "-DPATHIE_VERSION_MINOR=1",
        "-DPATHIE_VERSION_PATCH=0", '-DPATHIE_VERSION_GIT="jumanpp"',
    ]


def write_if_changed(path: Path, text: str) -> None:
    path.parent.mkdir(parents=True, exist_ok=True)
    if path.exists() and path.read_text() == text:
        return
    path.write_text(text)


def configured_headers() -> None:
    write_if_changed(OUT / "cfg" / "util_config.h",
                     "#ifndef __JPP_UTIL_CONFIG_H__\n#define __JPP_UTIL_CONFIG_H__ 1\n#endif\n")
    write_if_changed(OUT / "cfg" / "core_config.h", """#pragma once
#ifndef __JUMANPP_CORE_CONFIG_H__
#define __JUMANPP_CORE_CONFIG_H__
#include <cstddef>
namespace jumanpp {
namespace core {
constexpr size_t JPP_MAX_DIC_FIELDS = 16;
#define JPP_PREFETCH_FEATURE_WEIGHTS 1
#define JPP_TRAIN_VIOLATION_INVALID 1
static constexpr char JPP_DEFAULT_CONFIG_DIR[]{"/usr/local/libexec/jumanpp/"};
}
}
#endif
""")
    write_if_changed(OUT / "cfg" / "core_version.h", """#pragma once
#ifndef __JUMANPP_CORE_VERSION_H__
#define __JUMANPP_CORE_VERSION_H__
#include <util/string_piece.h>
namespace jumanpp {
namespace core {
static const StringPiece JPP_VERSION_STRING { "2.0.0-dev-oracle" };
}
}
#endif
""")


def ref_sources() -> list[Path]:
    src = REF / "src"
    dirs = ["util", "core", "core/analysis", "core/dic", "core/impl",
            "core/input", "core/spec", "core/codegen", "core/training", "rnn"]
    out: list[Path] = []
    for d in dirs:
        for f in sorted((src / d).glob("*.c*")):
            n = f.name
            if n.endswith("_test.cc") or n.endswith("_test.cpp"):
                continue
            if n in ("win32_utils.cc",):
                continue
            if f.suffix not in (".cc", ".cpp"):
                continue
            out.append(f)
    for n in ["errors.cpp", "entry_iterator.cpp", "path.cpp", "temp.cpp", "pathie.cpp"]:
        out.append(REF / "libs" / "pathie-cpp" / "src" / n)
    jd = src / "jumandic" / "shared"
    for n in ["jumandic_spec.cc", "jumandic_spec_lexdata.cc", "juman_format.cc",
              "jumandic_env.cc", "morph_format.cc", "jumandic_ids.cc",
              "jumandic_id_resolver.cc", "mdic_format.cc", "subset_format.cc",
              "lattice_format.cc", "jumanpp_args.cc"]:
        out.append(jd / n)
    return out


def obj_for(src: Path) -> Path:
    try:
        rel = src.relative_to(REF)
    except ValueError:
        rel = Path("local") / src.name
    return OUT / "obj" / (str(rel).replace("/", "__") + ".o")


def needs(src: Path, obj: Path) -> bool:
    return (not obj.exists()) or obj.stat().st_mtime < src.stat().st_mtime


def compile_one(src: Path, extra: list[str] | None = None) -> Path:
    obj = obj_for(src)
    if needs(src, obj):
        obj.parent.mkdir(parents=True, exist_ok=True)
        cmd = [CXX, *CXXFLAGS, *includes(), *(extra or []), "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"compile failed: {src}\n{r.stderr[-4000:]}")
    return obj


def link(out: Path, objs: list[Path], libs: list[Path]) -> None:
    cmd = [CXX, "-pthread", "-o", str(out), *map(str, objs), *map(str, libs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed: {out}\n{r.stderr[-4000:]}")


def archive(out: Path, objs: list[Path]) -> None:
    if out.exists():
        out.unlink()
    subprocess.run(["ar", "rcs", str(out), *map(str, objs)], check=True)


def build(verbose: bool = True) -> None:
    if not REF.exists():
        raise RuntimeError(f"reference tree {REF} not present; oracle/_ref can only be built where it is mounted")
    jobs = int(os.environ.get("JPP_BUILD_JOBS", os.cpu_count() or 4))
    configured_headers()
    (OUT / "gen").mkdir(parents=True, exist_ok=True)
    all_srcs = ref_sources()
    # jumandic_env.cc includes the generated jpp_jumandic_cg.h: second phase
    late = [s for s in all_srcs if s.name == "jumandic_env.cc"]
    srcs = [s for s in all_srcs if s.name != "jumandic_env.cc"]
    with cf.ThreadPoolExecutor(jobs) as ex:
        objs = list(ex.map(compile_one, srcs))
    if verbose:
        print(f"[build_ref] {len(objs)} reference objects", flush=True)

    # --- static feature codegen (reference src/CMakeLists.txt:34-67) ---
    gen_cc = OUT / "gen" / "jpp_jumandic_cg.cc"
    gen_h = OUT / "gen" / "jpp_jumandic_cg.h"
    cg_main = OUT / "gen" / "jpp_jumandic_cg_codegen_main.cc"
    tmpl = (REF / "src/core/codegen/codegen_main.cc.in").read_text()
    tmpl = tmpl.replace("${CG_SPEC_INCLUDE}", "shared/jumandic_spec.h")
    tmpl = tmpl.replace("${CG_SPEC_NAME}", "jumanpp::jumandic::SpecFactory")
    write_if_changed(cg_main, tmpl)
    pre_lib = OUT / "libjpp_ref_nocg.a"
    archive(pre_lib, objs)
    if not (gen_cc.exists() and gen_h.exists()):
        cg_obj = compile_one(cg_main)
        cg_bin = OUT / "jpp_jumandic_codegen_binary"
        link(cg_bin, [cg_obj], [pre_lib])
        subprocess.run([str(cg_bin), "jpp_jumandic_cg", "JumandicStatic", str(OUT / "gen")], check=True)
    cg_o = compile_one(gen_cc)
    late_objs = [compile_one(s) for s in late]
    lib = OUT / "libjpp_ref.a"
    archive(lib, objs + late_objs + [cg_o])
    if verbose:
        print("[build_ref] static features generated + compiled", flush=True)

    # --- stock CLI + bootstrap + our harness ---
    mains = {
        "jumanpp_v2": REF / "src/jumandic/main/jumanpp.cc",
        "jpp_jumandic_bootstrap": REF / "src/jumandic/main/bootstrap.cc",
    }
    for name, src in mains.items():
        o = compile_one(src)
        link(OUT / name, [o], [lib])
    harness = HERE / "ref_harness.cc"
    if harness.exists():
        o = compile_one(harness, extra=[f"-I{HERE.parent / 'include'}"])
        link(OUT / "jpp_ref_harness", [o], [lib])
    # independent RNN check: the reference's vendored Eigen-free legacy implementation
    # (src/rnn/legacy/rnnlmlib_static.cpp, compiled in place) beside MikolovRnn + eigen_shim
    legacy_check = HERE / "rnn_legacy_check.cc"
    if legacy_check.exists():
        o1 = compile_one(legacy_check)
        o2 = compile_one(REF / "src/rnn/legacy/rnnlmlib_static.cpp")
        link(OUT / "rnn_legacy_check", [o1, o2], [lib])
    if verbose:
        print("[build_ref] done ->", OUT, flush=True)


if __name__ == "__main__":
    try:
        build()
    except RuntimeError as e:
        print(e, file=sys.stderr)
        sys.exit(1)
