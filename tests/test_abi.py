"""C ABI checks that need no GPU: the library loads, exports every symbol
include/jpp_b200.h declares, the ctypes mirror matches the C struct layout, and
scoring fails LOUDLY (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from jumanpp_b200 import abi

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "jpp_b200.h"


@pytest.fixture(scope="module")
def lib_path():
    from jumanpp_b200 import build
    return build.build()


def test_every_declared_symbol_is_exported(lib_path):
    text = HEADER.read_text()
    declared = set(re.findall(r"\b(jpp_b200_[a-z_]+)\s*\(", text))
    declared -= {"jpp_b200_status"}
    assert {"jpp_b200_create", "jpp_b200_score_batch", "jpp_b200_destroy", "jpp_b200_last_error"} <= declared
    lib = C.CDLL(str(lib_path))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/jpp_b200.h but not exported"
    from jumanpp_b200.scorer import EXPORTED_SYMBOLS
    assert set(EXPORTED_SYMBOLS) == declared


def test_ctypes_layout_matches_header(tmp_path):
    structs = {"jpp_b200_primitive": abi.Primitive, "jpp_b200_computation": abi.Computation, "jpp_b200_pattern": abi.Pattern,
               "jpp_b200_ngram": abi.Ngram, "jpp_b200_rnn_model": abi.RnnModel, "jpp_b200_model": abi.Model,
               "jpp_b200_config": abi.Config, "jpp_b200_batch": abi.Batch, "jpp_b200_result": abi.Result,
               "jpp_b200_counters": abi.Counters}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "jpp_b200.h"', "int main(void){"]
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append('printf("jpp_b200_beam_element %zu\\n", sizeof(jpp_b200_beam_element));')
    lines.append('printf("jpp_b200_node_info %zu\\n", sizeof(jpp_b200_node_info));')
    lines.append('printf("jpp_b200_node_ptr %zu\\n", sizeof(jpp_b200_node_ptr));')
    lines.append('printf("jpp_b200_codepoint %zu\\n", sizeof(jpp_b200_codepoint));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct in structs.items():
        assert int(out[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"
    assert int(out["jpp_b200_beam_element"]) == abi.BEAM_EL_DT.itemsize == 8
    assert int(out["jpp_b200_node_info"]) == abi.NODE_INFO_DT.itemsize == 8
    assert int(out["jpp_b200_node_ptr"]) == abi.NODE_PTR_DT.itemsize == 4
    assert int(out["jpp_b200_codepoint"]) == abi.CODEPOINT_DT.itemsize == 8


def test_no_cpu_fallback(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from conftest import load_golden
    from jumanpp_b200.scorer import LatticeScorer, ScorerError
    d, model, cfg, batch = load_golden("default")
    with pytest.raises(ScorerError) as e:
        LatticeScorer(model, cfg)
    assert e.value.code == abi.CUDA_ERROR and "no CPU path" in e.value.message


def test_product_does_not_reference_the_oracle():
    """Nothing under jumanpp_b200/ may import, load or link anything under oracle/."""
    for p in (ROOT / "jumanpp_b200").rglob("*"):
        if p.suffix in (".py", ".h", ".cu", ".cpp") and p.is_file():
            t = p.read_text(errors="ignore")
            for pat in (r'#include\s+"[^"]*oracle', r"^\s*(import|from)\s+\S*oracle", r"libjpp_oracle", r"oracle_lib",
                        r"libjpp_b200_emul", r"jpp_oracle_[a-z_]+\("):
                assert not re.search(pat, t, re.M), (p, pat)
    out = subprocess.run(["ldd", str(ROOT / "jumanpp_b200" / "libjpp_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
