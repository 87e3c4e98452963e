#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED
reference (oracle/_ref/jpp_ref_harness, built by oracle/build_ref.py from
/root/reference).  Run in the build container only; the .jppb.gz files are
committed so that the tests need neither /root/reference nor the harness.

Each file holds a staged batch (the pre-state of computeScores), the compiled
Jumandic feature program, and the reference's post-state: stored pattern
hashes, every beam element, per-connection scores.
"""
from __future__ import annotations

import gzip
import re
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
from jumanpp_b200 import synth  # noqa: E402

HARNESS = ROOT / "oracle" / "_ref" / "jpp_ref_harness"
REF = Path("/root/reference")

RNN_PARAMS = dict(layer_size=32, maxent_order=3, maxent_size=1 << 16, nce_lnz=9.0, seed=2, vocab=5000)
RNN2_PARAMS = dict(layer_size=64, maxent_order=4, maxent_size=1 << 15, nce_lnz=7.5, seed=5, vocab=3000)
RNN3_PARAMS = dict(layer_size=128, maxent_order=3, maxent_size=1 << 16, nce_lnz=9.0, seed=11, vocab=4000)  # bench size
RNN5_PARAMS = dict(layer_size=256, maxent_order=3, maxent_size=1 << 15, nce_lnz=9.0, seed=17, vocab=2000)  # output split over 2 blocks
RNN4_PARAMS = dict(layer_size=48, maxent_order=2, maxent_size=1 << 14, nce_lnz=8.0, seed=13, vocab=2000)   # not a multiple of 32

CONFIGS = {
    # name: (harness args, sentence count)
    "default": ("beam=5 gbeam=6 rcheck=1 rbeam=5 log2w=20 wseed=1", 20),
    "beam20": ("beam=20 gbeam=20 rcheck=1 rbeam=5 log2w=18 wseed=3", 12),
    "beam20_all": ("beam=20 gbeam=20 rcheck=20 rbeam=20 log2w=18 wseed=3", 10),
    "gbeam20_partition": ("beam=5 gbeam=20 rcheck=2 rbeam=3 log2w=16 wseed=7", 14),
    "beam3_ties": ("beam=3 gbeam=4 rcheck=2 rbeam=3 log2w=12 wseed=9", 20),
    "beam1": ("beam=1 gbeam=1 rcheck=1 rbeam=1 log2w=16 wseed=11", 12),
    "rcheck0": ("beam=5 gbeam=6 rcheck=0 rbeam=0 log2w=16 wseed=13", 14),
    "rnn_default": ("beam=5 gbeam=6 rcheck=1 rbeam=5 log2w=18 wseed=1 rnn=RNN1", 16),
    "rnn_beam20": ("beam=20 gbeam=20 rcheck=1 rbeam=5 log2w=16 wseed=7 rnn=RNN2 rnn_weight=0.05 nce_bias=3.5", 10),
    "rnn_h128": ("beam=5 gbeam=6 rcheck=1 rbeam=5 log2w=18 wseed=5 rnn=RNN3", 12),
    "rnn_h48": ("beam=4 gbeam=5 rcheck=1 rbeam=4 log2w=16 wseed=6 rnn=RNN4", 10),
    "rnn_h256": ("beam=20 gbeam=20 rcheck=1 rbeam=5 log2w=16 wseed=8 rnn=RNN5", 8),
    # the reference WITHOUT its generated code (CoreHolder::initialize(nullptr): PartialNgramDynamicFeatureApply,
    # every sum through computeUnrolled4RawPerceptron) -> jpp_b200_model.float_order = JPP_B200_ORDER_DYNAMIC
    "dyn_default": ("beam=5 gbeam=6 rcheck=1 rbeam=5 log2w=20 wseed=1 dynamic=1", 20),
    "dyn_beam3_ties": ("beam=3 gbeam=4 rcheck=2 rbeam=3 log2w=12 wseed=9 dynamic=1", 20),
    "dyn_full": ("beam=5 gbeam=0 rcheck=0 rbeam=0 log2w=14 wseed=4 dynamic=1", 10),
    "dyn_rnn": ("beam=5 gbeam=6 rcheck=1 rbeam=5 log2w=18 wseed=1 rnn=RNN1 dynamic=1", 12),
}


def run(args: list[str]) -> None:
    r = subprocess.run(args, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{' '.join(args)}\n{r.stdout}\n{r.stderr}")


def gz(src: Path, dst: Path) -> None:
    with open(src, "rb") as f, gzip.GzipFile(dst, "wb", compresslevel=9, mtime=0) as g:
        shutil.copyfileobj(f, g)


def main() -> None:
    if not HARNESS.exists():
        sys.exit("build oracle/_ref first: python oracle/build_ref.py")
    tmp = Path(tempfile.mkdtemp(prefix="jpp_golden_"))
    dic = tmp / "dict.csv"
    surfaces = synth.write_dictionary(dic, 20000, 20260923)
    sents = synth.generate_corpus(64, 6, 40, 77, surfaces)
    sents[0] = sents[0][:1]                 # one-codepoint sentence
    sents[1] = "abc12３４" + sents[1][:6]     # ascii / digit unk makers
    sents[2] = sents[2][:3] + "ゴロゴロ" + "ドキドキ"  # onomatopoeia maker
    for name, prm in (("RNN1", RNN_PARAMS), ("RNN2", RNN2_PARAMS), ("RNN3", RNN3_PARAMS), ("RNN4", RNN4_PARAMS),
                      ("RNN5", RNN5_PARAMS)):
        words = synth.rnn_vocabulary(dic, prm["vocab"])
        synth.write_rnn_model(tmp / name, words, prm["layer_size"], prm["maxent_order"], prm["maxent_size"],
                              prm["nce_lnz"], prm["seed"])
    only = set(sys.argv[1:])   # names to (re)generate; default: all
    for name, (args, nsent) in CONFIGS.items():
        if only and name not in only:
            continue
        corpus = tmp / f"{name}.txt"
        chosen = sents[:nsent]
        if name == "rcheck0":
            # the reference itself reads result[-1] when a boundary has right nodes but an
            # empty global beam and rightCheck == 0 (score_processor.cc:348-359 ->
            # feature_impl_ngram_partial_kernels.h:106-109): keep the sentences it survives
            ok = []
            for s in sents:
                one = tmp / "one.txt"
                one.write_text(s + "\n", encoding="utf-8")
                r = subprocess.run([str(HARNESS), "stage", f"dict={dic}", f"corpus={one}", f"out={tmp / 'one.jppb'}",
                                    *args.split()], capture_output=True)
                if r.returncode == 0:
                    ok.append(s)
                if len(ok) == nsent:
                    break
            chosen = ok
        corpus.write_text("\n".join(chosen) + "\n", encoding="utf-8")
        a = [re.sub(r"^rnn=(RNN\d)$", lambda m: f"rnn={tmp / m.group(1)}", x) for x in args.split()]
        out = tmp / f"{name}.jppb"
        run([str(HARNESS), "stage", f"dict={dic}", f"corpus={corpus}", f"out={out}", *a])
        gz(out, HERE / f"{name}.jppb.gz")
        print(name, (HERE / f"{name}.jppb.gz").stat().st_size)

    # the reference's own differential fixture: test/jumandic/codegen.mdic, the
    # sentence and the 64-weight table of src/jumandic/shared/jumandic_codegen_test.cc:52-81
    test_src = (REF / "src/jumandic/shared/jumandic_codegen_test.cc").read_text(encoding="utf-8")
    block = test_src[test_src.index("alignas(64) float weights[] = {"):]
    block = block[:block.index("};")]
    vals = [float(x.rstrip("f")) for x in re.findall(r"[-+]\d+\.\d+f", block)]
    assert len(vals) == 64, len(vals)
    wfile = tmp / "w64.bin"
    np.array(vals, np.float32).tofile(wfile)
    sentence = re.search(r'fullAnalyze\("([^"]+)"', test_src).group(1)
    (tmp / "cg.txt").write_text(sentence + "\n", encoding="utf-8")
    for name, extra in (("ref_codegen_fixture", "beam=5 gbeam=0 rcheck=0 rbeam=0"),
                        ("ref_codegen_fixture_gbeam", "beam=5 gbeam=6 rcheck=1 rbeam=5")):
        if only and name not in only:
            continue
        out = tmp / f"{name}.jppb"
        run([str(HARNESS), "stage", f"dict={REF / 'test/jumandic/codegen.mdic'}", f"corpus={tmp / 'cg.txt'}",
             f"out={out}", "log2w=6", f"weights_file={wfile}", *extra.split()])
        gz(out, HERE / f"{name}.jppb.gz")
        print(name, (HERE / f"{name}.jppb.gz").stat().st_size)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
