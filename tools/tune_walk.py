#!/usr/bin/env python3
"""Build tuning variants of libjpp_b200.so (here, no GPU needed) and time the beam-walk
kernel of each on the GPU box:   python tools/tune_walk.py build | run"""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
VARIANTS = {
    "mb4_b1_s0_k32": ["JPP_WALK_MINBLOCKS=4", "JPP_GATHER_BATCH=1", "JPP_STAGE_CANDS=0", "JPP_KCHUNK=32"],
    "mb4_b1_s1_k16": ["JPP_WALK_MINBLOCKS=4", "JPP_GATHER_BATCH=1", "JPP_STAGE_CANDS=1", "JPP_KCHUNK=16"],
    "mb6_b1_s1_k16": ["JPP_WALK_MINBLOCKS=6", "JPP_GATHER_BATCH=1", "JPP_STAGE_CANDS=1", "JPP_KCHUNK=16"],
    "mb8_b1_s1_k16": ["JPP_WALK_MINBLOCKS=8", "JPP_GATHER_BATCH=1", "JPP_STAGE_CANDS=1", "JPP_KCHUNK=16"],
    "mb4_b2_s1_k16": ["JPP_WALK_MINBLOCKS=4", "JPP_GATHER_BATCH=2", "JPP_STAGE_CANDS=1", "JPP_KCHUNK=16"],
    "mb4_b4_s1_k16": ["JPP_WALK_MINBLOCKS=4", "JPP_GATHER_BATCH=4", "JPP_STAGE_CANDS=1", "JPP_KCHUNK=16"],
    "mb6_b2_s1_k16": ["JPP_WALK_MINBLOCKS=6", "JPP_GATHER_BATCH=2", "JPP_STAGE_CANDS=1", "JPP_KCHUNK=16"],
}
VDIR = ROOT / "tools" / "_variants"
if sys.argv[1] == "build":
    from jumanpp_b200 import build as b
    VDIR.mkdir(exist_ok=True)
    for name, defs in VARIANTS.items():
        print(name, b.build(force=True, defines=defs, out=VDIR / f"libjpp_b200_{name}.so"))
else:
    extra = sys.argv[2:]
    for name in VARIANTS:
        env = dict(os.environ, JPP_B200_LIB=str(VDIR / f"libjpp_b200_{name}.so"))
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "3", "--warmup", "3", "--no-cpu-baseline", *extra],
                           env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(name, round(d["value"]), {k: round(v, 1) for k, v in d["stage_ms"].items()}, flush=True)
        except Exception as e:
            print(name, "FAILED", r.stderr[-500:], flush=True)
