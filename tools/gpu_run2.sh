# round 2, run 2: warp-scope walk + compiled-in n-gram program: tests, A/B benches
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 150 env JPP_B200_WAVEFRONT=1 python -m pytest tests/test_gpu_failure_isolation.py -x -q --timeout 45 -k beam20 2>&1 | tail -40
timeout 700 python -m pytest tests -m gpu -x -q --timeout 150 2>&1 | tail -25
B="timeout 400 python bench.py --warmup 3"
$B --steps 5 > gpurun_out/r2b_default.json 2> gpurun_out/r2b_default.err; tail -2 gpurun_out/r2b_default.err
JPP_B200_NO_STATIC=1 $B --steps 3 --no-cpu-baseline > gpurun_out/r2b_nostatic.json 2> gpurun_out/r2b_nostatic.err
JPP_B200_BLOCK_WALK=1 $B --steps 3 --no-cpu-baseline > gpurun_out/r2b_blockwalk.json 2> gpurun_out/r2b_blockwalk.err
JPP_B200_LIB=$PWD/jumanpp_b200/variants/libjpp_b200_mb3.so $B --steps 3 --no-cpu-baseline > gpurun_out/r2b_mb3.json 2> gpurun_out/r2b_mb3.err
JPP_B200_LIB=$PWD/jumanpp_b200/variants/libjpp_b200_mb5.so $B --steps 3 --no-cpu-baseline > gpurun_out/r2b_mb5.json 2> gpurun_out/r2b_mb5.err
$B --steps 3 --workload jumandic_beam20 > gpurun_out/r2b_beam20.json 2> gpurun_out/r2b_beam20.err; tail -2 gpurun_out/r2b_beam20.err
for f in default nostatic blockwalk mb3 mb5 beam20; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2b_$f.json"))
    print("$f", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms"].items()}, d.get("parity"), d.get("valid"))
except Exception as e:
    print("$f", "FAILED", e); print(open("gpurun_out/r2b_$f.err").read()[-1500:])
PY
done
