# round 2, run 3: block walk with the compiled-in program; all GPU tests; binding
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -40
B="timeout 400 python bench.py --warmup 3"
$B --steps 5 > gpurun_out/r2c_default.json 2> gpurun_out/r2c_default.err; tail -2 gpurun_out/r2c_default.err
JPP_B200_LIB=$PWD/jumanpp_b200/variants/libjpp_b200_st3.so $B --steps 3 --no-cpu-baseline > gpurun_out/r2c_st3.json 2> gpurun_out/r2c_st3.err
$B --steps 3 --workload jumandic_beam20 --no-cpu-baseline > gpurun_out/r2c_beam20.json 2> gpurun_out/r2c_beam20.err; tail -2 gpurun_out/r2c_beam20.err
JPP_B200_NO_STATIC=1 $B --steps 3 --workload jumandic_beam20 --no-cpu-baseline > gpurun_out/r2c_beam20_nostatic.json 2> gpurun_out/r2c_beam20_nostatic.err
for f in default st3 beam20 beam20_nostatic; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2c_$f.json"))
    print("$f", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms"].items()}, d.get("parity"), d.get("valid"))
except Exception as e:
    print("$f", "FAILED", e); print(open("gpurun_out/r2c_$f.err").read()[-1500:])
PY
done
