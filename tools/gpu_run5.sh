# round 2, run 5: all GPU tests; block size / register A/B of the fused walk; binding end to end
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30
B="timeout 500 python bench.py --warmup 3"
$B --steps 5 --e2e-full > gpurun_out/r2d_default.json 2> gpurun_out/r2d_default.err; tail -3 gpurun_out/r2d_default.err
JPP_B200_FUSED_SENT=16 $B --steps 3 --no-cpu-baseline > gpurun_out/r2d_sent16.json 2> gpurun_out/r2d_sent16.err
JPP_B200_LIB=$PWD/jumanpp_b200/variants/libjpp_b200_mb4.so $B --steps 3 --no-cpu-baseline > gpurun_out/r2d_mb4.json 2> gpurun_out/r2d_mb4.err
JPP_B200_FUSED_SENT=16 JPP_B200_LIB=$PWD/jumanpp_b200/variants/libjpp_b200_mb4.so $B --steps 3 --no-cpu-baseline > gpurun_out/r2d_mb4_sent16.json 2> gpurun_out/r2d_mb4_sent16.err
$B --steps 3 --batch-form resident --no-cpu-baseline > gpurun_out/r2d_resident.json 2> gpurun_out/r2d_resident.err
JPP_B200_FUSED_SENT=16 $B --steps 3 --workload jumandic_beam20 --no-cpu-baseline > gpurun_out/r2d_beam20_sent16.json 2> gpurun_out/r2d_beam20_sent16.err
for f in default sent16 mb4 mb4_sent16 resident beam20_sent16; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2d_$f.json"))
    print("$f", round(d["value"]), round(d["e2e"]["value"]), d["e2e"]["h2d_bytes_per_step"], {k: round(v,1) for k,v in d["stage_ms"].items()}, d.get("parity"), d.get("valid"))
    if "e2e_full" in d: print("   e2e_full", json.dumps(d["e2e_full"])[:1500])
except Exception as e:
    print("$f", "FAILED", e); print(open("gpurun_out/r2d_$f.err").read()[-1500:])
PY
done
