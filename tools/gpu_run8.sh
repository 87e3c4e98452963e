# round 2, run 8: binding tests; tcgen05 RNN (two pipelines per block); warp-scope walk at 64 / 85 registers; configs[2] with its CPU arm
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_binding.py -q --timeout 300 2>&1 | tail -15
JPP_B200_RNN_TC=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -q --timeout 200 -k "rnn_h128 or properties or reference_post or idempotent or invariants or prefix or order" 2>&1 | tail -5
JPP_B200_WARP_WALK=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_sweep.py tests/test_gpu_failure_isolation.py -q --timeout 200 2>&1 | tail -5
B="timeout 600 python bench.py --warmup 3"
JPP_B200_RNN_TC=1 $B --steps 5 --no-cpu-baseline > gpurun_out/r2g_tc.json 2> gpurun_out/r2g_tc.err
JPP_B200_WARP_WALK=1 $B --steps 3 --no-cpu-baseline > gpurun_out/r2g_ww8.json 2> gpurun_out/r2g_ww8.err
JPP_B200_WARP_WALK=1 JPP_B200_LIB=$PWD/jumanpp_b200/variants/libjpp_b200_ww6.so $B --steps 3 --no-cpu-baseline > gpurun_out/r2g_ww6.json 2> gpurun_out/r2g_ww6.err
JPP_B200_WARP_WALK=1 JPP_B200_NO_STATIC=1 $B --steps 3 --no-cpu-baseline > gpurun_out/r2g_ww8_nostatic.json 2> gpurun_out/r2g_ww8_nostatic.err
$B --steps 3 --workload jumandic_beam20 > gpurun_out/r2g_beam20.json 2> gpurun_out/r2g_beam20.err
JPP_B200_WARP_WALK=1 $B --steps 3 --workload jumandic_beam20 --no-cpu-baseline > gpurun_out/r2g_beam20_ww.json 2> gpurun_out/r2g_beam20_ww.err
for f in tc ww8 ww6 ww8_nostatic beam20 beam20_ww; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2g_$f.json"))
    print("$f", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms"].items()}, d.get("parity"), d.get("valid"), d.get("cpu_baseline",{}).get("value"))
except Exception as e:
    print("$f", "FAILED", e); print(open("gpurun_out/r2g_$f.err").read()[-1500:])
PY
done
