# round 2, run 7: binding tests; tcgen05 RNN (batched loads); pipelined e2e_full; BASELINE configs[3] at 1M sentences
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_binding.py tests/test_gpu_resident_form.py tests/test_gpu_failure_isolation.py -q --timeout 300 2>&1 | tail -30
JPP_B200_RNN_TC=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -q --timeout 200 -k "rnn_h128 or properties or reference_post or idempotent or invariants or prefix or order" 2>&1 | tail -5
B="timeout 600 python bench.py --warmup 3"
JPP_B200_RNN_TC=1 $B --steps 5 --e2e-full > gpurun_out/r2f_tc.json 2> gpurun_out/r2f_tc.err; tail -2 gpurun_out/r2f_tc.err
$B --steps 5 --no-cpu-baseline > gpurun_out/r2f_default.json 2> gpurun_out/r2f_default.err
timeout 1500 python bench.py --warmup 3 --steps 3 --workload jumandic_beam5_linear --sentences 1000000 > gpurun_out/r2f_linear_1m.json 2> gpurun_out/r2f_linear_1m.err; tail -4 gpurun_out/r2f_linear_1m.err
for f in tc default linear_1m; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2f_$f.json"))
    print("$f", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms"].items()}, d.get("parity"), d.get("valid"), d.get("cpu_baseline",{}).get("value"))
    if "e2e_full" in d: print("   e2e_full", json.dumps(d["e2e_full"])[:1600])
except Exception as e:
    print("$f", "FAILED", e); print(open("gpurun_out/r2f_$f.err").read()[-1500:])
PY
done
