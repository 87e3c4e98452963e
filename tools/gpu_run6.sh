# round 2, run 6: all GPU tests; tcgen05 RNN level kernel on/off; ncu launch list + full captures
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -80
JPP_B200_RNN_TC=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_resident_form.py -q --timeout 200 -k "rnn_h128 or properties or reference_post or idempotent or invariants or prefix or order" 2>&1 | tail -40
B="timeout 500 python bench.py --warmup 3"
$B --steps 5 > gpurun_out/r2e_default.json 2> gpurun_out/r2e_default.err; tail -2 gpurun_out/r2e_default.err
JPP_B200_RNN_TC=1 $B --steps 5 > gpurun_out/r2e_tc.json 2> gpurun_out/r2e_tc.err; tail -2 gpurun_out/r2e_tc.err
JPP_B200_LIB=$PWD/jumanpp_b200/variants/libjpp_b200_mb5.so $B --steps 3 --no-cpu-baseline > gpurun_out/r2e_mb5.json 2> gpurun_out/r2e_mb5.err
for f in default tc mb5; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2e_$f.json"))
    print("$f", round(d["value"]), round(d["e2e"]["value"]), {k: round(v,1) for k,v in d["stage_ms"].items()}, d.get("parity"), d.get("valid"))
except Exception as e:
    print("$f", "FAILED", e); print(open("gpurun_out/r2e_$f.err").read()[-1500:])
PY
done
# ncu: every launch of one step with its device time, then full captures of the dominant kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2e_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_walk_fused|k_t0|k_rnn_tree" -c 3 -o gpurun_out/r2e_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2e_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
