# round 2, run 11: compiled-in per-node program (k_t0_static), software-pipelined TC level kernel with level records;
# register variants of both; e2e A/B (NUMA binding, lanes)
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_config_sweep.py tests/test_gpu_failure_isolation.py tests/test_gpu_resident_form.py -x -q --timeout 300 2>&1 | tail -4
B="timeout 600 python bench.py --warmup 3 --steps 3"
$B --steps 5 --parity-sample 10000 > gpurun_out/r2j_default.json 2> gpurun_out/r2j_default.err; tail -2 gpurun_out/r2j_default.err
for v in t0m8 t0m10 t0m4 tc3x32 tc4x16; do
  JPP_B200_LIB=$PWD/jumanpp_b200/variants/libjpp_b200_$v.so $B --no-cpu-baseline --no-gather-probe > gpurun_out/r2j_$v.json 2> gpurun_out/r2j_$v.err; tail -1 gpurun_out/r2j_$v.err
done
JPP_B200_NO_STATIC_T0=1 $B --no-cpu-baseline --no-gather-probe > gpurun_out/r2j_tablet0.json 2> gpurun_out/r2j_tablet0.err
$B --no-cpu-baseline --no-gather-probe --no-numa-bind > gpurun_out/r2j_nonuma.json 2> gpurun_out/r2j_nonuma.err
JPP_B200_LANES=4 $B --no-cpu-baseline --no-gather-probe > gpurun_out/r2j_lanes4.json 2> gpurun_out/r2j_lanes4.err
JPP_B200_LANES=8 $B --no-cpu-baseline --no-gather-probe > gpurun_out/r2j_lanes8.json 2> gpurun_out/r2j_lanes8.err
$B --no-cpu-baseline --no-gather-probe --batch-form resident > gpurun_out/r2j_resident.json 2> gpurun_out/r2j_resident.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['e2e']['value']), {k:round(v,2) for k,v in d['stage_ms'].items() if k not in ('upload',)}, d.get('parity'), d.get('valid'))
    except Exception as e: print(f, 'ERR', e)
PY
