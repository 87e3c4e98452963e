# round 2, run 12: end-to-end experiments -- uploads on a high-priority stream, one window computing at a time, TC kernel /
# walk shared-memory footprints that let kernels of different lanes share an SM; lane timeline traces
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="timeout 600 python bench.py --warmup 3 --steps 3 --no-cpu-baseline --no-gather-probe"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2l_$name.json 2> gpurun_out/r2l_$name.err; grep "jpp_b200 trace" gpurun_out/r2l_$name.err | tail -2; }
run default A=1
run trace JPP_B200_TRACE=1
run flat JPP_B200_FLAT_PRIORITY=1
run flat_trace JPP_B200_FLAT_PRIORITY=1 JPP_B200_TRACE=1
run turns JPP_B200_LANE_TURNS=1
run turns_trace JPP_B200_LANE_TURNS=1 JPP_B200_TRACE=1
run lanes3 JPP_B200_LANES=3
run lanes8 JPP_B200_LANES=8
for v in tc4x16 co smr16; do
  run $v JPP_B200_LIB=$PWD/jumanpp_b200/variants/libjpp_b200_$v.so
done
run co_trace JPP_B200_LIB=$PWD/jumanpp_b200/variants/libjpp_b200_co.so JPP_B200_TRACE=1
JPP_B200_LIB=$PWD/jumanpp_b200/variants/libjpp_b200_co.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_sweep.py -x -q --timeout 300 2>&1 | tail -3
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2l_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1), round(d['ms_per_step'],1), d.get('valid'))
    except Exception as e: print(f, 'ERR', e)
PY
