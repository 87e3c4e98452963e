# round 2, run 13: upload turn passed when the copies are queued; L2 fetch granularity 32 / 64 / 128; co-residency build
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B="timeout 600 python bench.py --warmup 3 --steps 3 --no-cpu-baseline"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2m_$name.json 2> gpurun_out/r2m_$name.err; grep "jpp_b200 trace\|L2 fetch" gpurun_out/r2m_$name.err | tail -2; }
CO=JPP_B200_LIB=$PWD/jumanpp_b200/variants/libjpp_b200_co.so
run default JPP_B200_TRACE=1
run co $CO JPP_B200_TRACE=1
run oldturns JPP_B200_UPLOAD_TURNS=1
run l2f32 JPP_B200_L2_FETCH=32
run l2f128 JPP_B200_L2_FETCH=128
run co_l2f32 $CO JPP_B200_L2_FETCH=32
run co_lanes4 $CO JPP_B200_LANES=4
run co_lanes8 $CO JPP_B200_LANES=8
env $CO $B --batch-form resident > gpurun_out/r2m_co_resident.json 2> gpurun_out/r2m_co_resident.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2m_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        g=d.get('gather_roofline') or {}
        print(f, round(d['value']), round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1), {k:round(v,2) for k,v in d['stage_ms'].items() if k not in ('upload',)}, round((g.get('peak_gathers_per_s') or 0)/1e9,1), d.get('valid'))
    except Exception as e: print(f, 'ERR', e)
PY
