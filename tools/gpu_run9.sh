# round 2, run 9: full GPU suite with the tcgen05 level kernel as the default; k_t0 with shared-memory primitives;
# gather probe; TC pipeline variants (3x32, 4x16, 6x16); 10 000-sentence parity sample against the reference
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 400 2>&1 | tail -8
B="timeout 600 python bench.py --warmup 3"
$B --steps 5 --parity-sample 10000 > gpurun_out/r2h_default.json 2> gpurun_out/r2h_default.err; tail -3 gpurun_out/r2h_default.err
for v in g3t32 g4t16 g6t16; do
  JPP_B200_LIB=$PWD/jumanpp_b200/variants/libjpp_b200_$v.so $B --steps 3 --no-cpu-baseline --no-gather-probe > gpurun_out/r2h_$v.json 2> gpurun_out/r2h_$v.err; tail -2 gpurun_out/r2h_$v.err
done
JPP_B200_RNN_TC=0 $B --steps 3 --no-cpu-baseline --no-gather-probe > gpurun_out/r2h_ffma.json 2> gpurun_out/r2h_ffma.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2h_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['e2e']['value']), {k:round(v,2) for k,v in d['stage_ms'].items()}, d.get('gather_roofline'), d.get('parity'), d.get('valid'))
    except Exception as e: print(f, 'ERR', e)
PY
