# round 2, run 15: the committed build once more -- all GPU tests, smoke(), the default bench line
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --warmup 3 --steps 5 > gpurun_out/r2o_default.json 2> gpurun_out/r2o_default.err; tail -2 gpurun_out/r2o_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2o_default.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['e2e']['value']), {k:round(v,2) for k,v in d['stage_ms'].items()}, d.get('parity'), d['cpu_baseline']['value'], d.get('valid'))
PY
