# round 2, final run: every GPU test, the default bench line (CPU arm, 10 000-sentence parity, gather probe, e2e_full through
# the binding), ncu launch list + full captures of the final kernels, then configs[2] and configs[3] with the final build
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 2>&1 | tail -6
B="timeout 900 python bench.py --warmup 3"
$B --steps 5 --parity-sample 10000 --e2e-full > gpurun_out/r2n_default.json 2> gpurun_out/r2n_default.err; tail -2 gpurun_out/r2n_default.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gather-probe > gpurun_out/r2n_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_walk_fused|k_t0_static|k_rnn_tree" -c 3 -o gpurun_out/r2n_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gather-probe > gpurun_out/r2n_ncu_full.log 2>&1
ncu -i gpurun_out/r2n_full.ncu-rep --page raw --csv > gpurun_out/r2n_full_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:"k_rnn_level_tc" -s 8 -c 2 -o gpurun_out/r2n_tc python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gather-probe > gpurun_out/r2n_ncu_tc.log 2>&1
ncu -i gpurun_out/r2n_tc.ncu-rep --page raw --csv > gpurun_out/r2n_tc_raw.csv 2>/dev/null
$B --steps 3 --workload jumandic_beam20 > gpurun_out/r2n_beam20.json 2> gpurun_out/r2n_beam20.err; tail -1 gpurun_out/r2n_beam20.err
$B --steps 2 --workload jumandic_beam5_linear --sentences 1000000 > gpurun_out/r2n_linear_1m.json 2> gpurun_out/r2n_linear_1m.err; tail -1 gpurun_out/r2n_linear_1m.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2n_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['e2e']['value']), round(d['e2e']['ms_per_step'],1), {k:round(v,2) for k,v in d['stage_ms'].items() if k not in ('upload',)}, d.get('parity'), (d.get('cpu_baseline') or {}).get('value'), (d.get('e2e_full') or {}).get('value'), d.get('valid'))
    except Exception as e: print(f, 'ERR', e)
PY
