# round 2, run 10: k_t0 back on the private array (+ 16-byte-store and shared-memory variants), TC level kernel 6 x 16,
# gather cache-hint variants, deeper gather batches, small-table runs (how much of the walk is table misses), ncu of the probe
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -x -q --timeout 300 2>&1 | tail -4
B="timeout 600 python bench.py --warmup 3 --steps 3 --no-cpu-baseline"
$B > gpurun_out/r2i_default.json 2> gpurun_out/r2i_default.err; tail -2 gpurun_out/r2i_default.err
for v in t0v16 t0smem gna gl gnal b16 b16m3; do
  JPP_B200_LIB=$PWD/jumanpp_b200/variants/libjpp_b200_$v.so $B --no-gather-probe > gpurun_out/r2i_$v.json 2> gpurun_out/r2i_$v.err; tail -1 gpurun_out/r2i_$v.err
done
for w in 20 24; do
  $B --log2w $w > gpurun_out/r2i_w$w.json 2> gpurun_out/r2i_w$w.err; tail -1 gpurun_out/r2i_w$w.err
done
timeout 600 ncu --set full --clock-control none -k regex:k_gather_probe -c 3 -o gpurun_out/r2i_probe python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2i_ncu.log 2>&1
ncu -i gpurun_out/r2i_probe.ncu-rep --page raw --csv > gpurun_out/r2i_probe_raw.csv 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        g=d.get('gather_roofline') or {}
        print(f, round(d['value']), round(d['e2e']['value']), {k:round(v,2) for k,v in d['stage_ms'].items() if k not in ('upload',)}, g.get('peak_gathers_per_s'), d.get('valid'))
    except Exception as e: print(f, 'ERR', e)
PY
