# round 2, run 12 (8 GPUs): BASELINE configs[4] -- ONE corpus of 1 000 000 sentences cut into 8 contiguous shards
# (balanced by codepoints), one rank per GPU, 1-best paths gathered back on rank 0 in corpus order over NCCL
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2k_topo.txt 2>&1
nproc; cat /sys/fs/cgroup/cpu.max
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 8 --steps 3 --warmup 3 --shard-corpus --sentences 1000000 > gpurun_out/r2k_8gpu_sharded.json 2> gpurun_out/r2k_8gpu_sharded.err
tail -5 gpurun_out/r2k_8gpu_sharded.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2k_8gpu_sharded.json').read().strip().splitlines()[-1])
    print(round(d['value']), round(d['e2e']['value']), d.get('sharded'), d.get('parity'), d.get('valid'), d.get('numa'))
except Exception as e: print('ERR', e)
PY
