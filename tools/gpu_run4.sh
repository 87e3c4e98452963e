# round 2, run 4: why the binding tests fail
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_binding.py -x -q --timeout 200 -k "rnn and default" 2>&1 | tail -60
