# round 2, run 16: kernel-selection switches read at handle creation; GPU tests incl. the table-driven / FFMA paths; smoke()
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
