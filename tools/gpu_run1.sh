set -x
cd $GRAFT_REPO_ROOT
nproc; cat /sys/fs/cgroup/cpu.max; python -c "import os; print(len(os.sched_getaffinity(0)))"; lscpu | head -20; free -g | head -2; nvidia-smi topo -m | head -20
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err; tail -3 gpurun_out/r2a_bench_default.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload jumandic_beam20 > gpurun_out/r2a_bench_beam20.json 2> gpurun_out/r2a_bench_beam20.err; tail -3 gpurun_out/r2a_bench_beam20.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload jumandic_beam5_linear > gpurun_out/r2a_bench_linear.json 2> gpurun_out/r2a_bench_linear.err; tail -3 gpurun_out/r2a_bench_linear.err
