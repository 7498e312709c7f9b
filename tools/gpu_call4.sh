#!/bin/bash
# regression tests + bench + ncu (launch list and full capture of the v4 apply kernel)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-newton --e2e-steps 1 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply3d_v4 -s 6 -c 2 -o gpurun_out/prof_v4 -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-newton --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
cat gpurun_out/bench_n1.json
