#!/bin/bash
# GPU call: full gpu test-suite, v2 vs v4 apply kernel, 1- and 2-GPU bench incl. Newton
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest default rc=$?" | tee -a gpurun_out/summary.txt
( PF_APPLY_VARIANT=16 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q ) > gpurun_out/pytest_v4.log 2>&1
echo "pytest v4 rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err
echo "bench v2 rc=$?" | tee -a gpurun_out/summary.txt
PF_APPLY_VARIANT=16 timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err
echo "bench v4 rc=$?" | tee -a gpurun_out/summary.txt
PF_APPLY_VARIANT=16 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/bench_v4_n2.json 2> gpurun_out/bench_v4_n2.err
echo "bench v4 n2 rc=$?" | tee -a gpurun_out/summary.txt
tail -3 gpurun_out/pytest_gpu.log gpurun_out/pytest_v4.log
cat gpurun_out/bench_v2.json gpurun_out/bench_v4.json gpurun_out/bench_v4_n2.json
