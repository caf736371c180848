#!/bin/bash
# Round-2 multi-GPU checks (gpurun --gpus N): sharding invariance + library-side collective (NCCL loaded at run time, also
# from inside a CUDA graph), then the bench line at N ranks exactly as the driver launches it.
cd $GRAFT_REPO_ROOT 2>/dev/null || true
N=${1:-2}
T=${2:-r2m}
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 600 > gpurun_out/${T}_multi_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_multi_tests.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/${T}_bench_${N}gpu.log 2> gpurun_out/${T}_bench_${N}gpu.err; echo "bench rc=$?" >> gpurun_out/${T}_bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/${T}_refarm_${N}gpu.log 2>&1; echo "ref rc=$?" >> gpurun_out/${T}_refarm_${N}gpu.log
tail -3 gpurun_out/${T}_multi_tests.log; tail -2 gpurun_out/${T}_bench_${N}gpu.err; tail -c 300 gpurun_out/${T}_refarm_${N}gpu.log
