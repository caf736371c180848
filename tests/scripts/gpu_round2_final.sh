#!/bin/bash
# Round-2 GPU checks on the final tree: both suites, smoke, bench line, launch list of the bench command, ncu --set full of
# the C3 step kernel and of the shared-covariance step kernel (the instantiation with the proposed state in TMEM), probes.
cd $GRAFT_REPO_ROOT 2>/dev/null || true
T=${1:-r2h}
timeout 1200 python -m pytest tests/test_gpu_k4.py -q --timeout 600 > gpurun_out/${T}_k4tests.log 2>&1; echo "k4 pytest rc=$?" >> gpurun_out/${T}_k4tests.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --ignore=tests/test_gpu_k4.py > gpurun_out/${T}_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; echo "bench rc=$?" >> gpurun_out/${T}_bench.err
timeout 300 python tests/scripts/k4_probe.py > gpurun_out/${T}_k4_probe.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k4_steps -s 10 -c 1 -o gpurun_out/${T}_ncu_k4 python tests/scripts/k4_probe.py > gpurun_out/${T}_ncu_k4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_run_small -s 3 -c 1 -o gpurun_out/${T}_ncu_c1 python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu --no-workloads > gpurun_out/${T}_ncu_c1.log 2>&1
tail -3 gpurun_out/${T}_k4tests.log; tail -3 gpurun_out/${T}_gputests.log; tail -2 gpurun_out/${T}_smoke.log; tail -1 gpurun_out/${T}_bench.err; head -3 gpurun_out/${T}_k4_probe.txt
