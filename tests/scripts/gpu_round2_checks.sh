#!/bin/bash
# Round-2 GPU checks: the shared-covariance suite first (own process: a trapped kernel poisons its CUDA context), then the
# rest, the smoke entry, kernel probes (v2, v1 baseline, v2 without TMA), the bench line and two ncu captures.
cd $GRAFT_REPO_ROOT 2>/dev/null || true
T=${1:-r2c}
timeout 1200 python -m pytest tests/test_gpu_k4.py -q --timeout 600 > gpurun_out/${T}_k4tests.log 2>&1; echo "k4 pytest rc=$?" >> gpurun_out/${T}_k4tests.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --ignore=tests/test_gpu_k4.py > gpurun_out/${T}_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
timeout 300 python tests/scripts/k4_probe.py > gpurun_out/${T}_k4_probe_v2.txt 2>&1
ME_K4_NO_TMA=1 timeout 300 python tests/scripts/k4_probe.py > gpurun_out/${T}_k4_probe_v2_notma.txt 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; echo "bench rc=$?" >> gpurun_out/${T}_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k4_steps -s 10 -c 1 -o gpurun_out/${T}_ncu_k4 python tests/scripts/k4_probe.py > gpurun_out/${T}_ncu_k4.log 2>&1
# (register-cap variants: profiles/r02_nreg_probe_small_shapes.txt)
ME_B200_LIB=$PWD/metropolisengine_b200/lib/variants/k4wide_libme_b200.so timeout 300 python tests/scripts/k4_probe.py > gpurun_out/${T}_k4_probe_32warps.txt 2>&1
ME_B200_LIB=$PWD/metropolisengine_b200/lib/variants/k4wide_libme_b200.so timeout 600 python -m pytest tests/test_gpu_k4.py -q --timeout 600 -k oracle > gpurun_out/${T}_k4tests_32warps.log 2>&1
timeout 300 python tests/scripts/scale_probe.py > gpurun_out/${T}_scale_probe.txt 2>&1
tail -3 gpurun_out/${T}_k4tests.log; tail -3 gpurun_out/${T}_gputests.log; tail -2 gpurun_out/${T}_smoke.log; head -3 gpurun_out/${T}_k4_probe_v2.txt; tail -1 gpurun_out/${T}_bench.err
