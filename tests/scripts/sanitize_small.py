"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck):
   compute-sanitizer --tool memcheck python tests/scripts/sanitize_small.py
   (compute-sanitizer is closed on the round-1 GPU pool; the script also runs bare as an end-to-end smoke.)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import metropolisengine_b200 as me

os.environ["ME_SEGMENTS"] = "4"
# fused small shape, segmented launch through the work queue, time series, pooled moments
eng = me.MetropolisEngine(("xy_well", 1.0), initial_real_params=np.array([0.1, -0.2]), temp=.1, n_chains=32 * 592 + 7, seed=3)
eng.run(8, 5)
eng.pooled_statistics()
eng.statistical_inefficiency(column=0, n_chains=64)
eng.detect_equilibration(column=0, n_chains=64, nskip=2)
# mixed shape: shared-memory log table, group and magnitude-phase moves
eng = me.MetropolisEngine(("mixed_well", 1.0, -1.0, 0.5, 1.0), initial_real_params=np.zeros(3),
                          initial_complex_params=np.full(4, 0.3 + 0.1j), temp=.1, n_chains=200, seed=4,
                          complex_sample_method="magnitude-phase")
eng.run(55, 3)
eng.step_real_group(2)
eng.step_complex_group(2)
eng.measure()
# runtime-compiled functor
src = """__device__ double me_user_energy(const double* x, const double* cr, const double* ci, const double* k) {
    return k[0] * (x[0] * x[0] + x[1] * x[1]); }"""
eng = me.MetropolisEngine(me.CudaEnergy(src, consts=[1.0]), initial_real_params=np.zeros(2), temp=.1, n_chains=32 * 600,
                          seed=5, record=False)
eng.run(8, 5)
# runtime-shape (large) per-chain path
eng = me.MetropolisEngine(me.BuiltinEnergy("cylinder", 10.0, -1.0, 0.05, 1.0, reject=True), initial_real_params=np.array([0.2]),
                          initial_complex_params=np.zeros(64, dtype=complex), temp=.1, sampling_width=0.012, n_chains=64, seed=6)
eng.run(3, 2)
eng.step_complex_group_magnitude()
eng.step_complex_group_phase()
eng.measure()
eng.check_status()
# shared-covariance tcgen05 path with the asynchronous factor refresh
eng = me.SharedCovarianceEngine(temp=.1, n_chains=128 * 5, seed=7, sampling_width=0.004)
eng.run(54, 2)
eng.synchronize_refresh()
_ = eng.covariance_matrix_complex
torch.cuda.synchronize()
print("sanitize_small: done")
