"""Warm per-kernel timings of the shared-covariance (config 4) pipeline with CUDA events: stepping, measure, pooled
moments, factor refresh.  Run on a GPU box: python tests/scripts/k4_probe.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import metropolisengine_b200 as me
from metropolisengine_b200.engine import _ptr

n = 32768
eng = me.SharedCovarianceEngine(energy_consts=(10.0, -1.0, 0.05, 1.0), temp=.1, n_chains=n, seed=1, record=False)
eng.run(60, 10)
torch.cuda.synchronize()


def timed(label, fn, reps=50):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / reps
    print("%-46s %9.1f us" % (label, us))
    return us


us = timed("k4_steps x10 (one launch)", lambda: eng.step(10))
print("   -> steps only: %.3e chain-steps/s" % (n * 10 / us * 1e6))
timed("k4_steps x100 (one launch)", lambda: eng.step(100), reps=10)
L = eng._lib
timed("k4_measure (no rows)", lambda: L.me_k4_measure(eng._h, None, 0, eng._stream()))
ts = torch.empty((4, eng._lay.TS_COLS, n), dtype=torch.float64, device=eng.device)
timed("k4_measure (row store)", lambda: L.me_k4_measure(eng._h, _ptr(ts), 0, eng._stream()))
timed("k4_moments (stage 1 + 2a + 2b)", lambda: L.me_k4_moments(eng._h, _ptr(eng._shift), _ptr(eng._scratch), eng._scratch.numel(),
                                                                _ptr(eng._inc_full), None, None, eng._stream()))
eng.synchronize_refresh()
spare = torch.zeros_like(eng._factors[0])
spare_sa = torch.ones(1, dtype=torch.float64, device=eng.device)
timed("k4_refactor", lambda: L.me_k4_refactor(eng._h, _ptr(eng._mom), _ptr(eng._inc_full), 0, _ptr(eng._cov_c), _ptr(eng._cov_a),
                                              _ptr(spare), _ptr(spare_sa), _ptr(eng._psd_status), eng._stream()))
if hasattr(L, "me_k4_step_measure"):
    timed("k4_steps x10 + measure tail + stage 2a/2b", lambda: L.me_k4_step_measure(
        eng._h, 10, _ptr(eng._s_a), None, 0, _ptr(eng._shift), _ptr(eng._scratch), eng._scratch.numel(), _ptr(eng._inc_full),
        None, None, eng._stream()))
us = timed("run(1, 10), separate measure / FP64 moments kernels", lambda: eng.run(1, 10, fused_measure=False))
print("   -> %.3e chain-steps/s" % (n * 10 / us * 1e6))
us = timed("run(1, 10): one-launch block + adaptation", lambda: eng.run(1, 10))
print("   -> with measure every 10 (refresh on the side stream): %.3e chain-steps/s" % (n * 10 / us * 1e6))
seq = me.SharedCovarianceEngine(energy_consts=(10.0, -1.0, 0.05, 1.0), temp=.1, n_chains=n, seed=1, record=False,
                                async_refresh=False)
seq.run(60, 10)
us = timed("run(1, 10), async_refresh=False", lambda: seq.run(1, 10))
print("   -> with measure every 10 (sequential refresh): %.3e chain-steps/s" % (n * 10 / us * 1e6))
print("acceptance", eng.acceptance_rate, "sigma", eng.sampling_width)
