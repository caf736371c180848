import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import metropolisengine_b200 as me
n, M, spm = 131072, 50, 10
eng = me.MetropolisEngine(me.BuiltinEnergy("cylinder", 10.0, -1.0, 0.05, 1.0, reject=True), initial_real_params=np.array([0.0]),
                          initial_complex_params=np.zeros(8, dtype=complex), temp=.1, n_chains=n, seed=5, record=False, sampling_width=0.02)
eng.run(60, 2)
best = None
for it in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.run(M, spm); b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b); best = t if best is None else min(best, t)
print("1r+8c  %8.3f ms  %.4e chain-steps/s  acceptance %.4f" % (best, n * M * spm / best * 1e3, eng.acceptance_rate))
