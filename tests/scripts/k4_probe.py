import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import metropolisengine_b200 as me
n = 32768
eng = me.SharedCovarianceEngine(energy_consts=(10.0, -1.0, 0.05, 1.0), temp=.1, n_chains=n, seed=1, record=False)
eng.run(60, 10); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); eng.step(200); b.record(); torch.cuda.synchronize()
print("k4 steps only: %.3e chain-steps/s (%.2f us per ensemble step)" % (n * 200 / a.elapsed_time(b) * 1e3, a.elapsed_time(b) * 1e3 / 200))
a.record(); eng.run(50, 10); b.record(); torch.cuda.synchronize()
print("k4 run(50,10) incl. measure + pooled covariance: %.3e chain-steps/s" % (n * 500 / a.elapsed_time(b) * 1e3))
print("acceptance", eng.acceptance_rate, "sigma", eng.sampling_width)
