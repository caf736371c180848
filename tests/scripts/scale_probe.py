import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import metropolisengine_b200 as me
for n in (65536, 262144, 1048576, 4194304, 16777216):
    for wl in ("c2", "c3"):
        if wl == "c2":
            eng = me.MetropolisEngine(("xy_well", 1.0), initial_real_params=np.zeros(2), temp=.1, n_chains=n, record=False)
        else:
            if n > 4194304: continue
            eng = me.MetropolisEngine(("mixed_well", 1.0, -1.0, 0.5, 1.0), initial_real_params=np.zeros(3), initial_complex_params=np.zeros(4, dtype=complex), temp=.1, n_chains=n, record=False)
        steps = max(200, int(2e9 // n // 10 * 10)) if wl == "c2" else max(100, int(3e8 // n // 10 * 10))
        eng.run(5, 10); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.run(steps // 10, 10); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print("%s n=%9d block=%3d grid=%7d steps=%6d  %.3e chain-steps/s" % (wl, n, eng._block, eng._grid, steps, n * steps / ms * 1e3), flush=True)
        del eng
