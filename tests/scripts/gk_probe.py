"""Runtime-shape kernel (gk_run, me_generic.cu) at the cylinder shape with PER-CHAIN covariance, for several ensemble
sizes: the kernel is one thread per chain with all operands in global memory, so small ensembles leave the GPU idle.

usage: python tests/scripts/gk_probe.py [chain counts ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import metropolisengine_b200 as me

nc, spm, measures = 64, 10, 10
for chains in [int(a) for a in sys.argv[1:]] or [8192, 32768, 131072]:
    eng = me.MetropolisEngine(me.BuiltinEnergy("cylinder", 10.0, -1.0, 0.05, 1.0, reject=True),
                              initial_real_params=np.array([0.0]), initial_complex_params=np.zeros(nc, dtype=complex),
                              temp=.1, n_chains=chains, seed=5, record=False, sampling_width=0.02)
    eng.run(55, 2)
    torch.cuda.synchronize()
    for what, m, s in (("steps + measure every 10", measures, spm), ("steps only (one measure per 100)", 1, 100)):
        best = None
        for it in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.run(m, s)
            b.record()
            torch.cuda.synchronize()
            t = a.elapsed_time(b)
            best = t if best is None else min(best, t)
        d = 1 + 2 * nc
        gbs = chains * m * s * 8.0 * (nc * nc + 7 * d) / (best * 1e-3) / 1e9
        print("chains %7d  %-34s %9.2f ms  %.3e chain-steps/s  %.0f GB/s algorithmic" % (chains, what, best, chains * m * s / best * 1e3, gbs),
              flush=True)
    del eng
    torch.cuda.empty_cache()
