"""Throughput of the C2 step kernel as a function of the warps per SM sub-partition:
592 sub-partitions x {1, 2, 3, 4} warps exactly, plus the BASELINE count (65,536 chains = 3.46 warps on average)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import metropolisengine_b200 as me

n_smsp = 4 * torch.cuda.get_device_properties(0).multi_processor_count
for label, chains in [("1 warp / SMSP", n_smsp * 32), ("2 warps / SMSP", n_smsp * 64), ("3 warps / SMSP", n_smsp * 96),
                      ("4 warps / SMSP", n_smsp * 128), ("65,536 chains", 65536), ("8 warps worth / SMSP", n_smsp * 256)]:
    eng = me.MetropolisEngine(("xy_well", 1.0), initial_real_params=np.zeros(2), temp=.1, n_chains=chains, seed=3,
                              record=False)
    eng.run(300, 10)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.run(3000, 10)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print("%-22s chains=%8d grid=%5d block=%3d  %8.2f ms  %.3e chain-steps/s  (%.1f cycles per warp-step per SMSP-warp)"
          % (label, chains, eng._grid, eng._block, ms, chains * 30000 / ms * 1e3,
             ms * 1e-3 * 1.965e9 / 30000))
