"""Throughput of the unfused plugin path (torch-vectorised energy callable: me_propose -> callable -> me_accept)
against the fused device functor, same physics (xy-well), 65,536 chains."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import metropolisengine_b200 as me

n = 65536
for graph in (False, True):
    eng = me.MetropolisEngine(lambda r, c: (r * r).sum(dim=1), initial_real_params=np.zeros(2), temp=.1, n_chains=n, seed=3,
                              record=False, graph_callable=graph)
    eng.run(5, 10)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.run(50, 10)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print("torch callable, graph_callable=%-5s: %.3e chain-steps/s, %.1f us per ensemble step"
          % (graph, n * 500 / ms * 1e3, ms * 1e3 / 500))
