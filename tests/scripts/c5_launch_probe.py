"""Per-launch time of short fused launches (config 5 per-GPU share on 8 GPUs: 131,072 xy-well chains x 100 steps, 10 measures,
pooled reduction after every launch, 20 launches per CUDA graph) against the number of time segments of the work queue
(ME_SEGMENTS, see plan_segments in me_api.cu).  Each setting runs in its own process.

usage: python tests/scripts/c5_launch_probe.py [chains] [segment counts ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(chains):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import metropolisengine_b200 as me
    eng = me.MetropolisEngine(("xy_well", 1.0), initial_real_params=np.zeros(2), temp=.1, n_chains=chains, seed=7, record=False)
    M, spm, L = 10, 10, 20
    eng.run_graphed(M, spm, L)
    eng.run_graphed(M, spm, L)
    best = None
    for it in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.run_graphed(M, spm, L)
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b)
        best = t if best is None else min(best, t)
    ps = eng.pooled_statistics()
    print("  chains %d  %7.1f us per launch  %.4e chain-steps/s  (grid %d x block %d)  var x0 %.6f" % (
        chains, best * 1e3 / L, chains * M * spm * L / best * 1e3, eng._grid, eng._block, float(ps["cov_real"][0, 0])), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(int(sys.argv[2]))
    else:
        chains = sys.argv[1] if len(sys.argv) > 1 else "131072"
        for segs in (sys.argv[2:] or ["auto", "1", "2", "3", "4", "5"]):
            env = dict(os.environ)
            if segs != "auto":
                env["ME_SEGMENTS"] = segs
            print("ME_SEGMENTS", segs, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "child", chains], env=env)
