"""Kernel time of the mixed 3r+4c shape (BASELINE config 3: 262,144 chains, measure every 10) for library builds with
different register caps / measure-block variants (metropolisengine_b200/lib/variants/libme_b200_c3*.so) and CTA sizes
(ME_BLOCK).  Each (library, CTA size) runs in its own process: the library is chosen at import time (ME_B200_LIB).

usage: python tests/scripts/c3_probe.py [block sizes, default "64 128"]     (parent: loops over the variants found)
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import metropolisengine_b200 as me
    n, M, spm = 262144, 100, 10
    eng = me.MetropolisEngine(("mixed_well", 1.0, -1.0, 0.5, 1.0), initial_real_params=np.zeros(3),
                              initial_complex_params=np.zeros(4, dtype=complex), temp=.1, n_chains=n, seed=2024,
                              ts_chunk_bytes=M * 14 * n * 8)
    eng.reserve_rows(M)
    best = None
    for it in range(6):
        eng.clear_time_series(keep_storage=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.run(M, spm)
        b.record()
        torch.cuda.synchronize()
        if it >= 2:
            best = a.elapsed_time(b) if best is None else min(best, a.elapsed_time(b))
    try:
        ps = eng.pooled_statistics()
    except RuntimeError:                 # a build with the pooled moments switched off
        ps = dict(mean_real=[float("nan")], cov_real=np.full((1, 1), np.nan), observables_mean=[float("nan")] * 4)
    print("  c3  %8.3f ms  %.4e chain-steps/s  (grid %d x block %d)  acceptance %.4f  pooled mean x0 %.6f var x0 %.6f <|c0|> %.6f" % (
        best, n * M * spm / best * 1e3, eng._grid, eng._block, eng.acceptance_rate,
        float(ps["mean_real"][0]), float(ps["cov_real"][0, 0]), float(ps["observables_mean"][3])), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        blocks = sys.argv[1:] or ["64", "128"]
        libs = [None] + sorted(glob.glob(os.path.join(ROOT, "metropolisengine_b200", "lib", "variants", "libme_b200_c3*.so")))
        for lib in libs:
            for blk in blocks:
                env = dict(os.environ)
                if lib:
                    env["ME_B200_LIB"] = lib
                env["ME_BLOCK"] = blk
                print("library:", os.path.basename(lib) if lib else "libme_b200.so (shipped)", "ME_BLOCK", blk, flush=True)
                subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
