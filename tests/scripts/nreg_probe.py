"""Kernel time of the small fused shapes (C1, C2 at their bench sizes) for builds with different register caps
(make SMALL_NREG=... OUT=../lib/variants/libme_b200_r<cap>.so): more resident warps against spills / re-materialised
loop invariants.  Each build runs in its own process (the library is chosen at import time through ME_B200_LIB).

usage: python tests/scripts/nreg_probe.py            (parent: loops over the variants found)
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
    for name, energy, nr, temp, M, spm in (("c2", ("xy_well", 1.0), 2, .1, 2000, 10), ("c1", ("x2",), 1, .01, 10000, 1)):
        n = 65536
        eng = me.MetropolisEngine(energy, initial_real_params=np.zeros(nr), temp=temp, n_chains=n, seed=1,
                                  ts_chunk_bytes=M * (nr + 2) * n * 8)
        eng.reserve_rows(M)
        best = None
        for it in range(5):
            eng.clear_time_series(keep_storage=True)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.run(M, spm)
            b.record()
            torch.cuda.synchronize()
            if it >= 2:
                best = a.elapsed_time(b) if best is None else min(best, a.elapsed_time(b))
        print("  %s  %8.3f ms  %.4e chain-steps/s  (grid %d x block %d)" % (name, best, n * M * spm / best * 1e3, eng._grid, eng._block),
              flush=True)
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        libs = [None] + sorted(glob.glob(os.path.join(ROOT, "metropolisengine_b200", "lib", "variants", "libme_b200_r*.so")))
        for lib in libs:
            env = dict(os.environ)
            if lib:
                env["ME_B200_LIB"] = lib
            print("library:", os.path.basename(lib) if lib else "libme_b200.so (shipped)", flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
