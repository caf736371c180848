"""Debug probe: first divergence between the runtime-shape magnitude/phase path and the C oracle (chain 0)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import metropolisengine_b200 as me
from oracle import c_oracle as co

consts = [10.0, -1.0, 0.05, 1.0]
c0 = np.concatenate([np.full(8, 0.05 + 0.02j), np.zeros(8, dtype=complex)])
eng = me.MetropolisEngine(me.BuiltinEnergy("cylinder", *consts, reject=True), initial_real_params=np.array([0.2]),
                          initial_complex_params=c0, temp=.1, sampling_width=0.3, n_chains=48, seed=31,
                          complex_sample_method="magnitude-phase")
lay = eng._lay
o = co.CChain(1, 16, "cylinder", consts=consts, temp=.1, sampling_width=0.3,
              x0=np.concatenate([[0.2], c0.real, c0.imag]), use_reject=True)
names = ["X", "E", "SIG", "MEAN", "COVR", "COVC", "OBSM", "FACR", "FACC", "NACC"]
offs = [getattr(lay, n) for n in names] + [lay.WORDS]


def report(tag):
    st = eng.state.cpu().numpy()[:, 0]
    out = []
    for n, a, b in zip(names, offs[:-1], offs[1:]):
        dd = np.abs(st[a:b] - o.state[a:b])
        if dd.size and dd.max() > 1e-12:
            k = int(dd.argmax())
            out.append("%s[%d] %.3e (gpu %.12g oracle %.12g)" % (n, k, dd.max(), st[a + k], o.state[a + k]))
    if out:
        print(tag, "; ".join(out), flush=True)
    return bool(out)


step, shown = 0, 0
for im in range(54):
    for grp in (1, 3, 4, 3):
        if grp == 1:
            eng.step_real_group()
        elif grp == 3:
            eng.step_complex_group_magnitude()
        else:
            eng.step_complex_group_phase()
        o.run(1, 1, False, seed=31, chain_id=0, step0=step, group=grp)
        step += 1
        if shown < 12 and report("measure %d step %d group %d:" % (im, step - 1, grp)):
            shown += 1
    eng.measure()
    o.run(1, 0, True, seed=31, chain_id=0, step0=step)
    if shown < 12 and report("after measure %d:" % im):
        shown += 1
report("final:")
print("done")
