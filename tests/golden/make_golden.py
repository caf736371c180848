#!/usr/bin/env python
"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE.

Only runs in the build container (needs /root/reference).  The GPU box never runs this; it reads
the committed ``*.npz`` files.  Usage::

    python tests/golden/make_golden.py            # all single-chain cases
    python tests/golden/make_golden.py --ensembles  # also the M-chain distribution fixtures (slow)

Recipe:
  * stub ``matplotlib``/``pymbar`` (imported by metropolisengine/statistics.py:2,4, unused on the hot path)
  * seed both global generators (``np.random.seed(s); random.seed(s)``)
  * tap ``np.random.multivariate_normal`` (metropolis_engine.py:268,300) and ``random.uniform``
    (metropolis_engine.py:335) to record, per step, the proposal increment ``a = z @ (sqrt(s)[:,None]*v)``
    (so that ``proposal == a + mean`` bit-for-bit), the underlying standard normals ``z`` and the
    accept uniform ``u`` (NaN when the reference did not draw one).

Array layout in the fixtures: ``d = n_r + 2*n_c`` with order ``[real..., Re c..., Im c...]`` —
the reference's own embedding order (metropolis_engine.py:288).
"""
import argparse
import contextlib
import io
import os
import random
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "pymbar", "pymbar.timeseries"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pymbar"].timeseries = sys.modules["pymbar.timeseries"]
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    import metropolisengine as me  # noqa: E402  (the reference package)
    assert me.__file__.startswith(REFERENCE), me.__file__
    return me


class Tap:
    """Records every Gaussian draw and accept-uniform the reference consumes."""

    def __init__(self):
        self.draws = []   # list of (increment a, standard normals z)
        self.uniforms = []
        self._orig_mvn = np.random.multivariate_normal
        self._orig_uniform = random.uniform

    def __enter__(self):
        tap = self

        def mvn(mean, cov, *args, **kwargs):
            mean = np.asarray(mean)
            state0 = np.random.get_state()
            out = tap._orig_mvn(mean, cov, *args, **kwargs)
            state1 = np.random.get_state()
            np.random.set_state(state0)
            z = np.random.standard_normal(mean.shape[0])
            # the legacy sampler consumes exactly len(mean) normals
            s1 = np.random.get_state()
            assert s1[2] == state1[2] and np.array_equal(s1[1], state1[1])
            np.random.set_state(state1)
            (u, s, v) = np.linalg.svd(np.asarray(cov, dtype=np.float64))
            a = np.dot(z, np.sqrt(s)[:, None] * v)
            assert np.array_equal(a + mean, out), "numpy sampler restatement is not bit-exact"
            tap.draws.append((a.copy(), z.copy()))
            return out

        def uniform(lo, hi):
            val = tap._orig_uniform(lo, hi)
            if lo == 0:                     # accept test (metropolis_engine.py:335); the phase redraw (ME:317)
                tap.uniforms.append(val)    # calls uniform(-pi, pi) and is recorded through its proposal instead
            return val

        np.random.multivariate_normal = mvn
        random.uniform = uniform
        return self

    def __exit__(self, *exc):
        np.random.multivariate_normal = self._orig_mvn
        random.uniform = self._orig_uniform


def run_case(me, name, energy, n_measures, steps_per_measure, seed=0, reject=None, df=True, schedule="all", **ctor):
    """Drive the reference exactly like the README/demo loops and record everything."""
    np.random.seed(seed)
    random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        eng = me.MetropolisEngine(energy, **ctor)
    if reject is not None:
        eng.set_reject_condition(reject)
    n_r, n_c = eng.num_real_params, eng.num_complex_params
    d = n_r + 2 * n_c
    S = n_measures * steps_per_measure
    rec = dict(
        n_r=n_r, n_c=n_c, n_measures=n_measures, steps_per_measure=steps_per_measure, seed=seed,
        temp=float(eng.temp), target_acceptance=float(eng.target_acceptance),
        sampling_width0=np.asarray(ctor.get("sampling_width", 0.05), dtype=np.float64),   # scalar, or [sigma_r, sigma_c]
        alpha=float(eng.alpha), ratio=float(eng.ratio), m=int(eng.m),
        x0=np.array(eng.real_params, dtype=np.float64),
        c0=np.array(eng.complex_params, dtype=np.complex128),
        cov_r0=np.array(eng.covariance_matrix_real, dtype=np.float64) if n_r else np.zeros((0, 0)),
        cov_c0=np.array(eng.covariance_matrix_complex, dtype=np.complex128) if n_c else np.zeros((0, 0), complex),
        energy0=float(np.real(eng.energy_total)),
    )
    delta = np.zeros((S, d))
    zs = np.zeros((S, d))
    us = np.full(S, np.nan)
    acc = np.zeros(S, dtype=np.bool_)
    group = np.zeros(S, dtype=np.int32)      # 0 = step_all, 1 = step_real_group, 2 = step_complex_group,
                                             # 3 / 4 = magnitude / phase half of the magnitude-phase complex move
    step_x = np.zeros((S, d))
    step_sig = np.zeros((S, 3))
    step_energy = np.zeros(S)
    snap = {k: [] for k in ("x", "c", "sigma", "sigma_r", "sigma_c", "energy", "real_mean", "complex_mean",
                            "cov_r", "cov_c", "obs_mean", "obs")}
    mixed = n_r > 0 and n_c > 0

    def live_energy():
        # all-real / all-complex engines and group steps keep the live energy in eng.energy;
        # the mixed step_all keeps it in eng.energy_total (metropolis_engine.py:255)
        if mixed and schedule not in ("groups", "magphase"):
            return float(np.real(eng.energy_total))
        return float(np.real(sum(eng.energy.values())))

    with Tap() as tap:
        s = 0
        for im in range(n_measures):
            for _ in range(steps_per_measure):
                nd, nu = len(tap.draws), len(tap.uniforms)
                if schedule == "magphase":     # real group, then the two halves of the magnitude-phase complex move
                    # (metropolis_engine.py:168-207; step_complex_group is rebound to magnitude + phase, ME:129-130)
                    group[s] = (1, 3, 4)[s % 3]
                    if group[s] == 1:
                        a = eng.step_real_group()
                        (inc, z), = tap.draws[nd:]
                        delta[s, :n_r] = inc
                        zs[s, :n_r] = z
                    else:
                        name_ = "draw_complex_magnitudes" if group[s] == 3 else "draw_complex_phases"
                        orig_draw, seen = getattr(eng, name_), []
                        setattr(eng, name_, lambda o=orig_draw, seen=seen: seen.append(o()) or seen[-1])
                        with warnings.catch_warnings():
                            warnings.simplefilter("ignore")        # ComplexWarning at ME:310
                            a = (eng.step_complex_group_magnitude() if group[s] == 3
                                 else eng.step_complex_group_phase())
                        setattr(eng, name_, orig_draw)
                        (prop,) = seen
                        delta[s, n_r:n_r + n_c] = np.real(prop)     # ABSOLUTE proposed values, not increments
                        delta[s, n_r + n_c:] = np.imag(prop)
                elif schedule == "groups":     # alternate the two group steps (how the cylinder app drives it)
                    group[s] = 1 + (s % 2)
                    a = eng.step_real_group() if group[s] == 1 else eng.step_complex_group()
                    (inc, z), = tap.draws[nd:]
                    sl = slice(0, n_r) if group[s] == 1 else slice(n_r, d)
                    delta[s, sl] = inc
                    zs[s, sl] = z
                else:
                    a = eng.step_all()
                    new = tap.draws[nd:]
                    delta[s] = np.concatenate([inc for (inc, _z) in new])
                    zs[s] = np.concatenate([z for (_inc, z) in new])
                if len(tap.uniforms) > nu:
                    assert len(tap.uniforms) == nu + 1
                    us[s] = tap.uniforms[-1]
                acc[s] = bool(a)
                step_x[s, :n_r] = eng.real_params
                step_x[s, n_r:n_r + n_c] = np.real(eng.complex_params)
                step_x[s, n_r + n_c:] = np.imag(eng.complex_params)
                step_sig[s] = (getattr(eng, "sampling_width", np.nan), eng.real_group_sampling_width,
                               eng.complex_group_sampling_width)
                step_energy[s] = live_energy()
                s += 1
            eng.measure()
            snap["x"].append(np.array(eng.real_params, dtype=np.float64))
            snap["c"].append(np.array(eng.complex_params, dtype=np.complex128))
            snap["sigma"].append(float(getattr(eng, "sampling_width", np.nan)))
            snap["sigma_r"].append(float(eng.real_group_sampling_width))
            snap["sigma_c"].append(float(eng.complex_group_sampling_width))
            snap["energy"].append(live_energy())
            snap["real_mean"].append(np.array(eng.real_mean, dtype=np.float64))
            snap["complex_mean"].append(np.array(eng.complex_mean, dtype=np.complex128))
            snap["cov_r"].append(np.array(eng.covariance_matrix_real, dtype=np.float64) if n_r else np.zeros((0, 0)))
            snap["cov_c"].append(np.array(eng.covariance_matrix_complex, dtype=np.complex128) if n_c
                                 else np.zeros((0, 0), complex))
            snap["obs_mean"].append(np.array(eng.observables_mean, dtype=np.float64))
            snap["obs"].append(np.array(eng.observables, dtype=np.float64))
    rec.update(group=group, delta=delta, z=zs, u=us, accept=acc, step_x=step_x, step_sigma=step_sig, step_energy=step_energy)
    for k, v in snap.items():
        rec["m_" + k] = np.array(v)
    rec["measure_step_counter"] = int(eng.measure_step_counter)
    rec["observables_names"] = np.array(eng.observables_names)
    rec["params_names"] = np.array(eng.params_names)
    if df:
        with contextlib.redirect_stdout(io.StringIO()):
            eng.save_time_series()
        rec["df_columns"] = np.array(list(eng.df.columns))
        for i, col in enumerate(eng.df.columns):
            vals = eng.df[col].to_numpy()
            if np.iscomplexobj(vals):
                rec["df_%d" % i] = vals.astype(np.complex128)
            else:
                rec["df_%d" % i] = np.real(vals).astype(np.float64)
    out = os.path.join(HERE, name + ".npz")
    np.savez_compressed(out, **rec)
    print("%-22s steps=%6d accepts=%6d  -> %s (%.0f KiB)" % (name, S, int(acc.sum()), os.path.basename(out),
                                                            os.path.getsize(out) / 1024))
    return eng, rec


def check_survey_kats(me):
    """Reproduce SURVEY.md §4 KAT1-3 (energies written with ``**2`` as in the README/demos) to prove
    this harness drives the reference the same way the survey did."""
    np.random.seed(0); random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        e = me.MetropolisEngine(lambda r, c: r[0] ** 2, initial_real_params=[0.0], temp=.01)
    acc = 0
    for _ in range(1000):
        acc += bool(e.step_all()); e.measure()
    assert acc == 441, acc
    assert e.real_params[0] == 0.020564709191852493
    assert e.real_group_sampling_width == 0.560479460029389
    assert e.real_mean[0] == 0.002385584461589615
    assert e.covariance_matrix_real[0, 0] == 0.2260946623964727
    assert e.sampling_width == 0.05
    np.random.seed(0); random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        e = me.MetropolisEngine(lambda r, c: r[0] ** 2 + r[1] ** 2, initial_real_params=np.array([0., 0.]), temp=.1)
    acc = 0
    for _ in range(1000):
        for _ in range(10):
            acc += bool(e.step_all())
        e.measure()
    assert acc == 3167, acc
    assert e.real_group_sampling_width == 0.6737107772882164
    assert e.covariance_matrix_real[0, 0] == 0.479512502307138

    def e3(r, c):
        a = (c[0] * c[0].conjugate()).real
        return (1 - r[0]) ** 2 + (1 - r[1]) ** 2 + r[0] * r[1] * (-1 * a + .5 * a * a)
    np.random.seed(0); random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        e = me.MetropolisEngine(e3, initial_real_params=np.array([0., 0.]), initial_complex_params=np.array([0j]),
                                temp=.1)
    acc = 0
    for _ in range(100):
        for _ in range(10):
            acc += bool(e.step_all())
        e.measure()
    assert acc == 422, acc
    assert e.sampling_width == 0.3153914717010739
    assert e.covariance_matrix_complex[0, 0].real == 1.3303170224979906
    for m, want in ((1, 4.761904761904762), (2, 3.4922480938910487), (7, 2.585350473881254)):
        with contextlib.redirect_stdout(io.StringIO()):
            e = me.MetropolisEngine(lambda r, c: 0.0, initial_real_params=[0.0] * m, temp=.1)
        assert e.ratio == want, (m, e.ratio)
    print("SURVEY KAT1-3 + constants reproduced bit-for-bit by the live reference")


def single_chain_cases(me, only=None):
    from tests.golden.cases import cases, fresh_ctor
    for name, case in cases().items():
        if only and name not in only:
            continue
        run_case(me, name, case["energy"], case["n_measures"], case["steps_per_measure"], seed=case["seed"],
                 reject=case.get("reject"), schedule=case.get("schedule", "all"), **fresh_ctor(case))


def _ensemble_worker(args):
    """One reference process = one chain (global RNG state forbids sharing a process)."""
    cfg, seed = args
    me = import_reference()
    from oracle import energies as en
    np.random.seed(seed); random.seed(seed)
    if cfg == "c1":
        kw = dict(initial_real_params=[0.0], temp=.01); fn = en.x2; M, K = 1000, 1
    elif cfg == "c2":
        kw = dict(initial_real_params=np.array([0., 0.]), temp=.1); fn = en.xy_well; M, K = 1000, 10
    else:                 # "c3": the demo-style energy; "c3b": its bounded |x0 x1| variant — the one bench.py runs
        kw = dict(initial_real_params=np.zeros(3), initial_complex_params=np.zeros(4, dtype=complex), temp=.1)
        fn = en.mixed_3r4c_bounded if cfg == "c3b" else en.mixed_3r4c; M, K = 300, 10
    with contextlib.redirect_stdout(io.StringIO()):
        e = me.MetropolisEngine(fn, **kw)
    acc = np.zeros(M * K, dtype=np.bool_)
    s = 0
    for _ in range(M):
        for _ in range(K):
            acc[s] = bool(e.step_all()); s += 1
        e.measure()
    n_r, n_c = e.num_real_params, e.num_complex_params
    sig = e.sampling_width if (n_r and n_c) else e.real_group_sampling_width
    row = [sig, acc.mean(), acc[len(acc) // 2:].mean()]
    row += list(e.real_params) + list(e.real_mean) + list(np.diag(e.covariance_matrix_real))
    if n_c:
        row += list(np.abs(e.complex_params)) + list(np.real(np.diag(e.covariance_matrix_complex)))
    row += list(e.observables_mean)
    return row


def ensembles(only=None):
    """Per-chain end-of-run quantities of M independent reference chains (seed = chain index); the GPU
    ensemble tests compare distributions against these (two-sample KS, z-tests).  SURVEY.md §4."""
    import multiprocessing as mp
    for cfg, M in (("c1", 512), ("c2", 256), ("c3", 96), ("c3b", 192)):
        if only and cfg not in only:
            continue
        with mp.get_context("fork").Pool(os.cpu_count()) as pool:
            rows = pool.map(_ensemble_worker, [(cfg, s) for s in range(M)], chunksize=4)
        rows = np.array(rows, dtype=np.float64)
        if cfg == "c1":
            cols = ["sigma", "acc_all", "acc_half2", "x0", "mean0", "cov00", "obs_abs0", "obs_sq0"]
        elif cfg == "c2":
            cols = ["sigma", "acc_all", "acc_half2", "x0", "x1", "mean0", "mean1", "cov00", "cov11",
                    "obs_abs0", "obs_abs1", "obs_sq0", "obs_sq1"]
        else:
            cols = (["sigma", "acc_all", "acc_half2"] + ["x%d" % i for i in range(3)] + ["mean%d" % i for i in range(3)]
                    + ["cov%d%d" % (i, i) for i in range(3)] + ["absc%d" % j for j in range(4)]
                    + ["covc%d%d" % (j, j) for j in range(4)] + ["obs_abs%d" % i for i in range(7)]
                    + ["obs_sq%d" % i for i in range(3)])
        assert rows.shape[1] == len(cols), (rows.shape, len(cols))
        out = os.path.join(HERE, "ensemble_%s.npz" % cfg)
        np.savez_compressed(out, rows=rows, columns=np.array(cols), M=M)
        print("ensemble %s: M=%d  sigma=%.5f  acc2=%.5f  -> %s" % (cfg, M, rows[:, 0].mean(), rows[:, 2].mean(),
                                                                os.path.basename(out)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ensembles", action="store_true")
    ap.add_argument("--skip-single", action="store_true")
    ap.add_argument("--only", nargs="*", help="regenerate only these single-chain cases")
    a = ap.parse_args()
    ref = import_reference()
    if not a.skip_single:
        check_survey_kats(ref)
        single_chain_cases(ref, a.only)
    if a.ensembles:
        ensembles(a.only)
