"""Pins the numpy port (oracle/py_port.py) against fixtures produced by the unmodified reference.

Two modes per case:
  * seeded  — the port draws from the same seeded global generators the reference used
              (np.random.seed / random.seed): every decision and every number must be bit-identical;
  * injected — the recorded increments/uniforms are replayed (parity level L-A of SURVEY §8c).
"""
import random

import numpy as np
import pytest

from oracle.py_port import InjectedDraws, PortChain, adaptation_constants, statistical_inefficiency
from tests.conftest import load_golden
from tests.golden.cases import cases, fresh_ctor

CASES = cases()


def _drive(chain, g):
    M, K = int(g["n_measures"]), int(g["steps_per_measure"])
    n_r, n_c = int(g["n_r"]), int(g["n_c"])
    s = 0
    for im in range(M):
        for _ in range(K):
            grp = int(g["group"][s]) if "group" in g else 0
            a = chain.step({0: None, 1: "real", 2: "complex", 3: "magnitude", 4: "phase"}[grp])
            assert bool(a) == bool(g["accept"][s]), "decision differs at step %d" % s
            got = np.concatenate([chain.real_params, np.real(chain.complex_params), np.imag(chain.complex_params)])
            assert np.array_equal(got, g["step_x"][s]), "state differs at step %d" % s
            assert chain.live_energy() == g["step_energy"][s]
            s += 1
        chain.measure()
        if n_r:
            assert np.array_equal(chain.real_mean, g["m_real_mean"][im])
            assert np.array_equal(chain.covariance_matrix_real, g["m_cov_r"][im])
            assert chain.real_group_sampling_width == g["m_sigma_r"][im]
        if n_c:
            assert np.array_equal(chain.complex_mean, g["m_complex_mean"][im])
            assert np.array_equal(chain.covariance_matrix_complex, g["m_cov_c"][im])
            assert chain.complex_group_sampling_width == g["m_sigma_c"][im]
        assert np.array_equal(chain.observables_mean, g["m_obs_mean"][im])
    assert chain.measure_step_counter == int(g["measure_step_counter"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_seeded_is_bit_identical_to_reference(name):
    case, g = CASES[name], load_golden(name)
    np.random.seed(case["seed"])
    random.seed(case["seed"])
    chain = PortChain(case["energy"], reject_condition=case.get("reject"), **fresh_ctor(case))
    assert chain.ratio == float(g["ratio"]) and chain.alpha == float(g["alpha"])
    _drive(chain, g)


@pytest.mark.parametrize("name", sorted(CASES))
def test_port_injected_draws_reproduce_reference(name):
    case, g = CASES[name], load_golden(name)
    draws = InjectedDraws(g["delta"], g["u"], int(g["n_r"]), int(g["n_c"]))
    chain = PortChain(case["energy"], reject_condition=case.get("reject"), draws=draws, **fresh_ctor(case))
    _drive(chain, g)


def test_constants_match_survey_kat():
    # SURVEY.md §4 "constants" row (values produced by the live reference)
    alpha, m, ratio = adaptation_constants(1, 0)
    assert alpha == 1.0364333894937898
    for (n_r, n_c), want in (((1, 0), 4.761904761904762), ((2, 0), 3.4922480938910487),
                             ((2, 1), 3.0690292045531455), ((3, 4), 2.585350473881254),
                             ((1, 64), 2.261657784893143)):
        assert adaptation_constants(n_r, n_c)[2] == want


def test_statistical_inefficiency_ar1():
    rng = np.random.default_rng(0)
    phi = 0.8
    x = np.zeros(200000)
    e = rng.standard_normal(x.size)
    for i in range(1, x.size):
        x[i] = phi * x[i - 1] + e[i]
    g = statistical_inefficiency(x)
    assert abs(g - (1 + phi) / (1 - phi)) < 0.8   # exact value 9


def test_equilibration_detection_on_known_transient():
    """oracle/py_port.py::detect_equilibration (restated pymbar algorithm, parity unpinned — see its header):
    AR(1) noise plus an exponential transient of 5 correlation-free time constants must be cut near its end, and
    the inefficiency of the production region must be the AR(1) value."""
    from oracle.py_port import detect_equilibration, pymbar_statistical_inefficiency
    rng = np.random.default_rng(1)
    phi, n = 0.7, 3000
    x = np.zeros(n)
    e = rng.standard_normal(n)
    for i in range(1, n):
        x[i] = phi * x[i - 1] + e[i]
    assert abs(pymbar_statistical_inefficiency(x) - (1 + phi) / (1 - phi)) < 0.5
    y = x.copy()
    y[:400] += 8 * np.exp(-np.arange(400) / 60.)
    t0, g, neff = detect_equilibration(y, nskip=10)
    assert 60 <= t0 <= 400 and abs(g - (1 + phi) / (1 - phi)) < 1.5 and neff > 300
    assert detect_equilibration(np.ones(50)) == (0, 1.0, 1.0)
    t0_flat, _g, _n = detect_equilibration(x, nskip=10)
    assert t0_flat < 300                                        # no transient: (almost) nothing is discarded
