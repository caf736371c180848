"""Pins the C oracle (oracle/me_oracle.c) against fixtures recorded from the unmodified reference
(parity level L-A of SURVEY §8c: recorded increments and uniforms are injected), and checks its Philox /
Cholesky proposal path against distributional facts."""
import numpy as np
import pytest

from oracle import c_oracle as co
from tests.conftest import load_golden
from tests.golden.cases import cases, fresh_ctor

CASES = cases()
TOL = 1e-13


def flat_energy(case, n_r, n_c):
    fn = case["energy"]
    if isinstance(fn, dict):                      # dict-of-terms: the total is the sum over the "all" group
        terms = fn["all"]
        fn = lambda r, c: sum(t(r, c) for t in terms.values())

    def f(x):
        r = x[:n_r].copy()
        c = x[n_r:n_r + n_c] + 1j * x[n_r + n_c:]
        return float(np.real(fn(r, c)))
    return f


def make_c_chain(name, use_builtin=True):
    case, g = CASES[name], load_golden(name)
    n_r, n_c = int(g["n_r"]), int(g["n_c"])
    ctor = fresh_ctor(case)
    x0 = np.concatenate([g["x0"], np.real(g["c0"]), np.imag(g["c0"])])
    kw = dict(temp=float(g["temp"]), sampling_width=g["sampling_width0"], x0=x0,
              cov_r=ctor.get("covariance_matrix_real"), cov_c=ctor.get("covariance_matrix_complex"))
    if use_builtin and case["builtin"] is not None:
        bname, consts = case["builtin"]
        ch = co.CChain(n_r, n_c, bname, consts=consts, use_reject=("reject" in case), **kw)
    else:
        rej = None
        if "reject" in case:
            rj = case["reject"]
            rej = lambda x: rj(x[:n_r], x[n_r:n_r + n_c] + 1j * x[n_r + n_c:])
        ch = co.CChain(n_r, n_c, flat_energy(case, n_r, n_c), reject=rej, **kw)
    return ch, g


def rel_close(a, b, tol=TOL):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b), 1e-300)
    return np.all(np.abs(a - b) <= tol * np.maximum(scale, np.max(np.abs(b)) if b.size else 1.0))


@pytest.mark.parametrize("builtin", [True, False])
@pytest.mark.parametrize("name", sorted(CASES))
def test_c_oracle_injected_matches_reference(name, builtin):
    if builtin and CASES[name]["builtin"] is None:
        pytest.skip("no built-in functor for this case")
    ch, g = make_c_chain(name, use_builtin=builtin)
    M, K = int(g["n_measures"]), int(g["steps_per_measure"])
    n_r, n_c, d = ch.n_r, ch.n_c, ch.d
    for im in range(M):
        sl = slice(im * K, (im + 1) * K)
        if "group" in g and g["group"].any():          # group-wise schedule: one step at a time, then the measure
            acc = []
            for s in range(im * K, (im + 1) * K):
                a, _ = ch.run(1, 1, False, delta=g["delta"][s:s + 1], u=g["u"][s:s + 1], group=int(g["group"][s]))
                acc.append(a[0])
            _, ts = ch.run(1, 0, True, delta=g["delta"][:1], u=g["u"][:1], want_ts=True)
            acc = np.array(acc)
        else:
            acc, ts = ch.run(1, K, True, delta=g["delta"][sl], u=g["u"][sl], want_ts=True)
        assert np.array_equal(acc, g["accept"][sl]), "decisions differ in block %d" % im
        assert rel_close(ch.x, g["step_x"][(im + 1) * K - 1])
        assert rel_close(ch.energy, g["m_energy"][im])
        assert rel_close(ts[0, :d], ch.x, 0)
        if n_r:
            assert rel_close(ch.sigma[0], g["m_sigma_r"][im])
            assert rel_close(ch.mean[:n_r], g["m_real_mean"][im])
            assert rel_close(ch.cov_r, g["m_cov_r"][im])
        if n_c:
            assert rel_close(ch.sigma[1], g["m_sigma_c"][im])
            cm = ch.mean[n_r:n_r + n_c] + 1j * ch.mean[n_r + n_c:]
            assert rel_close(cm, g["m_complex_mean"][im])
            assert rel_close(ch.cov_c, g["m_cov_c"][im])
        assert rel_close(ch.obs_mean, g["m_obs_mean"][im])
    assert ch.n_measure.value == int(g["measure_step_counter"])


@pytest.mark.parametrize("name", ["kat1_x2", "kat2_xy", "c3_3r4c", "pure_2c", "warm_3r2c"])
def test_c_oracle_is_bit_identical_where_arithmetic_coincides(name):
    """With energies written as x*x (no libm pow) the restatement reproduces the reference bit-for-bit for the
    parameters, sigma, means and covariance matrices (observable means may differ by the pow-vs-multiply ulp,
    SURVEY App. B-16)."""
    ch, g = make_c_chain(name, use_builtin=CASES[name]["builtin"] is not None)
    M, K = int(g["n_measures"]), int(g["steps_per_measure"])
    ch.run(M, K, True, delta=g["delta"], u=g["u"])
    n_r, n_c = ch.n_r, ch.n_c
    if CASES[name]["builtin"] is not None:
        assert np.array_equal(ch.x, g["step_x"][-1])
    if n_r and CASES[name]["builtin"] is not None:
        assert ch.sigma[0] == g["m_sigma_r"][-1]
        assert np.array_equal(ch.mean[:n_r], g["m_real_mean"][-1])
        assert np.array_equal(ch.cov_r, g["m_cov_r"][-1])
    if n_c:
        # numpy's SIMD complex multiply uses fused multiply-adds (the reference's own diagonal carries a
        # ~1e-20 imaginary residue), so the complex block agrees to rounding, not bit-for-bit
        assert rel_close(ch.cov_c, g["m_cov_c"][-1], 1e-14)


def test_philox_known_answer():
    """Philox4x32 known-answer vectors from the Random123 distribution (kat_vectors): zero and all-ones
    counter/key and the pi-digits vector, for the 10-round default and for the 7 rounds of this stream definition."""
    import ctypes
    out = (ctypes.c_uint32 * 4)()
    L = co.lib()

    def run(seed, chain, step, slot, rounds):
        L.meo_philox_rounds(ctypes.c_uint64(seed), ctypes.c_uint64(chain), ctypes.c_uint32(step), ctypes.c_uint32(slot),
                            ctypes.c_int(rounds), out)
        return [hex(v) for v in out]

    ones64, ones32 = 0xffffffffffffffff, 0xffffffff
    pi = (0x299f31d0a4093822, 0x85a308d3243f6a88, 0x13198a2e, 0x03707344)
    assert run(0, 0, 0, 0, 10) == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    assert run(ones64, ones64, ones32, ones32, 10) == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    assert run(*pi, 10) == ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']
    assert run(0, 0, 0, 0, 7) == ['0x5f6fb709', '0xd893f64', '0x4f121f81', '0x4f730a48']
    assert run(ones64, ones64, ones32, ones32, 7) == ['0x5207ddc2', '0x45165e59', '0x4d8ee751', '0x8c52f662']
    assert run(*pi, 7) == ['0x4dfccaba', '0x190a87f0', '0xc47362ba', '0xb6b5242a']
    # the stream itself uses 7 rounds
    L.meo_philox(ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint32(0), ctypes.c_uint32(0), out)
    assert [hex(v) for v in out] == ['0x5f6fb709', '0xd893f64', '0x4f121f81', '0x4f730a48']


def test_philox_normals_are_standard():
    import ctypes
    L = co.lib()
    z0, z1 = ctypes.c_double(), ctypes.c_double()
    zs = []
    for chain in range(20000):
        L.meo_normal_pair(ctypes.c_uint64(7), ctypes.c_uint64(chain), ctypes.c_uint32(3), ctypes.c_uint32(0),
                          ctypes.byref(z0), ctypes.byref(z1))
        zs.append((z0.value, z1.value))
    zs = np.array(zs)
    from scipy import stats
    assert stats.kstest(zs[:, 0], "norm").pvalue > 1e-3
    assert stats.kstest(zs[:, 1], "norm").pvalue > 1e-3
    assert abs(np.corrcoef(zs.T)[0, 1]) < 0.03


def test_cholesky_factors_reproduce_covariances():
    ch, g = make_c_chain("warm_3r2c", use_builtin=False)
    Lr, Lc = ch.fac_r, ch.fac_c
    assert np.allclose(Lr @ Lr.T, ch.cov_r, rtol=1e-14, atol=1e-15)
    assert np.allclose(Lc @ Lc.conj().T, ch.cov_c, rtol=1e-14, atol=1e-15)


def test_philox_proposals_have_reference_covariance():
    """Increments of the Philox/Cholesky path must have the distribution the reference samples:
    real block N(0, sigma^2 C_r); complex block CN(0, sigma^2 conj(C_c)), pseudo-covariance 0
    (metropolis_engine.py:288-302)."""
    case = CASES["warm_3r2c"]
    ctor = fresh_ctor(case)
    Cr, Cc, sig = ctor["covariance_matrix_real"], ctor["covariance_matrix_complex"], ctor["sampling_width"]
    n_r, n_c = 3, 2
    x0 = np.zeros(7)
    N = 40000
    inc = np.zeros((N, 7))
    for chain in range(N):
        ch = co.CChain(n_r, n_c, lambda x: 0.0, temp=1.0, sampling_width=sig, x0=x0, cov_r=Cr, cov_c=Cc)
        ch.run(1, 1, False, seed=11, chain_id=chain)
        inc[chain] = ch.x          # flat energy: every proposal is accepted (diff == 0)
    r = inc[:, :3]
    w = inc[:, 3:5] + 1j * inc[:, 5:]
    emp_r = r.T @ r / N
    emp_c = w.T @ w.conj() / N            # E[w w^H]
    pseudo = w.T @ w / N
    se = 4.0 / np.sqrt(N)
    assert np.allclose(emp_r, sig ** 2 * Cr, atol=se * sig ** 2)
    assert np.allclose(emp_c, sig ** 2 * np.conj(Cc), atol=se * sig ** 2)
    assert np.allclose(pseudo, 0, atol=se * sig ** 2)


@pytest.mark.parametrize("re0, im0", [(0.0, 0.0), (-0.0, 0.0), (-0.0, -0.0), (0.0, -0.0)])
def test_magnitude_move_from_zero_modulus_follows_cmath_polar(re0, im0):
    """The reference's magnitude move is cmath.rect(gauss(|c|, s), phase(c)) (ME:304-310).  At zero modulus the phase
    is atan2 of signed zeros (arg(-0) = +-pi), which a rejected first move plus a phase redraw does produce; the C
    oracle (and through it the kernels) must give what cmath gives."""
    import cmath
    import ctypes
    import math
    o = co.CChain(0, 1, lambda x: 0.0, temp=.1, sampling_width=0.5, x0=np.array([re0, im0]))
    lay = co.layout(0, 1)
    o.state[lay.X], o.state[lay.X + 1] = re0, im0          # np.array keeps the sign of zero; make sure of it
    z0, z1 = ctypes.c_double(), ctypes.c_double()
    co.lib().meo_normal_pair(ctypes.c_uint64(3), ctypes.c_uint64(11), ctypes.c_uint32(0), ctypes.c_uint32(0),
                             ctypes.byref(z0), ctypes.byref(z1))
    acc, _ = o.run(1, 1, False, seed=3, chain_id=11, step0=0, group=3)
    assert acc.sum() == 1                                   # flat energy: always accepted
    modulus, phase = cmath.polar(complex(re0, im0))
    want = cmath.rect(modulus + z0.value * (0.5 * 0.5 * 1.0), phase)      # s = sigma^2 C_jj, C = identity
    got = complex(o.state[lay.X], o.state[lay.X + 1])
    assert got.real == pytest.approx(want.real, rel=1e-15, abs=0) and math.copysign(1, got.real) == math.copysign(1, want.real)
    assert got.imag == pytest.approx(want.imag, rel=1e-12, abs=1e-300)
