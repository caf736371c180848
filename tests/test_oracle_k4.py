"""CPU-side checks of the shared-covariance oracle (oracle/me_oracle_k4.c, oracle/k4_oracle.py) — the checker the GPU tests
of csrc/me_k4_device.cuh run beside the kernel.  The reference has no shared-covariance mode, so what pins this oracle
is (i) the stream facts below, (ii) exact stationary laws, (iii) the ensemble cross-check against the reference-pinned
per-chain engine in tests/test_gpu_k4.py."""
import numpy as np
from scipy import stats

from oracle import k4_oracle as ko
from oracle.py_port import adaptation_constants


def test_stream_normals_are_standard_symmetric_and_bf16():
    e = ko.K4Ensemble(64, 512, (10, -1, 0.05, 1), 0.1, 2.26, seed=9)
    zb = e.normals(7)
    flat = zb.reshape(-1).astype(np.float64)
    assert stats.kstest(flat[::5], "norm").pvalue > 1e-4
    # the 4096-quantile BF16 table has variance 0.99984 and is cut at 3.84
    assert abs(flat.mean()) < 5 / np.sqrt(flat.size) and abs(flat.var() - 0.99984) < 5 * np.sqrt(2 / flat.size)
    assert np.all((zb.view(np.uint32) & 0xffff) == 0) and np.abs(flat).max() < 3.85           # BF16 values
    assert abs((flat > 0).mean() - 0.5) < 5 * 0.5 / np.sqrt(flat.size)
    assert not np.array_equal(zb, e.normals(8))                             # the step enters the counter
    za, u = e.scalars(7)
    assert np.all((u >= 0) & (u < 1)) and stats.kstest(u, "uniform").pvalue > 1e-4
    assert stats.kstest(za, "norm").pvalue > 1e-4


def test_generator_table_of_the_library_equals_the_independent_one():
    """The library's table (bisection on erfc, me_k4.cu) against scipy's inverse normal CDF: identical BF16 bit patterns."""
    from metropolisengine_b200 import _lib
    import ctypes
    out = (ctypes.c_uint16 * 4096)()
    assert _lib.load().me_k4_normal_table(out, 4096) == 0
    assert np.array_equal(np.frombuffer(out, dtype=np.uint16), ko.normal_table())
    t = (ko.normal_table().astype(np.uint32) << 16).view(np.float32).astype(np.float64)
    assert np.all(np.diff(t) >= 0) and abs((t * t).mean() - 0.99984) < 2e-5


def test_one_rounding_bf16_and_factor_embedding():
    v = np.array([1.0 + 2.0 ** -8, 1.0 + 2.0 ** -8 + 1e-12, 1.0 + 2.0 ** -8 - 1e-12, -0.3, 3.0e-5])
    b = ko.bf16_from_double(v)
    assert list(b[:3]) == [1.0, np.float32(1.0078125), 1.0]                # tie to even, just above, just below
    rng = np.random.default_rng(0)
    a = rng.standard_normal((8, 8)) + 1j * rng.standard_normal((8, 8))
    C = a @ a.conj().T / 8 + 0.5 * np.eye(8)
    B = ko.embed_factor(C)
    emb = B @ B.T                                     # covariance of the embedded increments, interleaved coordinates
    rr, ii, ir = emb[0::2, 0::2], emb[1::2, 1::2], emb[1::2, 0::2]
    assert np.allclose(rr + ii, np.conj(C).real, atol=1e-12)              # CN(0, conj(C)): ME:288-302, SURVEY App. B-8
    assert np.allclose(ir - ir.T, np.conj(C).imag, atol=1e-12)


def test_oracle_ensemble_samples_the_exact_law_of_decoupled_modes():
    """beta = 0 and a stiff amplitude decouple the modes: Re / Im c_q ~ N(0, T / (2 (alpha + gamma q^2))).  A 128-chain
    oracle-only ensemble (its own stream and its own float64 B.z), pooled covariance switching on at the 50th measure."""
    nc, n, T = 8, 128, 0.1
    kappa, alpha, gamma = 500.0, 1.0, 0.05
    ratio = adaptation_constants(1, nc)[2]
    e = ko.K4Ensemble(nc, n, (kappa, alpha, gamma, 0.0), T, ratio, seed=4)
    acc = tot = 0
    samples = []
    for im in range(140):
        for k in range(6):
            e.begin_step_launch()
            zb = e.normals(e.step)
            d, _ = e.delta(zb)
            za, u = e.scalars(e.step)
            a = e.step_injected(d, za, u)
            e.end_step_launch()
            if im >= 70:
                acc += a.sum(); tot += n
        e.measure()
        if im >= 70:
            samples.append(e.state[:, e.L.X + 1:e.L.X + 1 + 2 * nc].copy())
    assert 0.2 < acc / tot < 0.75          # the width is still growing towards the 0.3 target (gain 1/200 per step)
    s = np.concatenate(samples)                                             # [70 * 128, 16]
    q = np.arange(nc) - nc // 2
    exact = np.tile(T / (2.0 * (alpha + gamma * q ** 2)), 2)
    assert np.all(np.abs(s.var(axis=0) / exact - 1.0) < 0.15), s.var(axis=0) / exact
    assert np.all(np.real(np.diag(e.cov_c)) > 0)


def test_shared_covariance_functor_compiles_without_a_gpu():
    from metropolisengine_b200 import _lib
    import ctypes
    src = b"""
__device__ void me_k4_mode(double q, double re, double im, const double* k, double& s0, double& s1) { s0 += re * re + im * im; s1 += q * q * re; }
__device__ double me_k4_total(double a, double s0, double s1, const double* k, int nc) { return k[0] * a * a + s0 + k[1] * s1; }
"""
    buf = ctypes.create_string_buffer(1 << 15)
    L = _lib.load()
    assert L.me_k4_check_energy_source(src, 16, 0, buf, len(buf)) == 0, buf.value.decode()[-400:]
    assert L.me_k4_check_energy_source(b"int x = ;", 16, 0, buf, len(buf)) == _lib.ME_ERR_COMPILE
