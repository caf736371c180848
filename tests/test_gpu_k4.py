"""Shared-covariance tensor-core path (BASELINE config 4 shape: 1 real + 64 complex), csrc/me_k4.cu.

The reference has no shared-covariance mode (it is one chain), so parity here is: (i) the C restatement of this path
(oracle/me_oracle_k4.c + oracle/k4_oracle.py) run beside the kernel on a whole 256-chain ensemble: generator stream,
tensor-core increments (tolerance, then injected), identical accept decisions, bit-identical FP64 state, pooled
covariance, Cholesky factor and its one-measure lag; (ii) an ensemble cross-check against the per-chain engine (pinned by
goldens recorded from the live reference) on the full quartic energy; (iii) the tensor-core contraction equals B.z computed in
float64 from the same BF16 operands; the in-kernel normals are standard; ensembles sample the exact stationary law of
decoupled Gaussian modes; virial identity; hard wall, adaptation target, sharding invariance; user CUDA functors."""
import numpy as np
import pytest
import torch
from scipy import stats

pytestmark = pytest.mark.gpu


def _engine(**kw):
    import metropolisengine_b200 as me
    return me.SharedCovarianceEngine(**kw)


def _rand_cov(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    c = a @ a.conj().T / n + 0.5 * np.eye(n)
    return c


@pytest.mark.parametrize("with_cov", [False, True])
def test_tensor_core_increments_match_float64_matmul(with_cov):
    n = 256
    kw = dict(energy_consts=(10.0, 1.0, 0.05, 0.0), temp=.1, n_chains=n, seed=3)
    if with_cov:
        kw["covariance_matrix_complex"] = _rand_cov(64, 1)
    eng = _engine(**kw)
    dz = torch.zeros((128, n), dtype=torch.float32, device="cuda")
    dd = torch.zeros((128, n), dtype=torch.float32, device="cuda")
    eng.step(1, _dbg=(dz, dd))
    torch.cuda.synchronize()
    z = dz.double().cpu().numpy()                     # BF16 values of the normals, [k][chain]
    B = eng._B.to(torch.bfloat16).double().cpu().numpy()   # BF16 values of the factor, [n][k]
    want = B @ z
    got = dd.double().cpu().numpy()
    scale = np.abs(B) @ np.abs(z)
    assert np.all(np.abs(got - want) <= 2e-5 * scale + 1e-30)
    if with_cov:
        # the factor reproduces the reference's proposal law CN(0, conj(C)): E[w w^H] = conj(C) (ME:288-302)
        Bf = eng._B.cpu().numpy()
        emb = Bf @ Bf.T                                 # covariance of the embedded increments, interleaved coords
        C = kw["covariance_matrix_complex"]
        rr, ii, ir = emb[0::2, 0::2], emb[1::2, 1::2], emb[1::2, 0::2]
        assert np.allclose(rr + ii, np.conj(C).real, atol=1e-12)
        assert np.allclose(ir - ir.T, np.conj(C).imag, atol=1e-12)


def test_in_kernel_normals_are_standard_and_symmetric():
    n = 4096
    eng = _engine(energy_consts=(10.0, 1.0, 0.05, 0.0), temp=.1, n_chains=n, seed=11)
    dz = torch.zeros((128, n), dtype=torch.float32, device="cuda")
    dd = torch.zeros((128, n), dtype=torch.float32, device="cuda")
    eng.step(1, _dbg=(dz, dd))
    z = dz.double().cpu().numpy()
    flat = z.reshape(-1)
    assert abs(flat.mean()) < 5 / np.sqrt(flat.size)
    assert abs(flat.var() - 0.99984) < 5 * np.sqrt(2.0 / flat.size) + 2e-5        # variance of the 4096-quantile BF16 table
    assert stats.kstest(flat[::7], "norm").pvalue > 1e-4          # BF16 rounding is far below the KS resolution here
    cm = np.corrcoef(z[:16])                                        # different coordinates are uncorrelated
    assert np.max(np.abs(cm - np.eye(16))) < 6 / np.sqrt(n)
    assert abs((flat > 0).mean() - 0.5) < 5 * 0.5 / np.sqrt(flat.size)


def test_gaussian_modes_reach_exact_stationary_variances():
    """beta = 0 and a stiff amplitude decouple the modes: Re/Im c_q ~ N(0, T / (2 (alpha + gamma q^2 (1+a^2)))),
    a ~ 0.  Pooled variances over 8192 chains after burn-in, z-test with sd sqrt(2/N_eff)."""
    n, T = 8192, 0.1
    kappa, alpha, gamma = 500.0, 1.0, 0.05
    eng = _engine(energy_consts=(kappa, alpha, gamma, 0.0), temp=T, n_chains=n, seed=5, record=False)
    eng.run(600, 10)                                                # the soft modes (q ~ 0) equilibrate last
    acc0 = eng.accept_count_per_chain.clone()
    eng.run(30, 10)
    acc = ((eng.accept_count_per_chain - acc0).sum() / (300.0 * n)).item()
    assert 0.24 < acc < 0.36, acc                                   # Robbins-Monro target 0.3 (ME:101)
    c = eng.complex_params_per_chain.cpu().numpy()                  # [chains, 64]
    q = np.arange(64) - 32
    exact = T / (2.0 * (alpha + gamma * q ** 2))                    # per real component (a^2 ~ T/(2 kappa) negligible)
    for part in (c.real, c.imag):
        v = part.var(axis=0)
        z = (v - exact) / (exact * np.sqrt(2.0 / n))
        assert np.max(np.abs(z)) < 5.0, (np.argmax(np.abs(z)), np.max(np.abs(z)))
        assert np.max(np.abs(part.mean(axis=0)) / np.sqrt(exact / n)) < 5.0
    # after n > 50 the shared covariance is the pooled running covariance: its diagonal tracks 2 * exact
    diag = np.real(np.diag(eng.covariance_matrix_complex))
    assert np.all(diag > 0.5 * 2 * exact) and np.all(diag < 3.0 * 2 * exact + 0.2)


def test_hard_wall_and_api_surface():
    n = 256
    eng = _engine(energy_consts=(0.0, 1.0, 0.05, 1.0), temp=.5, n_chains=n, seed=9,
                  initial_real_params=np.array([0.95]), sampling_width=0.3)
    for _ in range(20):
        acc = eng.step_all()
    assert acc.dtype == torch.bool and acc.shape == (n,)
    eng.run(5, 20)
    a = eng.real_params_per_chain.cpu().numpy()
    assert np.all(np.abs(a) < 1.0)                                   # |a| >= 1 is never accepted (ME:247)
    assert eng.real_mean.shape == (1,) and eng.complex_mean.shape == (64,)
    assert eng.covariance_matrix_complex.shape == (64, 64) and eng.observables_mean.shape == (66,)
    df = eng.save_time_series()
    assert len(df) == 5 and list(df.columns)[:2] == ["abs_param_0", "abs_param_1"]
    assert list(df.columns)[-1] == "complex_group_sampling_width" and "total_energy" in df.columns
    # energy bookkeeping: the stored energy equals the energy of the stored state
    x = eng.state[:129].cpu().numpy()
    a0, c = x[0], x[1:65] + 1j * x[65:129]
    m2 = np.abs(c) ** 2
    qq = (np.arange(64) - 32)[:, None] ** 2
    e = 0.0 * a0 ** 2 + (1.0 * m2.sum(0) + 0.05 * (1 + a0 ** 2) * (qq * m2).sum(0)) + (1.0 / 128.0) * m2.sum(0) ** 2
    assert np.allclose(eng.energy_per_chain.cpu().numpy(), e, rtol=1e-10)


def test_results_do_not_depend_on_tiling_or_sharding():
    import metropolisengine_b200 as me
    kw = dict(energy_consts=(10.0, -1.0, 0.05, 1.0), temp=.1, seed=21, record=False)
    full = me.SharedCovarianceEngine(n_chains=512, **kw)
    full.step(7)
    lo = me.SharedCovarianceEngine(n_chains=256, **kw)
    lo.step(7)
    assert torch.equal(full.state[:, :256], lo.state)               # chain i depends on (seed, i) only


def test_quartic_energy_satisfies_the_virial_identity_per_mode():
    """Full cylinder energy (quartic coupling, a-dependent stiffness, hard wall).  For every unbounded real
    coordinate x of a law proportional to exp(-E/T), <x dE/dx> = T exactly; per complex mode q this reads
    < 2 |c_q|^2 (alpha + gamma q^2 (1 + a^2) + beta/64 sum_k |c_k|^2) > = 2 T.  Checked for all 64 modes on the
    equilibrated ensemble, averaged over 40 decorrelated snapshots (z-test over chains x snapshots, 5 sigma with the
    snapshot correlation bounded by using the chain-mean variance)."""
    T, (kappa, alpha, gamma, beta) = 0.1, (10.0, -1.0, 0.05, 1.0)
    n = 4096
    eng = _engine(energy_consts=(kappa, alpha, gamma, beta), temp=T, n_chains=n, seed=2, record=False)
    eng.run(800, 10)
    q2 = torch.tensor((np.arange(64) - 32.0) ** 2, device="cuda")
    acc = torch.zeros((n, 64), dtype=torch.float64, device="cuda")
    snaps = 40
    for _ in range(snaps):
        eng.run(5, 10)
        c = eng.complex_params_per_chain
        a = eng.real_params_per_chain[:, 0]
        m2 = c.real ** 2 + c.imag ** 2
        w = alpha + gamma * q2[None, :] * (1 + a[:, None] ** 2) + (beta / 64.0) * m2.sum(1, keepdim=True)
        acc += 2 * m2 * w
    v = (acc / snaps).cpu().numpy()                     # per chain time-average, [n, 64]
    z = (v.mean(0) - 2 * T) / (v.std(0, ddof=1) / np.sqrt(n))
    assert np.max(np.abs(z)) < 5.0, (int(np.argmax(np.abs(z))), float(np.max(np.abs(z))), v.mean(0)[:4])
    assert abs(v.mean() - 2 * T) < 5 * v.mean(1).std(ddof=1) / np.sqrt(n)
    assert np.all(np.abs(eng.real_params_per_chain.cpu().numpy()) < 1.0)


def test_pooled_moments_and_factor_match_float64_linear_algebra():
    """me_k4_moments (symmetric rank-k update of [Re c; Im c]) and me_k4_refactor (left-looking complex Cholesky)
    against torch float64: the increment of one measure, then the covariance and the BF16 factor built from it."""
    import metropolisengine_b200 as me
    n = 128 * 37 + 128                              # uneven split over the moment CTAs
    eng = me.SharedCovarianceEngine(temp=.1, n_chains=n, seed=4, record=False, sampling_width=0.05)
    eng.run(52, 4, fused_measure=False)             # crosses n > 50: the factor has been rebuilt from pooled moments
    eng.synchronize_refresh()                       # adopt the refresh of the last measure (it runs on a side stream)
    torch.cuda.synchronize()
    lay = eng._lay
    x = eng.state[lay.X:lay.X + lay.D].clone()      # [129, chains] as seen by the LAST measure
    a = x[0] - eng._shift[0]
    c = torch.complex(x[1:65] - eng._shift[1:65, None], x[65:129] - eng._shift[65:129, None])    # [64, chains]
    inc = eng._inc_full.clone()                     # increment of the last measure
    assert inc[0].real.item() == n
    assert torch.allclose(inc[2].real, a.sum(), rtol=1e-12, atol=1e-12)
    assert torch.allclose(inc[3].real, (a * a).sum(), rtol=1e-12)
    assert torch.allclose(inc[1].real, eng.state[lay.SIG].sum(), rtol=1e-12)
    assert torch.allclose(inc[4:68], c.sum(dim=1), rtol=1e-11, atol=1e-11)
    s2 = (c @ c.conj().t()).reshape(-1)
    assert torch.allclose(inc[68:], s2, rtol=1e-11, atol=1e-11 * s2.abs().max().item())
    # covariance from the accumulated moments and its factor
    mom = eng._mom
    N = mom[0].real
    s1 = mom[4:68]
    small = (inc[1].real / inc[0].real) ** 2 / eng.measure_step_counter
    cov = (mom[68:].reshape(64, 64) - torch.outer(s1, s1.conj()) / N) / (N - 1) + small * torch.eye(64, dtype=torch.complex128, device=x.device)
    assert torch.allclose(eng._cov_c, cov, rtol=1e-10, atol=1e-14)
    G = torch.linalg.cholesky(cov)
    gr, gi = G.real / 2 ** .5, G.imag / 2 ** .5
    B = torch.zeros((128, 128), dtype=torch.float64, device=x.device)
    B[0::2, 0::2] = gr; B[0::2, 1::2] = gi; B[1::2, 0::2] = -gi; B[1::2, 1::2] = gr
    want = B.view(128, 16, 8).permute(1, 0, 2).to(torch.bfloat16).float()
    got = eng._factor.float()
    assert (got - want).abs().max().item() <= 2 ** -7 * want.abs().max().item()      # equal up to one BF16 rounding
    assert int(eng._psd_status.item()) == 0


@pytest.mark.parametrize("nc,n_chains,async_refresh", [(64, 128 * 300, True), (64, 128 * 9, False), (16, 128 * 150, True),
                                                       (32, 128 * 301, False), (8, 128 * 40, True)])
def test_one_launch_block_equals_steps_then_measure(nc, n_chains, async_refresh, monkeypatch):
    """me_k4_step_measure (k steps + the measurement + the CTA partials of the pooled moments in ONE launch) against
    me_k4_step; me_k4_measure; me_k4_moments on a second engine with the same seed.  Until the first factor refresh (50
    blocks) the chains see identical proposals, so states, means, observable means and time series must be bit-identical;
    the tensor-core moments (two BF16 words per coordinate) agree with the FP64 ones to ~1e-6 of the diagonal
    scale, the scalar sums to FP64 rounding; so do the covariance and (up to single BF16 roundings) the
    factor built from them at the 50th block."""
    import metropolisengine_b200 as me
    if nc == 32:
        # the instantiation that parks the proposed state in TMEM, forced onto CTAs of several tiles: its measure tail
        # hands the moment sums over after every tile (the library would pick the other instantiation here)
        monkeypatch.setenv("ME_K4_XP", "1")
    rng = np.random.default_rng(nc)
    x0c = 0.05 * (rng.standard_normal((n_chains, nc)) + 1j * rng.standard_normal((n_chains, nc)))
    x0r = 0.1 * rng.standard_normal((n_chains, 1))
    kw = dict(temp=.1, n_chains=n_chains, seed=77, record=True, sampling_width=0.004, n_complex=nc, ts_chunk_rows=64,
              initial_real_params=x0r, initial_complex_params=x0c, async_refresh=async_refresh)
    a, b = me.SharedCovarianceEngine(**kw), me.SharedCovarianceEngine(**kw)
    lay = a._lay
    for blocks, k in ((7, 3), (43, 1)):
        a.run(blocks, k, fused_measure=True)
        b.run(blocks, k, fused_measure=False)
        torch.cuda.synchronize()
        assert a.measure_step_counter == b.measure_step_counter and a.steps_done == b.steps_done
        assert torch.equal(a.state, b.state)          # parameters, energy, widths, means, observable means, counts
        assert torch.equal(a.time_series(), b.time_series())
        ia, ib = a._inc_full, b._inc_full
        nw = 4 + nc
        assert ia[0].real.item() == n_chains
        scale1 = b.state[lay.X:lay.X + lay.D].abs().sum(dim=1).max().item()
        assert torch.allclose(ia[:4], ib[:4], rtol=1e-12, atol=1e-13 * scale1)          # count, sigma, a, a^2: FP64
        assert torch.allclose(ia[4:nw], ib[4:nw], rtol=0, atol=1e-6 * scale1)           # sum c: two BF16 words, FP32 per CTA
        sa, sb = ia[nw:].reshape(nc, nc), ib[nw:].reshape(nc, nc)
        d = torch.sqrt(sb.diagonal().real)
        rel = ((sa - sb).abs() / torch.outer(d, d)).max().item()
        assert rel < 2e-5, rel
    # the 50th block crossed n > 50: covariance and factor rebuilt from the accumulated moments
    a.synchronize_refresh(); b.synchronize_refresh()
    torch.cuda.synchronize()
    assert a.measure_step_counter == 51
    ca, cb = a._cov_c, b._cov_c
    d = torch.sqrt(cb.diagonal().real)
    assert ((ca - cb).abs() / torch.outer(d, d)).max().item() < 2e-5
    assert torch.allclose(a._cov_a, b._cov_a, rtol=1e-12)
    fa, fb = a._factor.float(), b._factor.float()
    assert (fa - fb).abs().max().item() <= 2 ** -7 * fb.abs().max().item()
    assert (fa != fb).float().mean().item() < 0.05
    assert int(a._psd_status.item()) == 0
    # and the sampler keeps going on the tensor-core moments
    a.run(20, 5)
    torch.cuda.synchronize()
    assert int(a._psd_status.item()) == 0 and a.acceptance_rate > 0.05


def test_asynchronous_factor_refresh_is_deterministic_and_lags_by_one_measure():
    """The factor refresh runs on a side stream beside the next block of steps; the block after measure b steps with
    the factor of measure b-1.  The schedule is fixed by the host, so two runs are bit-identical, and the sequential
    schedule (async_refresh=False) differs from it only through that one-measure lag."""
    import metropolisengine_b200 as me
    kw = dict(energy_consts=(10.0, -1.0, 0.05, 1.0), temp=.1, seed=33, record=False, n_chains=1024,
              sampling_width=0.004)      # small enough that the identity-covariance proposals are accepted from the start
    a = me.SharedCovarianceEngine(**kw)
    b = me.SharedCovarianceEngine(**kw)
    a.run(56, 3)
    b.run(56, 3)
    torch.cuda.synchronize()
    assert float(a.acceptance_rate) > 0.1
    assert torch.equal(a.state, b.state)
    assert np.array_equal(a.covariance_matrix_complex, b.covariance_matrix_complex)
    s = me.SharedCovarianceEngine(async_refresh=False, **kw)
    c = me.SharedCovarianceEngine(**kw)
    s.run(50, 3)                  # the first refresh happens at the 50th measure (n = 51 > 50, ME:389,396): up to and
    c.run(50, 3)                  # including it both schedules have stepped with the initial factor
    torch.cuda.synchronize()
    assert torch.equal(s.state, c.state)
    s.run(1, 3); c.run(1, 3)      # the next block uses the new factor only in the sequential schedule
    torch.cuda.synchronize()
    assert not torch.equal(s.state[:129], c.state[:129])


@pytest.mark.parametrize("nc,async_refresh", [(64, True), (16, True), (32, False)])
def test_oracle_runs_beside_the_kernel_with_identical_decisions(nc, async_refresh):
    """Level L-A of SURVEY §8c applied to the shared-covariance path.  A 256-chain ensemble is stepped one step per launch
    with the taps on; the C oracle holds the same ensemble.  Per step: (1) the oracle's own Philox / inverse-CDF BF16
    operand, real-parameter normal and accept uniform equal the kernel's EXACTLY (its quantile table is built independently
    with scipy); (2) the oracle's B.z in float64 equals the tensor-core increments within the
    FP32 accumulation tolerance 2e-5 |B||z|; (3) with the kernel's increments injected, every accept decision is identical
    and the FP64 state (parameters, energy, width, count) is BIT-identical over 56 measures x 4 steps — across the
    covariance switch-on at the 50th measure and the one-measure lag of the asynchronous factor refresh; (4) the pooled
    covariance and the BF16 factor the engine computes equal the oracle's."""
    import metropolisengine_b200 as me
    from oracle import k4_oracle as ko
    n, M, K = 256, 56, 4
    consts, T, seed = (10.0, -1.0, 0.05, 1.0), 0.1, 41
    x0c = 0.05 * np.exp(1j * np.arange(nc))
    eng = me.SharedCovarianceEngine(energy_consts=consts, temp=T, n_chains=n, seed=seed, record=False,
                                    initial_real_params=np.array([0.1]), initial_complex_params=x0c,
                                    sampling_width=0.02, async_refresh=async_refresh)
    x0 = np.concatenate([[0.1], x0c.real, x0c.imag])
    orc = ko.K4Ensemble(nc, n, consts, T, eng.ratio, seed=seed, use_wall=True, sampling_width=0.02, x0=x0,
                        async_refresh=async_refresh, groups=eng._lay.SUM_GROUPS)
    lay = eng._lay
    N = 2 * nc
    dz = torch.zeros((N, n), dtype=torch.float32, device="cuda")
    dd = torch.zeros((N, n), dtype=torch.float32, device="cuda")
    ds = torch.zeros((2, n), dtype=torch.float64, device="cuda")
    st0 = eng.state.cpu().numpy().T
    assert np.array_equal(st0[:, lay.X:lay.E], orc.state[:, lay.X:lay.E])                         # same initial parameters
    assert np.allclose(st0[:, :lay.NACC], orc.state[:, :lay.NACC], rtol=1e-14, atol=1e-16)        # (hypot, one-pass energy)
    orc.state[:, lay.E] = st0[:, lay.E]
    nacc = np.zeros(n)
    for im in range(M):
        for k in range(K):
            step = im * K + k
            eng.step(1, _dbg=(dz, dd, ds))
            torch.cuda.synchronize()
            zg, dg, sg = dz.cpu().numpy().T, dd.cpu().numpy().T, ds.cpu().numpy()
            orc.begin_step_launch()
            # (1) generator stream: integer arithmetic and table look-ups only — the operand is reproduced exactly
            assert np.array_equal(orc.normals(step), zg), step
            za_o, u_o = orc.scalars(step)
            assert np.array_equal(u_o, sg[1]) and np.array_equal(za_o, sg[0]), step
            # the factor in use is the oracle's own (after the switch-on it came through the engine's refresh kernel)
            Bg = eng._factor.float().permute(1, 0, 2).reshape(N, N).cpu().numpy()
            assert np.mean(Bg != orc.B_now) < 2e-3 and np.allclose(Bg, orc.B_now, rtol=2.0 ** -7, atol=1e-12), step
            orc.B_now = Bg
            assert np.isclose(float(eng._s_a.item()), orc.s_a_now, rtol=1e-12)
            orc.s_a_now = float(eng._s_a.item())
            # (2) tensor-core increments
            d_o, scale = orc.delta(zg)
            assert np.all(np.abs(d_o - dg) <= 2e-5 * scale + 1e-30), step
            # (3) injected step
            acc = orc.step_injected(dg, sg[0], sg[1])
            orc.end_step_launch()
            nacc += acc
            st = eng.state.cpu().numpy().T
            assert np.array_equal(st[:, lay.NACC], nacc), "accept/reject decisions differ at step %d" % step
            assert np.array_equal(st[:, lay.X:lay.SIG + 1], orc.state[:, lay.X:lay.SIG + 1]), step     # x, E, sigma
        eng.measure()
        orc.measure()
        st = eng.state.cpu().numpy().T
        assert np.allclose(st[:, lay.MEAN:lay.NACC], orc.state[:, lay.MEAN:lay.NACC], rtol=1e-12, atol=1e-15), im
        if orc.n_measure > 50:
            assert np.allclose(eng.covariance_matrix_complex, orc.cov_c, rtol=1e-9, atol=1e-14), im
            assert np.isclose(float(eng.covariance_matrix_real[0, 0]), orc.cov_a, rtol=1e-9), im
    assert 0.05 < nacc.sum() / (n * M * K) < 0.95
    assert eng.measure_step_counter == orc.n_measure == M + 1


def test_user_functor_on_the_tensor_core_path_matches_the_builtin_one():
    """The energy plugin on the shared-covariance path: CUDA text with the sufficient-statistics contract
    (me_k4_set_energy_source), compiled by NVRTC into the same warp-specialised kernel.  Restating the cylinder energy
    must reproduce the built-in functor's chains; a different energy (no quartic term, no wall) must sample its exact law."""
    import metropolisengine_b200 as me
    src = """
__device__ void me_k4_mode(double q, double re, double im, const double* k, double& s0, double& s1) {
    const double m2 = fma(re, re, im * im);
    s0 += m2;
    s1 = fma(q * q, m2, s1);
}
__device__ double me_k4_total(double a, double s0, double s1, const double* k, int nc) {
    const double a2 = a * a;
    const double inner = fma(k[1], s0, (k[2] * (1.0 + a2)) * s1);
    return fma(k[3] / (2.0 * (double)nc), s0 * s0, fma(k[0], a2, inner));
}
__device__ bool me_k4_reject(double a, const double* k) { return fabs(a) >= 1.0; }
"""
    consts = (10.0, -1.0, 0.05, 1.0)
    kw = dict(temp=.1, n_chains=512, seed=6, record=False, n_complex=32)
    a = me.SharedCovarianceEngine(energy_consts=consts, **kw)
    b = me.SharedCovarianceEngine(energy=me.SharedEnergy(src, consts=consts, has_reject=True), **kw)
    a.run(53, 5)
    b.run(53, 5)
    torch.cuda.synchronize()
    assert torch.equal(a.accept_count_per_chain, b.accept_count_per_chain)
    assert torch.equal(a.state, b.state)
    # a functor of its own: independent modes with stiffness k0 + k1 q^4, harmonic amplitude, no wall
    src2 = """
__device__ void me_k4_mode(double q, double re, double im, const double* k, double& s0, double& s1) {
    const double m2 = re * re + im * im;
    s0 += m2;
    s1 += (q * q) * (q * q) * m2;
}
__device__ double me_k4_total(double a, double s0, double s1, const double* k, int nc) {
    return k[2] * a * a + k[0] * s0 + k[1] * s1;
}
"""
    T, k0, k1, k2 = 0.2, 1.0, 0.01, 4.0
    c = me.SharedCovarianceEngine(energy=me.SharedEnergy(src2, consts=(k0, k1, k2)), reject_condition=False, temp=T,
                                  n_chains=4096, seed=8, record=False, n_complex=8)
    c.run(500, 10)
    v = c.complex_params_per_chain.cpu().numpy()
    q = np.arange(8) - 4.0
    exact = T / (2.0 * (k0 + k1 * q ** 4))
    for part in (v.real, v.imag):
        z = (part.var(axis=0) - exact) / (exact * np.sqrt(2.0 / 4096))
        assert np.max(np.abs(z)) < 5.0, z
    av = c.real_params_per_chain.cpu().numpy()[:, 0].var()
    assert abs(av - T / (2 * k2)) < 5 * (T / (2 * k2)) * np.sqrt(2.0 / 4096)
    with pytest.raises(Exception):
        me.SharedCovarianceEngine(energy=me.SharedEnergy("not CUDA"), temp=T, n_chains=128, n_complex=8)


def test_pooled_engine_matches_the_per_chain_engine_on_the_quartic_energy():
    """Ensemble cross-check: the shared-covariance engine against the per-chain engine (the reference's own algorithm,
    pinned at this kind of shape by the goldens cyl_1r64c / magphase_1r16c recorded from the live reference) on the
    full cylinder energy with its wall, 1 real + 16 complex.  Both sample the same law: two-sample KS on the marginals of
    the amplitude, of the softest and the stiffest mode, z-test on <sum |c|^2>, acceptance near the target."""
    import metropolisengine_b200 as me
    nc, n, T = 16, 4096, 0.1
    consts = (10.0, -1.0, 0.05, 1.0)
    pooled = me.MetropolisEngine(me.BuiltinEnergy("cylinder", *consts, reject=True), initial_real_params=np.array([0.0]),
                                 initial_complex_params=np.zeros(nc, dtype=complex), temp=T, n_chains=n, seed=12,
                                 record=False, adapt="pooled")
    per = me.MetropolisEngine(me.BuiltinEnergy("cylinder", *consts, reject=True), initial_real_params=np.array([0.0]),
                              initial_complex_params=np.zeros(nc, dtype=complex), temp=T, n_chains=n, seed=13,
                              record=False)
    assert isinstance(pooled, me.SharedCovarianceEngine) and per._generic
    pooled.run(500, 10)
    per.run(500, 10)
    a0, a1 = pooled.accept_count_per_chain.clone(), per.accept_count_per_chain.clone()
    pooled.run(40, 10)
    per.run(40, 10)
    for eng, acc0 in ((pooled, a0), (per, a1)):
        acc = ((eng.accept_count_per_chain - acc0).sum() / (400.0 * n)).item()
        assert 0.22 < acc < 0.40, acc
    cp, cq = pooled.complex_params_per_chain.cpu().numpy(), per.complex_params_per_chain.cpu().numpy()
    ap, aq = pooled.real_params_per_chain.cpu().numpy()[:, 0], per.real_params_per_chain.cpu().numpy()[:, 0]
    assert stats.ks_2samp(ap, aq).pvalue > 1e-3
    for j in (0, nc // 2, nc - 1):                     # q = -8 (stiff), 0 (soft: alpha < 0, held by the quartic term), 7
        assert stats.ks_2samp(np.abs(cp[:, j]), np.abs(cq[:, j])).pvalue > 1e-3, j
        assert stats.ks_2samp(cp[:, j].real, cq[:, j].real).pvalue > 1e-3, j
    tp, tq = (np.abs(cp) ** 2).sum(1), (np.abs(cq) ** 2).sum(1)
    z = (tp.mean() - tq.mean()) / np.sqrt(tp.var() / n + tq.var() / n)
    assert abs(z) < 5.0, z
    assert np.all(np.abs(ap) < 1.0) and np.all(np.abs(aq) < 1.0)
