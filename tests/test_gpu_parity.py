"""GPU parity tests: the CUDA path (through the C ABI, via the Python facade) against the oracle.

Level L-A: a single strict chain driven by the reference's own recorded draws must reproduce the
reference chain's accept/reject decisions exactly and its state within 1e-12 relative (FP64).
Philox level: a throughput-mode ensemble must match the C oracle running the same Philox streams chain by chain.
"""
import numpy as np
import pytest
import torch

from tests.conftest import load_golden
from tests.golden.cases import cases, fresh_ctor

pytestmark = pytest.mark.gpu

CASES = cases()
RTOL = 1e-12          # north_star: parameters within 1e-12 relative (FP64)

USER_SOURCES = {
    # E = (|c0|^2 - 1)^2 + |c1 - c0|^2      (tests/golden/cases.py::_pure_2c)
    "pure_2c": """
__device__ double me_user_energy(const double* x, const double* cr, const double* ci, const double* k) {
    const double a = cr[0] * cr[0] + ci[0] * ci[0];
    const double dr = cr[1] - cr[0], di = ci[1] - ci[0];
    return (a - 1.0) * (a - 1.0) + (dr * dr + di * di);
}
""",
    # E = sum (1-r_i)^2 + r0 r1 mean_j(-|c_j|^2 + .5 |c_j|^4)   (tests/golden/cases.py::_warm_3r2c)
    "warm_3r2c": """
__device__ double me_user_energy(const double* x, const double* cr, const double* ci, const double* k) {
    double area = 0.0, s = 0.0;
    for (int i = 0; i < ME_NR; i++) { const double e = 1.0 - x[i]; area += e * e; }
    for (int j = 0; j < ME_NC; j++) { const double a = cr[j] * cr[j] + ci[j] * ci[j]; s += -1.0 * a + .5 * (a * a); }
    return area + x[0] * x[1] * (s / ME_NC);
}
""",
}


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.max(np.abs(b)) if b.size else 1.0
    return bool(np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(b), scale)))


def make_engine(name, me, **extra):
    case, g = CASES[name], load_golden(name)
    ctor = fresh_ctor(case)
    if case["builtin"] is not None:
        bname, consts = case["builtin"]
        energy = me.BuiltinEnergy(bname, *consts, reject=("reject" in case))
    elif name in USER_SOURCES:
        energy = me.CudaEnergy(USER_SOURCES[name])
    else:
        raise KeyError(name)
    return me.MetropolisEngine(energy, **ctor, **extra), g


def drive_injected_and_compare(eng, g):
    M, K = int(g["n_measures"]), int(g["steps_per_measure"])
    n_r, n_c = int(g["n_r"]), int(g["n_c"])
    nacc = 0
    for im in range(M):
        sl = slice(im * K, (im + 1) * K)
        eng.run_injected(g["delta"][sl], g["u"][sl], 1, K)
        nacc += int(g["accept"][sl].sum())
        assert int(eng.accept_count_per_chain.item()) == nacc, "accept/reject decisions differ in block %d" % im
        last = g["step_x"][(im + 1) * K - 1]
        if n_r:
            assert close(eng.real_params, last[:n_r]), im
            assert close(eng.real_group_sampling_width, g["m_sigma_r"][im]), im
            assert close(eng.real_mean, g["m_real_mean"][im]), im
            assert close(eng.covariance_matrix_real, g["m_cov_r"][im]), im
        if n_c:
            assert close(eng.complex_params, last[n_r:n_r + n_c] + 1j * last[n_r + n_c:]), im
            assert close(eng.complex_group_sampling_width, g["m_sigma_c"][im]), im
            assert close(eng.complex_mean, g["m_complex_mean"][im]), im
            assert close(eng.covariance_matrix_complex, g["m_cov_c"][im]), im
        assert close(eng.observables_mean, g["m_obs_mean"][im]), im
        assert close(eng.energy["total"], g["m_energy"][im], 1e-11), im
    assert eng.measure_step_counter == int(g["measure_step_counter"])
    eng.check_status()


@pytest.mark.parametrize("name", ["kat1_x2", "kat2_xy", "kat3_2r1c", "c3_3r4c", "cyl_1r8c_reject", "xy_temp0"])
def test_injected_parity_builtin(name):
    import metropolisengine_b200 as me
    eng, g = make_engine(name, me, strict=True)
    drive_injected_and_compare(eng, g)


def test_injected_parity_cylinder_shape_1r64c_per_chain_covariance():
    """BASELINE config 4 shape with the reference's own per-chain algorithm (128x128 embedded proposal covariance,
    metropolis_engine.py:274-302): runtime-shape kernels (me_generic.cu), draw-injected against the reference."""
    import metropolisengine_b200 as me
    eng, g = make_engine("cyl_1r64c", me, strict=True)
    assert eng._generic
    drive_injected_and_compare(eng, g)


def test_large_shape_philox_matches_c_oracle():
    """Same shape in Philox mode: proposals through the per-chain 64x64 complex Cholesky factor in global memory
    follow the C oracle chain by chain."""
    import metropolisengine_b200 as me
    from oracle import c_oracle as co
    n, M, K = 64, 54, 2
    consts = [10.0, -1.0, 0.05, 1.0]
    eng = me.MetropolisEngine(me.BuiltinEnergy("cylinder", *consts, reject=True), initial_real_params=np.array([0.2]),
                              initial_complex_params=np.zeros(64, dtype=complex), temp=.1, sampling_width=0.012,
                              n_chains=n, seed=99)
    eng.run(M, K)
    eng.check_status()
    st = eng.state.cpu().numpy()
    lay = eng._lay
    for ch in (0, 33, 63):
        o = co.CChain(1, 64, "cylinder", consts=consts, temp=.1, sampling_width=0.012,
                      x0=np.concatenate([[0.2], np.zeros(128)]), use_reject=True)
        acc, _ = o.run(M, K, True, seed=99, chain_id=ch)
        assert st[lay.NACC, ch] == acc.sum()
        assert close(st[:lay.WORDS - 2, ch], o.state[:lay.WORDS - 2], 1e-9), ch


def test_large_shape_one_launch_schedule_equals_the_step_by_step_path():
    """Runtime shapes (D > 32) with a built-in functor run a whole schedule in one launch (gk_run: propose -> energy + wall
    -> decide per step, measure per block, all inside the kernel).  It must equal the three-launches-per-step path bit
    for bit (same device functions, same Philox slots) — forced here by an always-false python predicate — including the
    stored rows, and group-wise steps."""
    import metropolisengine_b200 as me
    kw = dict(initial_real_params=np.array([0.2]), initial_complex_params=np.zeros(16, dtype=complex), temp=.1,
              sampling_width=0.05, n_chains=96, seed=17)
    consts = (10.0, -1.0, 0.05, 1.0)
    a = me.MetropolisEngine(me.BuiltinEnergy("cylinder", *consts, reject=True), **kw)
    b = me.MetropolisEngine(me.BuiltinEnergy("cylinder", *consts, reject=True), **kw)
    b.set_reject_condition(lambda r, c: torch.zeros(r.shape[0], dtype=torch.bool, device=r.device))
    assert a._generic and not a._unfused() and b._unfused()
    l0 = a.launch_count
    a.run(55, 4)
    assert a.launch_count - l0 == 1                                   # the whole schedule: one launch
    b.run(55, 4)
    a.step_real_group(3); b.step_real_group(3)
    a.step_complex_group(2); b.step_complex_group(2)
    torch.cuda.synchronize()
    assert torch.equal(a.state, b.state)
    assert torch.equal(a.time_series(), b.time_series()) and a.time_series().shape[0] == 55
    assert a.measure_step_counter == b.measure_step_counter == 56 and a.steps_done == b.steps_done == 225


@pytest.mark.parametrize("name", ["pure_2c", "warm_3r2c"])
def test_injected_parity_user_functor_nvrtc(name):
    """User CUDA functors compiled at run time (NVRTC, --fmad=false) and fused into the step kernel."""
    import metropolisengine_b200 as me
    eng, g = make_engine(name, me, strict=True)
    drive_injected_and_compare(eng, g)


def test_injected_parity_whole_schedule_single_launch():
    """The whole KAT2 schedule (1000 x (10 steps + measure)) in ONE launch reproduces the reference's final
    values of SURVEY §4 KAT2."""
    import metropolisengine_b200 as me
    eng, g = make_engine("kat2_xy", me, strict=True)
    eng.run_injected(g["delta"], g["u"], 1000, 10)
    assert int(eng.accept_count_per_chain.item()) == 3167
    assert close(eng.real_params, [0.2477655113830379, -0.0020188999856143724])
    assert close(eng.real_group_sampling_width, 0.6737107772882164)
    assert close(eng.real_mean, [-0.002096715114169222, 0.004685569814294245])
    assert close(eng.covariance_matrix_real[0, 0], 0.479512502307138)
    assert eng.sampling_width == 0.05            # all-real engines never adapt this attribute (App. B-1)
    # time series rows equal the per-measure snapshots
    ts = eng.time_series().cpu().numpy()[:, :, 0]
    assert ts.shape == (1000, 4)
    assert close(ts[:, :2], g["m_x"])
    assert close(ts[:, 3], g["m_sigma_r"])


def test_kat1_readme_loop_with_python_level_calls():
    """README.md:39-44 driven call by call: step_all() returns the reference's decisions (python bools)."""
    import metropolisengine_b200 as me
    eng, g = make_engine("kat1_x2", me, strict=True)
    decisions = []
    for s in range(200):
        eng.run_injected(g["delta"][s:s + 1], g["u"][s:s + 1], 1, 1, do_measure=False)
        decisions.append(bool(eng._last_accept.item()))
        eng.measure()
    assert decisions == [bool(a) for a in g["accept"][:200]]
    assert close(eng.real_mean, g["m_real_mean"][199])
    assert close(eng.observables_mean, g["m_obs_mean"][199])


def test_torch_callable_injected_parity_dict_terms():
    """Dict-of-terms energy (ME:111-115, demo/toymodel_complex_and_real.py:33-35) as torch callables on the
    unfused propose / callable / accept path, draw-injected against the reference."""
    import metropolisengine_b200 as me
    g = load_golden("dict_2r1c")
    k, al, be = 1.0, -1.0, 0.5
    area = lambda r, c: k * (1 - r[0]) ** 2 + k * (1 - r[1]) ** 2
    field = lambda r, c: r[0] * r[1] * (al * (c[0] * c[0].conj()).real + be * (c[0] * c[0].conj()).real ** 2)
    terms = {"complex": {"field": field}, "real": {"field": field, "area": area},
             "all": {"field": field, "area": area}}
    eng = me.MetropolisEngine(terms, initial_real_params=np.array([0.2, 0.1]),
                              initial_complex_params=np.array([0.5 + 0j]), temp=.1, strict=True,
                              callable_layout="params_first")
    M, K = int(g["n_measures"]), int(g["steps_per_measure"])
    eng.run_injected(g["delta"], g["u"], M, K)
    assert int(eng.accept_count_per_chain.item()) == int(g["accept"].sum())
    assert close(eng.real_params, g["step_x"][-1][:2], 1e-11)
    assert close(eng.sampling_width, g["m_sigma"][-1], 1e-11)
    assert close(eng.covariance_matrix_real, g["m_cov_r"][-1], 1e-10)
    assert close(eng.covariance_matrix_complex, g["m_cov_c"][-1], 1e-10)
    df = eng.save_time_series()
    want = [str(c) for c in g["df_columns"]]
    # the reference iterates a *set* of term names (ME:113-115,153-155), so the order of the <term>_energy
    # columns depends on PYTHONHASHSEED; everything else is ordered
    got = list(df.columns)
    assert [c for c in got if not c.endswith("_energy")] == [c for c in want if not c.endswith("_energy")]
    assert sorted(got) == sorted(want)
    assert got.index("field_energy") in (4, 5) and len(df) == M


def test_readme_example_with_torch_callable():
    """README.md:23-58 with the energy as a python callable (params_first layout lets the reference-style body
    vectorise unchanged); checks the reference's read API."""
    import metropolisengine_b200 as me
    eng = me.MetropolisEngine(lambda real_params, complex_params: real_params[0] ** 2, initial_real_params=[0.0],
                              temp=.01, callable_layout="params_first", seed=1)
    n_acc = 0
    for i in range(300):
        a = eng.step_all()
        assert isinstance(a, bool)
        n_acc += a
        eng.measure()
    assert 60 < n_acc < 260
    assert eng.real_mean.shape == (1,) and eng.covariance_matrix_real.shape == (1, 1)
    names = list(zip(eng.observables_names, eng.observables_mean))
    assert [n for n, _ in names] == ["abs_param_0", "param_0_squared"]
    eng.save_time_series()
    assert list(eng.df.columns) == ["abs_param_0", "param_0_squared", "total_energy", "param_0",
                                    "real_group_sampling_width"]
    assert len(eng.df) == 300


# ---------------------------------------------------------------------------------------------- Philox level
PHILOX_CASES = [
    ("x2", 1, 0, [], dict(temp=.01), 40, 1),
    ("xy_well", 2, 0, [1.0], dict(temp=.1), 70, 10),
    ("mixed_well", 2, 1, [1.0, -1.0, 0.5], dict(temp=.1), 70, 5),
    ("mixed_well", 3, 4, [1.0, -1.0, 0.5], dict(temp=.1), 70, 5),
    ("cylinder", 1, 8, [10.0, -1.0, 0.05, 1.0], dict(temp=.1, sampling_width=0.2), 60, 4),
]


@pytest.mark.parametrize("strict", [False, True])
@pytest.mark.parametrize("bname,n_r,n_c,consts,kw,M,K", PHILOX_CASES)
def test_philox_ensemble_matches_c_oracle_chain_by_chain(bname, n_r, n_c, consts, kw, M, K, strict):
    """Same Philox key/counter definition on both sides: every chain of the CUDA ensemble follows the oracle's
    chain (crosses n > 50, so the in-kernel Cholesky refactorisation is exercised)."""
    import metropolisengine_b200 as me
    from oracle import c_oracle as co
    n = 96
    x0r = np.full(n_r, 0.3) if bname == "cylinder" else np.zeros(n_r)
    x0c = np.zeros(n_c, dtype=complex)
    eng = me.MetropolisEngine(me.BuiltinEnergy(bname, *consts, reject=(bname == "cylinder")),
                              initial_real_params=x0r if n_r else None,
                              initial_complex_params=x0c if n_c else None, n_chains=n, seed=1234, strict=strict, **kw)
    eng.run(M, K)
    eng.check_status()
    st = eng.state.cpu().numpy()
    lay = eng._lay
    tol = 2e-10 if strict else 2e-9
    for ch in (0, 1, 31, 32, 95):
        o = co.CChain(n_r, n_c, bname, consts=consts, temp=kw["temp"], sampling_width=kw.get("sampling_width", 0.05),
                      x0=np.concatenate([x0r, x0c.real, x0c.imag]), use_reject=(bname == "cylinder"))
        acc, _ = o.run(M, K, True, seed=1234, chain_id=ch)
        assert st[lay.NACC, ch] == acc.sum(), (ch, st[lay.NACC, ch], acc.sum())
        assert close(st[:lay.WORDS - 2, ch], o.state[:lay.WORDS - 2], tol), ch


def test_results_do_not_depend_on_sharding():
    """Philox counters carry GLOBAL chain ids: one 96-chain engine == two engines owning [0,40) and [40,96)."""
    import metropolisengine_b200 as me
    kw = dict(initial_real_params=np.array([0., 0., 0.]), initial_complex_params=np.zeros(4, dtype=complex), temp=.1,
              seed=9)
    e_all = me.MetropolisEngine(("mixed_well", 1.0, -1.0, 0.5), n_chains=96, **kw)
    e_all.run(60, 4)
    parts = []
    for lo, hi in ((0, 40), (40, 96)):
        e = me.MetropolisEngine(("mixed_well", 1.0, -1.0, 0.5), n_chains=96, _shard=(lo, hi), **kw)
        e.run(60, 4)
        parts.append(e.state)
    assert torch.equal(torch.cat(parts, dim=1), e_all.state)


def test_checkpoint_resume_is_bit_identical():
    import metropolisengine_b200 as me
    kw = dict(initial_real_params=np.array([0., 0.]), temp=.1, n_chains=256, seed=3)
    a = me.MetropolisEngine(("xy_well", 1.0), **kw)
    a.run(30, 7)
    sd = a.state_dict()
    a.run(40, 7)
    b = me.MetropolisEngine(("xy_well", 1.0), **kw)
    b.load_state_dict(sd)
    b.run(40, 7)
    assert torch.equal(a.state, b.state)


def test_single_steps_equal_fused_run():
    """1000 python-level step_all()/measure() calls == one fused launch (same Philox counters)."""
    import metropolisengine_b200 as me
    kw = dict(initial_real_params=np.array([0., 0.]), temp=.1, n_chains=64, seed=3)
    a = me.MetropolisEngine(("xy_well", 1.0), **kw)
    for _ in range(60):
        for _ in range(3):
            acc = a.step_all()
        a.measure()
    assert acc.dtype == torch.bool and acc.shape == (64,)
    b = me.MetropolisEngine(("xy_well", 1.0), **kw)
    b.run(60, 3)
    assert torch.equal(a.state, b.state)
    assert torch.equal(a.time_series(), b.time_series())


def test_external_callable_matches_fused_functor():
    """The unfused propose / torch-callable / accept path consumes the same Philox slots as the fused kernel."""
    import metropolisengine_b200 as me
    kw = dict(initial_real_params=np.array([0., 0.]), temp=.1, n_chains=128, seed=21)
    a = me.MetropolisEngine(("xy_well", 1.0), **kw)
    a.run(55, 3)
    b = me.MetropolisEngine(lambda r, c: 1.0 * (r[:, 0] * r[:, 0] + r[:, 1] * r[:, 1]), **kw)
    b.run(55, 3)
    assert torch.equal(a.accept_count_per_chain, b.accept_count_per_chain)
    assert torch.allclose(a.state, b.state, rtol=1e-11, atol=1e-13)


def test_status_flags_surface_as_reference_exceptions():
    import metropolisengine_b200 as me
    eng = me.MetropolisEngine(me.CudaEnergy("""
__device__ double me_user_energy(const double* x, const double* cr, const double* ci, const double* k) {
    return x[0] > 0.2 ? nan("") : x[0] * x[0];
}"""), initial_real_params=[0.0], temp=1.0, n_chains=64, sampling_width=0.5)
    eng.run(5, 20)
    with pytest.raises(FloatingPointError):
        eng.check_status()
    with pytest.raises(ValueError):
        me.MetropolisEngine("x2", temp=.1)                          # ME:37-39
    with pytest.raises(AssertionError):
        me.MetropolisEngine("x2", initial_real_params=[0.0], temp=-1)   # ME:92
    with pytest.raises(_me_error()):
        me.MetropolisEngine(me.CudaEnergy("this is not CUDA"), initial_real_params=[0.0], temp=.1)


def _me_error():
    from metropolisengine_b200._lib import MeError
    return MeError


def test_statistical_inefficiency_kernel_matches_oracle_estimator():
    """ESS denominator: the device kernel equals the oracle's estimator (pymbar's definition) on the stored series."""
    import metropolisengine_b200 as me
    from oracle.py_port import statistical_inefficiency
    eng = me.MetropolisEngine(("xy_well", 1.0), initial_real_params=np.array([0., 0.]), temp=.1, n_chains=64, seed=4)
    eng.run(600, 2)
    g = eng.statistical_inefficiency(column=0, n_chains=8, burn_in=0.25).cpu().numpy()
    ts = eng.time_series().cpu().numpy()
    for ch in range(8):
        want = statistical_inefficiency(ts[150:, 0, ch])
        assert abs(g[ch] - want) <= 1e-9 * want, (ch, g[ch], want)
    assert np.all(g >= 1.0)


def test_on_disk_time_series_formats(tmp_path):
    """CSV in the reference's df.to_csv format (complex printed as (a+bj)) and the dense npz."""
    import pandas
    import metropolisengine_b200 as me
    eng = me.MetropolisEngine(("mixed_well", 1.0, -1.0, 0.5), initial_real_params=np.array([0., 0.]),
                              initial_complex_params=np.array([0j]), temp=.1, n_chains=8, seed=1)
    eng.run(12, 3)
    path = eng.to_csv(str(tmp_path / "series.csv"), chain=2)
    back = pandas.read_csv(path, index_col=0)
    assert list(back.columns) == ["abs_param_0", "abs_param_1", "abs_param_2", "param_0_squared", "param_1_squared",
                                  "total_energy", "param_0", "param_1", "real_group_sampling_width", "param_2",
                                  "complex_group_sampling_width"]
    assert len(back) == 12 and str(back["param_2"][0]).startswith("(") and str(back["param_2"][0]).endswith("j)")
    c = back["param_2"].apply(complex)
    assert np.allclose(np.abs(c), back["abs_param_2"], rtol=1e-12)
    z = np.load(eng.save_npz(str(tmp_path / "series.npz")))
    assert z["rows"].shape == (12, 7, 8) and list(z["columns"])[-3] == "energy"
    assert np.allclose(z["rows"][:, 0, 2], back["param_0"], rtol=1e-15)


@pytest.mark.parametrize("name", ["groups_2r1c", "groups_widths_2r1c"])
def test_group_wise_stepping_injected_parity(name):
    """step_real_group / step_complex_group on a mixed engine, each with its own width
    (ME:209-239, 440-456), alternating as the cylinder app drives them; draw-injected against the reference.
    groups_widths_2r1c starts from sampling_width=[sigma_real, sigma_complex] (ME:93-95)."""
    import metropolisengine_b200 as me
    eng, g = make_engine(name, me, strict=True)
    if name == "groups_widths_2r1c":
        assert eng.real_group_sampling_width == 0.21 and eng.complex_group_sampling_width == 0.04
    M, K = int(g["n_measures"]), int(g["steps_per_measure"])
    nacc = 0
    for im in range(M):
        for s in range(im * K, (im + 1) * K):
            eng.run_injected_group(int(g["group"][s]), g["delta"][s:s + 1], g["u"][s:s + 1], 1)
            nacc += int(g["accept"][s])
        eng.measure()
        assert int(eng.accept_count_per_chain.item()) == nacc, im
        assert close(eng.real_params, g["m_x"][im]) and close(eng.complex_params, g["m_c"][im]), im
        assert close(eng.real_group_sampling_width, g["m_sigma_r"][im]), im
        assert close(eng.complex_group_sampling_width, g["m_sigma_c"][im]), im
        assert close(eng.covariance_matrix_real, g["m_cov_r"][im]) and close(eng.covariance_matrix_complex, g["m_cov_c"][im])
        assert close(eng.energy["total"], g["m_energy"][im], 1e-11), im
    assert eng.real_group_sampling_width != eng.complex_group_sampling_width
    assert eng.step_counter == 1 + M * K // 2                      # only complex-group steps count (ME:450)
    df = eng.save_time_series()
    assert close(df["real_group_sampling_width"], g["m_sigma_r"]) and close(df["complex_group_sampling_width"], g["m_sigma_c"])


def test_group_wise_stepping_philox_matches_c_oracle():
    import metropolisengine_b200 as me
    from oracle import c_oracle as co
    kw = dict(initial_real_params=np.array([0.3, 0.2]), initial_complex_params=np.array([0.4 - 0.1j]), temp=.1)
    eng = me.MetropolisEngine(("mixed_well", 1.0, -1.0, 0.5), n_chains=32, seed=8, **kw)
    o = co.CChain(2, 1, "mixed_well", consts=[1.0, -1.0, 0.5], temp=.1, x0=np.array([0.3, 0.2, 0.4, -0.1]))
    step = 0
    for im in range(55):
        for _ in range(3):
            eng.step_real_group(2)
            o.run(1, 2, False, seed=8, chain_id=5, step0=step, group=1); step += 2
            acc = eng.step_complex_group()
            o.run(1, 1, False, seed=8, chain_id=5, step0=step, group=2); step += 1
        eng.measure()
        o.run(1, 0, True, seed=8, chain_id=5, step0=step)
    assert acc.shape == (32,)
    st = eng.state.cpu().numpy()
    assert close(st[:eng._lay.WORDS - 2, 5], o.state[:eng._lay.WORDS - 2], 2e-9)
    with pytest.raises(ValueError):
        me.MetropolisEngine("x2", initial_real_params=[0.0], temp=.1).step_complex_group()


@pytest.mark.parametrize("name", ["magphase_2r1c", "magphase_1r16c"])
def test_magnitude_phase_injected_parity(name):
    """complex_sample_method="magnitude-phase" (ME:129-130, 168-207, 304-317).  The reference is
    driven as real group / magnitude move / phase redraw; the recorded proposals are injected into the strict kernel.
    magphase_1r16c: the same on a large parameter space (runtime-shape kernels, hard wall, zero-modulus starts)."""
    import metropolisengine_b200 as me
    eng, g = make_engine(name, me, strict=True)
    assert eng._generic == (name == "magphase_1r16c")
    assert eng.complex_sample_method == "magnitude-phase"
    M, K = int(g["n_measures"]), int(g["steps_per_measure"])
    nacc = 0
    for im in range(M):
        for s in range(im * K, (im + 1) * K):
            eng.run_injected_group(int(g["group"][s]), g["delta"][s:s + 1], g["u"][s:s + 1], 1)
            nacc += int(g["accept"][s])
            assert int(eng.accept_count_per_chain.item()) == nacc, s
        eng.measure()
        assert close(eng.real_params, g["m_x"][im]) and close(eng.complex_params, g["m_c"][im]), im
        assert close(eng.real_group_sampling_width, g["m_sigma_r"][im]), im
        assert close(eng.complex_group_sampling_width, g["m_sigma_c"][im]), im     # phase redraws never adapt it
        assert close(eng.covariance_matrix_real, g["m_cov_r"][im]) and close(eng.covariance_matrix_complex, g["m_cov_c"][im])
        assert close(eng.energy["total"], g["m_energy"][im], 1e-11), im
    assert eng.step_counter == 1 + M * K // 3                      # only the magnitude half counts (ME:450)


@pytest.mark.parametrize("x0c", [np.array([0.4 - 0.1j, -0.3 + 0.2j]), np.array([0j, -0.3 + 0.2j])],
                         ids=["nonzero", "zero_modulus"])
def test_magnitude_phase_philox_matches_c_oracle(x0c):
    """Device-generated magnitude / phase moves (normal z_j of the step; angle word of Philox call j) against the C
    oracle running the same streams; step_complex_group() = magnitude + phase and returns None as in the reference.
    zero_modulus: a coefficient starting at 0 takes cmath.polar's signed-zero argument (arg(-0) = pi) until a move is
    accepted (tests/test_oracle_c.py pins the oracle to cmath for this)."""
    import metropolisengine_b200 as me
    from oracle import c_oracle as co
    kw = dict(initial_real_params=np.array([0.3, 0.2, 0.1]), initial_complex_params=x0c, temp=.1, sampling_width=0.6,
              complex_sample_method="magnitude-phase")
    src = USER_SOURCES["warm_3r2c"]
    eng = me.MetropolisEngine(me.CudaEnergy(src), n_chains=32, seed=21, **kw)

    def energy(x, n_r, n_c):
        r, c = x[:n_r], x[n_r:n_r + n_c] + 1j * x[n_r + n_c:]
        a = (c * c.conjugate()).real
        return float(np.sum((1 - r) ** 2) + r[0] * r[1] * np.mean(-1 * a + .5 * a ** 2))
    o = co.CChain(3, 2, lambda x: energy(x, 3, 2), temp=.1, x0=np.concatenate([[0.3, 0.2, 0.1], x0c.real, x0c.imag]),
                  sampling_width=0.6)
    step = 0
    for im in range(56):
        for _ in range(2):
            eng.step_real_group()
            o.run(1, 1, False, seed=21, chain_id=9, step0=step, group=1); step += 1
            assert eng.step_complex_group() is None
            o.run(1, 1, False, seed=21, chain_id=9, step0=step, group=3); step += 1
            o.run(1, 1, False, seed=21, chain_id=9, step0=step, group=4); step += 1
        eng.measure()
        o.run(1, 0, True, seed=21, chain_id=9, step0=step)
    st = eng.state.cpu().numpy()
    assert close(st[:eng._lay.WORDS - 2, 9], o.state[:eng._lay.WORDS - 2], 2e-9)
    # phases are redrawn uniformly: the ensemble's arguments of c_0 cover all four quadrants
    ang = np.angle(eng.complex_params_per_chain.cpu().numpy()[:, 0])
    assert len(set(np.floor(ang / (np.pi / 2)).astype(int))) == 4


def test_user_functor_on_a_runtime_shape_matches_the_builtin_functor():
    """A user CUDA functor beyond D = 32 (ME:20, 110-120: the energy plugin is the reference's extension point) runs in the
    runtime-shape kernels compiled around it.  The cylinder energy restated as a user functor (same operation order, hard
    wall included) must give the built-in functor's chains bit for bit: one-launch schedules and single steps."""
    import metropolisengine_b200 as me
    src = """
__device__ double me_user_energy(const double* x, const double* cr, const double* ci, const double* k) {
    const double a2 = x[0] * x[0];
    double quad = 0.0, tot = 0.0;
    for (int j = 0; j < ME_NC; j++) {
        const double q = (double)(j - ME_NC / 2);
        const double m2 = cr[j] * cr[j] + ci[j] * ci[j];
        quad = quad + (k[1] + (k[2] * (q * q)) * (1.0 + a2)) * m2;
        tot = tot + m2;
    }
    return (k[0] * a2 + quad) + (k[3] / (2.0 * (double)ME_NC)) * (tot * tot);
}
__device__ bool me_user_reject(const double* x, const double* cr, const double* ci, const double* k) { return fabs(x[0]) >= 1.0; }
"""
    nc, n = 20, 300
    consts = (10.0, -1.0, 0.05, 1.0)
    kw = dict(initial_real_params=np.array([0.3]), initial_complex_params=0.05 * np.exp(1j * np.arange(nc)), temp=.1,
              n_chains=n, seed=4, sampling_width=0.05)
    a = me.MetropolisEngine(me.BuiltinEnergy("cylinder", *consts, reject=True), **kw)
    b = me.MetropolisEngine(me.CudaEnergy(src, consts=consts, has_reject=True), **kw)
    assert a._generic and b._generic
    assert torch.equal(a.state, b.state)                      # initial energies
    for eng in (a, b):
        eng.run(60, 3)                                        # crosses n > 50: per-chain covariances and factors live
        for _ in range(4):
            eng.step_all()
        eng.measure()
    torch.cuda.synchronize()
    assert torch.equal(a.state, b.state)
    assert torch.equal(a.time_series(), b.time_series())
    assert 0.05 < a.acceptance_rate < 0.9


def test_magnitude_phase_large_shape_philox_matches_c_oracle():
    """Magnitude / phase moves of a runtime-shape engine (1 real + 16 complex: 33 words, me_generic.cu gk_propose /
    gk_accept with groups 3 / 4) against the C oracle on the same Philox streams.  Half of the coefficients start at
    zero modulus (cmath.polar's arg 0 = 0 branch)."""
    import metropolisengine_b200 as me
    from oracle import c_oracle as co
    consts = [10.0, -1.0, 0.05, 1.0]
    c0 = np.concatenate([np.full(8, 0.05 + 0.02j), np.zeros(8, dtype=complex)])
    eng = me.MetropolisEngine(me.BuiltinEnergy("cylinder", *consts, reject=True), initial_real_params=np.array([0.2]),
                              initial_complex_params=c0, temp=.1, sampling_width=0.3, n_chains=48, seed=31,
                              complex_sample_method="magnitude-phase")
    assert eng._generic
    oracles = {ch: co.CChain(1, 16, "cylinder", consts=consts, temp=.1, sampling_width=0.3,
                             x0=np.concatenate([[0.2], c0.real, c0.imag]), use_reject=True) for ch in (0, 17, 47)}
    step, nacc = 0, {ch: 0 for ch in oracles}
    for im in range(54):                                       # crosses n > 50: adapted C_jj enter the magnitude width
        for grp in (1, 3, 4, 3):
            if grp == 1:
                eng.step_real_group()
            elif grp == 3:
                eng.step_complex_group_magnitude()
            else:
                eng.step_complex_group_phase()
            for ch, o in oracles.items():
                acc, _ = o.run(1, 1, False, seed=31, chain_id=ch, step0=step, group=grp)
                nacc[ch] += int(acc.sum())
            step += 1
        eng.measure()
        for ch, o in oracles.items():
            o.run(1, 0, True, seed=31, chain_id=ch, step0=step)
    eng.check_status()
    st = eng.state.cpu().numpy()
    lay = eng._lay
    assert sum(nacc.values()) > 0
    for ch, o in oracles.items():
        assert st[lay.NACC, ch] == nacc[ch], ch
        assert close(st[:lay.WORDS - 2, ch], o.state[:lay.WORDS - 2], 2e-9), ch
    assert eng.step_counter == 1 + 54 * 2                       # the two magnitude halves per block count (ME:450)


def test_magnitude_phase_with_a_torch_callable_matches_the_device_functor():
    """Magnitude / phase moves on the unfused propose / torch-callable / accept path (me_propose and me_accept with
    groups 3 / 4) consume the same Philox words as the fused kernel: same decisions, same states."""
    import metropolisengine_b200 as me
    kw = dict(initial_real_params=np.array([0.3, 0.2, 0.1]), initial_complex_params=np.array([0.4 - 0.1j, -0.3 + 0.2j]),
              temp=.1, sampling_width=0.6, complex_sample_method="magnitude-phase", n_chains=64, seed=5)

    def energy(r, c):
        a = (c * c.conj()).real
        return ((1 - r) ** 2).sum(dim=1) + r[:, 0] * r[:, 1] * (-a + 0.5 * a * a).mean(dim=1)

    engines = [me.MetropolisEngine(me.CudaEnergy(USER_SOURCES["warm_3r2c"]), **kw), me.MetropolisEngine(energy, **kw)]
    for eng in engines:
        for _ in range(54):
            eng.step_real_group()
            assert eng.step_complex_group() is None           # magnitude move + phase redraw (ME:168-176)
            eng.step_complex_group_magnitude()
            eng.measure()
    a, b = engines
    assert torch.equal(a.accept_count_per_chain, b.accept_count_per_chain)
    assert int(a.accept_count_per_chain.sum()) > 0
    assert torch.allclose(a.state, b.state, rtol=1e-10, atol=1e-13)
    assert a.step_counter == b.step_counter
    assert torch.equal(a.time_series(), a.time_series()) and torch.allclose(a.time_series(), b.time_series(), rtol=1e-10,
                                                                            atol=1e-13)


def test_graph_replay_of_callable_steps_is_bit_identical():
    """graph_callable=True: the steps between two measures of a torch-callable engine are captured once in a CUDA graph
    (device-resident step / measure counters, me_device_counters) and replayed; states and series must equal the
    step-by-step path exactly, across measures (the gain depends on the measure counter) and several blocks."""
    import metropolisengine_b200 as me

    def energy(r, c):
        a = (c * c.conj()).real
        return ((1 - r) ** 2).sum(dim=1) + r[:, 0] * r[:, 1] * (-a + 0.5 * a * a).mean(dim=1)

    def run(graph):
        eng = me.MetropolisEngine(energy, initial_real_params=np.array([0.3, 0.2]),
                                  initial_complex_params=np.array([0.4 - 0.1j]), temp=.1, n_chains=512, seed=9,
                                  graph_callable=graph)
        eng.run(30, 4)
        eng.step_all()                    # an ordinary step after replays uses the handle's counters again
        eng.run(25, 4)                    # crosses n > 50
        torch.cuda.synchronize()
        return eng.state.clone(), eng.time_series().clone(), eng.launch_count
    s0, t0, l0 = run(False)
    s1, t1, l1 = run(True)
    assert torch.equal(s0, s1) and torch.equal(t0, t1)
    assert l1 < l0 / 2                    # one graph launch per block instead of two kernel launches per step


def test_python_reject_condition_on_a_device_functor_engine():
    """ME:142-146: a python predicate on an engine whose energy is a device functor.  The engine then steps unfused
    (me_propose -> me_energy_builtin -> predicate -> me_accept); the chain must equal the one whose wall sits inside
    the functor (same Philox slots, same decisions), on a fused shape and on a runtime-compiled user functor."""
    import metropolisengine_b200 as me
    kw = dict(initial_real_params=np.array([0.9]), initial_complex_params=np.zeros(8, dtype=complex), temp=.1,
              sampling_width=0.2, n_chains=64, seed=31)
    consts = (10.0, -1.0, 0.05, 1.0)
    a = me.MetropolisEngine(me.BuiltinEnergy("cylinder", *consts, reject=True), **kw)
    a.run(55, 4)
    b = me.MetropolisEngine(me.BuiltinEnergy("cylinder", *consts), **kw)
    b.set_reject_condition(lambda r, c: r[:, 0].abs() >= 1.0)
    b.run(55, 4)
    assert torch.equal(a.accept_count_per_chain, b.accept_count_per_chain)
    assert torch.allclose(a.state, b.state, rtol=1e-11, atol=1e-13)
    assert float(a.real_params_per_chain.abs().max()) < 1.0
    # constructor argument form (the reference drops it; honoured here) with a user CUDA functor
    src = """
__device__ double me_user_energy(const double* x, const double* cr, const double* ci, const double* k) {
    return x[0] * x[0] + x[1] * x[1];
}"""
    kw2 = dict(initial_real_params=np.array([0., 0.]), temp=.1, n_chains=64, seed=3)
    c = me.MetropolisEngine(me.CudaEnergy(src), reject_condition=lambda r, c_: r[:, 0] > 0.3, **kw2)
    c.run(60, 5)
    assert float(c.real_params_per_chain[:, 0].max()) <= 0.3
    d = me.MetropolisEngine(me.CudaEnergy(src + """
__device__ bool me_user_reject(const double* x, const double* cr, const double* ci, const double* k) { return x[0] > 0.3; }
""", has_reject=True), **kw2)
    d.run(60, 5)
    assert torch.equal(c.accept_count_per_chain, d.accept_count_per_chain)
    assert torch.allclose(c.state, d.state, rtol=1e-11, atol=1e-13)


def test_reference_import_name_front_door_and_equilibrium_stats_with_external_frames():
    """`import metropolisengine as me` (reference README.md:15); adapt="pooled" picks the shared-covariance engine;
    save_equilibrium_stats(external_df=...) (ME:481-504) re-averages external frames from the global cut-off and
    leaves eq_means_error empty exactly like statistics.py:59,64."""
    import pandas
    import metropolisengine as ref_name
    import metropolisengine_b200 as me
    assert ref_name.MetropolisEngine is me.MetropolisEngine
    eng = ref_name.MetropolisEngine("x2", initial_real_params=[0.5], temp=.01, n_chains=4, seed=1)
    eng.run(300, 2)
    rows = 300
    ext = [pandas.DataFrame({"profile_a": np.linspace(0., 1., rows), "profile_b": np.ones(rows)}),
           pandas.DataFrame({"abs_a": np.arange(rows, dtype=float)}),
           pandas.DataFrame({"label": ["x"] * rows})]
    eq = eng.save_equilibrium_stats(external_df=ext)
    assert "profile_a" in eq and "abs_a" in eq and "label" not in eq and "profile_b" not in eq   # constant / string columns
    cut = eng.global_eq_point
    assert eng.eq_means_error == {}
    assert eng.field_profile["profile_a"] == pytest.approx(np.linspace(0., 1., rows)[cut:].mean())
    assert eng.field_abs_profile["abs_a"] == pytest.approx(np.arange(rows)[cut:].mean())
    assert eng.equilibrated_means["global_cutoff"] == cut and "param_0" in eng.equilibrated_means
    pooled = me.MetropolisEngine(me.BuiltinEnergy("cylinder", 10.0, -1.0, 0.05, 1.0, reject=True),
                                 initial_real_params=np.array([0.1]), initial_complex_params=np.zeros(64, dtype=complex),
                                 temp=.1, n_chains=256, seed=2, adapt="pooled")
    assert isinstance(pooled, me.SharedCovarianceEngine)
    pooled.run(2, 3)
    assert pooled.complex_mean.shape == (64,)
    with pytest.raises(NotImplementedError):
        me.MetropolisEngine("x2", initial_real_params=[0.0], temp=.1, adapt="pooled")
    lst = me.MetropolisEngine("x2", initial_real_params=[0.0], temp=.1, sampling_width=[0.3, 0.1])
    assert lst.real_group_sampling_width == 0.3
    with pytest.raises(AttributeError):
        lst.sampling_width
