"""Definitions of the single-chain golden cases (shared by make_golden.py, which feeds them to the
unmodified reference, and by the tests, which feed them to the oracle ports and to the CUDA engine).

Each case: energy plugin in the reference's form, schedule ``n_measures x (steps_per_measure steps + measure)``
as in README.md:39-44 / demo loops, seed for both global generators, optional hard-wall predicate,
constructor keywords (reference signature, metropolis_engine.py:17).

``builtin`` names the CUDA built-in functor (metropolisengine_b200/csrc/energies.cuh) that restates the same
energy with the same operation order, with its constants; ``None`` means "no built-in: use a user functor".
"""
import numpy as np

from oracle import energies as en


def _pure_2c(r, c):
    return float((abs(c[0]) ** 2 - 1.0) ** 2 + abs(c[1] - c[0]) ** 2)


def _warm_3r2c(r, c):
    a = (c * c.conjugate()).real
    return float(np.sum((1 - r) ** 2) + r[0] * r[1] * np.mean(-1 * a + .5 * a ** 2))


def _dict_terms():
    k, al, be = 1.0, -1.0, 0.5
    area = lambda r, c: k * (1 - r[0]) ** 2 + k * (1 - r[1]) ** 2
    field = lambda r, c: r[0] * r[1] * (al * (c[0] * c[0].conjugate()).real
                                        + be * (c[0] * c[0].conjugate()).real ** 2)
    return {"complex": {"field": field}, "real": {"field": field, "area": area},
            "all": {"field": field, "area": area}}


def cases():
    return {
        # README minimal example
        "kat1_x2": dict(energy=en.x2, builtin=("x2", []), n_measures=1000, steps_per_measure=1, seed=0,
                        ctor=dict(initial_real_params=[0.0], temp=.01)),
        # KAT2: demo 1, xy potential well
        "kat2_xy": dict(energy=en.xy_well, builtin=("xy_well", [1.0]), n_measures=1000, steps_per_measure=10, seed=0,
                        ctor=dict(initial_real_params=np.array([0., 0.]), temp=.1)),
        # KAT3: demo 2 (2 real + 1 complex), single real-valued callable
        "kat3_2r1c": dict(energy=en.demo_2r1c, builtin=("mixed_well", [1.0, -1.0, 0.5]), n_measures=100,
                          steps_per_measure=10, seed=0,
                          ctor=dict(initial_real_params=np.array([0., 0.]), initial_complex_params=np.array([0j]),
                                    temp=.1)),
        # BASELINE config 3 shape: 3 real + 4 complex; crosses n > 50 so both covariance recursions run
        "c3_3r4c": dict(energy=en.mixed_3r4c, builtin=("mixed_well", [1.0, -1.0, 0.5]), n_measures=120,
                        steps_per_measure=10, seed=0,
                        ctor=dict(initial_real_params=np.array([0., 0., 0.]),
                                  initial_complex_params=np.zeros(4, dtype=complex), temp=.1)),
        # all-complex engine: step_all is the complex-group step (metropolis_engine.py:46)
        "pure_2c": dict(energy=_pure_2c, builtin=None, n_measures=90, steps_per_measure=5, seed=3,
                        ctor=dict(initial_complex_params=np.array([0.3 + 0.1j, -0.2j]), temp=.2)),
        # hard wall via set_reject_condition (metropolis_engine.py:142-146,247-249), cylinder-shaped 1r+8c
        "cyl_1r8c_reject": dict(energy=en.make_cylinder(8), builtin=("cylinder", [10.0, -1.0, 0.05, 1.0]),
                                reject=en.cylinder_reject, n_measures=70, steps_per_measure=6, seed=5,
                                ctor=dict(initial_real_params=np.array([0.9]),
                                          initial_complex_params=np.zeros(8, dtype=complex), temp=.1,
                                          sampling_width=0.2)),
        # BASELINE config 4 shape with the reference's own per-chain algorithm: 1 real + 64 complex, 128x128 embedded
        # proposal covariance (metropolis_engine.py:274-302); crosses n > 50 so the 64x64 complex recursion runs
        "cyl_1r64c": dict(energy=en.make_cylinder(64), builtin=("cylinder", [10.0, -1.0, 0.05, 1.0]),
                          reject=en.cylinder_reject, n_measures=53, steps_per_measure=3, seed=13,
                          ctor=dict(initial_real_params=np.array([0.2]),
                                    initial_complex_params=np.zeros(64, dtype=complex), temp=.1,
                                    sampling_width=0.012)),
        # group-wise stepping of a mixed engine: step_real_group / step_complex_group alternate,
        # each with its own width (metropolis_engine.py:209-239, 440-456)
        "groups_2r1c": dict(energy=en.demo_2r1c, builtin=("mixed_well", [1.0, -1.0, 0.5]), n_measures=70,
                            steps_per_measure=8, seed=17, schedule="groups",
                            ctor=dict(initial_real_params=np.array([0.3, 0.2]),
                                      initial_complex_params=np.array([0.4 - 0.1j]), temp=.1)),
        # the same schedule started from sampling_width=[sigma_real, sigma_complex] (metropolis_engine.py:93-95): the
        # list form sets the two group widths only (the reference's step_all would raise AttributeError at ME:431)
        "groups_widths_2r1c": dict(energy=en.demo_2r1c, builtin=("mixed_well", [1.0, -1.0, 0.5]), n_measures=60,
                                   steps_per_measure=6, seed=19, schedule="groups",
                                   ctor=dict(initial_real_params=np.array([0.3, 0.2]),
                                             initial_complex_params=np.array([0.4 - 0.1j]), temp=.1,
                                             sampling_width=[0.21, 0.04])),
        # complex_sample_method="magnitude-phase" (metropolis_engine.py:129-130, 168-207, 304-317):
        # the user alternates step_real_group() and step_complex_group(), the latter being a Gaussian magnitude
        # move followed by a uniform phase redraw
        "magphase_2r1c": dict(energy=en.demo_2r1c, builtin=("mixed_well", [1.0, -1.0, 0.5]), n_measures=70,
                              steps_per_measure=9, seed=23, schedule="magphase",
                              ctor=dict(initial_real_params=np.array([0.3, 0.2]),
                                        initial_complex_params=np.array([0.4 - 0.1j]), temp=.1,
                                        complex_sample_method="magnitude-phase")),
        # the same schedule on a large parameter space (1 real + 16 complex = 33 words: runtime-shape kernels); half of
        # the coefficients start at zero modulus, where cmath.polar's argument is atan2 of signed zeros
        "magphase_1r16c": dict(energy=en.make_cylinder(16), builtin=("cylinder", [10.0, -1.0, 0.05, 1.0]),
                               reject=en.cylinder_reject, n_measures=54, steps_per_measure=6, seed=29,
                               schedule="magphase",
                               ctor=dict(initial_real_params=np.array([0.2]),
                                         initial_complex_params=np.concatenate([np.full(8, 0.05 + 0.02j),
                                                                                np.zeros(8, dtype=complex)]),
                                         temp=.1, sampling_width=0.3, complex_sample_method="magnitude-phase")),
        # temp = 0 (the constructor default): greedy descent, no uniform is ever drawn (metropolis_engine.py:331-332)
        "xy_temp0": dict(energy=en.xy_well, builtin=("xy_well", [1.0]), n_measures=60, steps_per_measure=4, seed=7,
                         ctor=dict(initial_real_params=np.array([1.0, -2.0]), temp=0)),
        # warm start: constructor-supplied covariance matrices and sampling width (metropolis_engine.py:63-70,93-99)
        "warm_3r2c": dict(energy=_warm_3r2c, builtin=None, n_measures=80, steps_per_measure=3, seed=11,
                          ctor=dict(initial_real_params=np.array([0.5, 0.5, 0.5]),
                                    initial_complex_params=np.array([0.1j, 0.2 + 0j]),
                                    covariance_matrix_real=np.array([[0.5, 0.2, 0.0], [0.2, 0.4, -0.1],
                                                                     [0.0, -0.1, 0.3]]),
                                    covariance_matrix_complex=np.array([[0.6, 0.1 + 0.2j], [0.1 - 0.2j, 0.5]],
                                                                       dtype=complex),
                                    sampling_width=0.11, temp=.3)),
        # dict-of-terms energy (metropolis_engine.py:111-115; demo/toymodel_complex_and_real.py:33-35) on a mixed
        # engine: the <term>_energy columns stay at their initial values
        "dict_2r1c": dict(energy=_dict_terms(), builtin=None, n_measures=60, steps_per_measure=10, seed=2,
                          ctor=dict(initial_real_params=np.array([0.2, 0.1]),
                                    initial_complex_params=np.array([0.5 + 0j]), temp=.1)),
    }


def fresh_ctor(case):
    """Deep-copied constructor kwargs (the reference mutates covariance arguments in place)."""
    out = {}
    for k, v in case["ctor"].items():
        out[k] = v.copy() if isinstance(v, np.ndarray) else (list(v) if isinstance(v, list) else v)
    return out
