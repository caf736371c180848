"""Workload energy definitions in the reference's plugin form (TEST INFRASTRUCTURE).

This file is part of ``oracle/`` — checker code only.  Nothing in the product package
(``metropolisengine_b200/``) may import it; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / reference arm do.

Every function has the reference's plugin signature
``(real_params: ndarray[n_r] f64, complex_params: ndarray[n_c] c128) -> float``
(reference contract: metropolisengine/metropolis_engine.py:20, call site :250 / :231).

The definitions follow SURVEY.md §8(d):

* ``x2``        README minimal example, ``E = x**2``                     (README.md:26-27)
* ``xy_well``   demo 1, ``E = const*(x**2+y**2)``                        (demo/toymodel_xypotentialwell.py:13-18)
* ``demo_2r1c`` demo 2 with a real-valued result                         (demo/toymodel_complex_and_real.py:17-26)
* ``mixed_3r4c`` bounded 3 real + 4 complex scale-up of demo 2           
* ``cylinder``  synthetic 1 real + n_c complex Fourier-mode field energy  (shape from
                the legacy API metropolis_engine.py:103,139-143)

The operation order in each function is bit-relevant: the CUDA functors in
``metropolisengine_b200/csrc/energies.cuh`` restate the same order.
"""
import numpy as np


def x2(real_params, complex_params):
    x = real_params[0]
    return x * x


def make_xy_well(const=1.0):
    def xy_well(real_params, complex_params):
        x = real_params[0]
        y = real_params[1]
        return const * (x * x + y * y)
    return xy_well


xy_well = make_xy_well(1.0)


def make_demo_2r1c(k=1.0, alpha=-1.0, beta=0.5):
    def demo_2r1c(real_params, complex_params):
        x = real_params[0]
        y = real_params[1]
        c = complex_params[0]
        a = c.real * c.real + c.imag * c.imag
        ex = 1.0 - x
        ey = 1.0 - y
        area = k * (ex * ex) + k * (ey * ey)
        field = (x * y) * (alpha * a + beta * (a * a))
        return area + field
    return demo_2r1c


demo_2r1c = make_demo_2r1c()


def make_mixed_3r4c(k=1.0, alpha=-1.0, beta=0.5, bounded=False):
    """E = k*sum_i (1-x_i)^2 + w*(1/n_c)*sum_j (alpha*|c_j|^2 + beta*|c_j|^4); sequential sums; w = x0*x1 (the
    demo's form, unbounded below where x0*x1 < 0) or |x0*x1| when ``bounded``."""
    def mixed_3r4c(real_params, complex_params):
        area = 0.0
        for i in range(len(real_params)):
            e = 1.0 - real_params[i]
            area = area + k * (e * e)
        s = 0.0
        n_c = len(complex_params)
        for j in range(n_c):
            c = complex_params[j]
            a = c.real * c.real + c.imag * c.imag
            s = s + (alpha * a + beta * (a * a))
        w = real_params[0] * real_params[1]
        return area + (abs(w) if bounded else w) * (s / n_c)
    return mixed_3r4c


mixed_3r4c = make_mixed_3r4c()
mixed_3r4c_bounded = make_mixed_3r4c(bounded=True)


def make_cylinder(n_c, kappa=10.0, alpha=-1.0, gamma=0.05, beta=1.0):
    """E = kappa*a^2 + sum_k (alpha + gamma*k^2*(1+a^2))*|c_k|^2 + (beta/(2 n_c))*(sum_k |c_k|^2)^2,
    modes k = -n_c/2 .. n_c/2-1 in storage order; sequential sums."""
    def cylinder(real_params, complex_params):
        a = real_params[0]
        a2 = a * a
        quad = 0.0
        tot = 0.0
        for j in range(n_c):
            kk = float(j - n_c // 2)
            c = complex_params[j]
            m2 = c.real * c.real + c.imag * c.imag
            quad = quad + (alpha + (gamma * (kk * kk)) * (1.0 + a2)) * m2
            tot = tot + m2
        return (kappa * a2 + quad) + (beta / (2.0 * n_c)) * (tot * tot)
    return cylinder


def cylinder_reject(real_params, complex_params):
    """Hard wall of the cylinder app: proposals with |a| >= 1 are rejected before the energy is
    evaluated (legacy metropolis_engine.py:103,139; engine hook metropolis_engine.py:142-146,247)."""
    return bool(abs(real_params[0]) >= 1.0)
