/* me_energies.cuh — built-in device energy functors (the plugin surface, reference contract
 * metropolis_engine.py:20: `[real values] [complex values] -> float`; hard-wall predicate
 * metropolis_engine.py:28,142-146).
 *
 * Functor interface (the same one a user functor compiled through NVRTC implements):
 *     static double eval  (const double* x [NR], const double* c_re [NC], const double* c_im [NC], const double* k);
 *     static bool   reject(same)                       - evaluated BEFORE the energy (metropolis_engine.py:247)
 * x / c_re / c_im point into the chain's register-resident parameter vector; `k` are the functor constants
 * (kernel parameter space).  Operation order matches oracle/energies.py, which the goldens were recorded with.
 * Fed to NVRTC as text: no #includes. */
#ifndef ME_ENERGIES_CUH
#define ME_ENERGIES_CUH

namespace me {

/* README minimal example E = x^2 (README.md:26-27) */
template <int NR, int NC>
struct EnergyX2 {
    __device__ static __forceinline__ double eval(const double *x, const double *, const double *, const double *) {
        return x[0] * x[0];
    }
    __device__ static __forceinline__ bool reject(const double *, const double *, const double *, const double *) {
        return false;
    }
};

/* demo 1: E = const (x^2 + y^2) (demo/toymodel_xypotentialwell.py:13-18); k[0] = const */
template <int NR, int NC>
struct EnergyXYWell {
    __device__ static __forceinline__ double eval(const double *x, const double *, const double *, const double *k) {
        return k[0] * (x[0] * x[0] + x[1] * x[1]);
    }
    __device__ static __forceinline__ bool reject(const double *, const double *, const double *, const double *) {
        return false;
    }
};

/* demo 2 and its scale-up (demo/toymodel_complex_and_real.py:17-26):
 * E = k0 sum_i (1-x_i)^2 + w (1/NC) sum_j (k1 |c_j|^2 + k2 |c_j|^4),  w = x0 x1 (the demo's form), or
 * w = |x0 x1| when k3 != 0.  The demo's form is unbounded below where x0 x1 < 0 (a few chains per thousand run
 * away, in the reference too); the |.| variant is the bounded workload used for large ensembles. */
template <int NR, int NC>
struct EnergyMixedWell {
    __device__ static __forceinline__ double eval(const double *x, const double *cr, const double *ci, const double *k) {
        double area = 0.0, s = 0.0;
#pragma unroll
        for (int i = 0; i < NR; i++) { const double e = 1.0 - x[i]; area = area + k[0] * (e * e); }
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const double a = cr[j] * cr[j] + ci[j] * ci[j];
            s = s + (k[1] * a + k[2] * (a * a));
        }
        const double w = x[0] * x[1];
        return area + (k[3] != 0.0 ? fabs(w) : w) * (s / (double)NC);
    }
    __device__ static __forceinline__ bool reject(const double *, const double *, const double *, const double *) {
        return false;
    }
};

/* cylinder-style Fourier-mode field (shape from legacy metropolis_engine.py:103,139-143):
 * E = k0 a^2 + sum_q (k1 + k2 q^2 (1+a^2)) |c_q|^2 + (k3/(2 NC)) (sum_q |c_q|^2)^2, q = j - NC/2;
 * hard wall |a| >= 1 */
template <int NR, int NC>
struct EnergyCylinder {
    __device__ static __forceinline__ double eval(const double *x, const double *cr, const double *ci, const double *k) {
        const double a2 = x[0] * x[0];
        double quad = 0.0, tot = 0.0;
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const double q = (double)(j - NC / 2);
            const double m2 = cr[j] * cr[j] + ci[j] * ci[j];
            quad = quad + (k[1] + (k[2] * (q * q)) * (1.0 + a2)) * m2;
            tot = tot + m2;
        }
        return (k[0] * a2 + quad) + (k[3] / (2.0 * (double)NC)) * (tot * tot);
    }
    __device__ static __forceinline__ bool reject(const double *x, const double *, const double *, const double *) {
        return fabs(x[0]) >= 1.0;
    }
};

/* placeholder for engines whose energy is evaluated by the caller between me_propose and me_accept
 * (torch-vectorised callable); the fused kernel is then only used for measure() */
template <int NR, int NC>
struct EnergyNone {
    __device__ static __forceinline__ double eval(const double *, const double *, const double *, const double *) {
        return 0.0;
    }
    __device__ static __forceinline__ bool reject(const double *, const double *, const double *, const double *) {
        return false;
    }
};

}  // namespace me

#endif
