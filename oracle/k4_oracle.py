"""ctypes front-end of the shared-covariance oracle (oracle/me_oracle_k4.c) plus the pooled-covariance algebra in numpy.
TEST INFRASTRUCTURE ONLY — see the header of me_oracle_k4.c.  The product package never imports this module.

``K4Ensemble`` carries a whole ensemble (the shared covariance couples the chains): per-chain FP64 state stepped by the C
restatement, pooled moments / covariance / Cholesky factor / BF16 operand in numpy with the engine's one-measure lag
(metropolisengine_b200/engine_shared.py)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "me_oracle_k4.c")
LIB = os.path.join(HERE, "libme_oracle_k4.so")


class Config(ctypes.Structure):
    _fields_ = [("nc", ctypes.c_int), ("use_wall", ctypes.c_int), ("consts", ctypes.c_double * 16),
                ("temp", ctypes.c_double), ("target", ctypes.c_double), ("ratio", ctypes.c_double),
                ("groups", ctypes.c_int)]


class Layout(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int) for k in ("D", "X", "E", "SIG", "MEAN", "OBSM", "NOBS", "NACC", "STATUS", "WORDS")]


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.k4o_step.restype = ctypes.c_int
        _lib.k4o_step.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_double,
                                  ctypes.c_double, ctypes.c_int64]
        _lib.k4o_measure.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        _lib.k4o_init.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double]
        _lib.k4o_normals.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_void_p]
        _lib.k4o_scalars.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p]
        _lib.k4o_delta.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        _lib.k4o_bf16_from_double.restype = ctypes.c_float
        _lib.k4o_bf16_from_double.argtypes = [ctypes.c_double]
    return _lib


def bf16_from_double(a):
    """Round float64 values to BF16 in ONE rounding (nearest even); returns float32 holding the BF16 values."""
    a = np.asarray(a, dtype=np.float64)
    f = a.astype(np.float32)
    b = f.view(np.uint32).copy()
    tie = ((b & 0xffff) == 0x8000) & (f.astype(np.float64) != a)
    up = (f.astype(np.float64) < a) == ((b >> 31) == 0)           # move the magnitude towards |a|
    b[tie & up] += 1
    b[tie & ~up] -= 1
    r = b + 0x7fff + ((b >> 16) & 1)
    return (r & 0xffff0000).astype(np.uint32).view(np.float32)


def normal_table():
    """The generator's table, built independently of the library: BF16 bit patterns (uint16) of the half-normal quantiles
    Phi^-1(1/2 + (i + 1/2) / 8192), i < 4096, from scipy's inverse normal CDF."""
    from scipy.special import ndtri
    z = ndtri(0.5 + (np.arange(4096) + 0.5) / 8192.0)
    return (bf16_from_double(z).view(np.uint32) >> 16).astype(np.uint16)


def embed_factor(cov_c):
    """C_c = G G^H -> B[2i][2j] = Re G/sqrt2, B[2i][2j+1] = Im G/sqrt2, B[2i+1][2j] = -Im G/sqrt2, B[2i+1][2j+1] =
    Re G/sqrt2 (row n = output coordinate, column k = normal), as float64."""
    G = np.linalg.cholesky(cov_c)
    nc = G.shape[0]
    gr, gi = G.real * 0.70710678118654752440, G.imag * 0.70710678118654752440
    B = np.zeros((2 * nc, 2 * nc))
    B[0::2, 0::2] = gr
    B[0::2, 1::2] = gi
    B[1::2, 0::2] = -gi
    B[1::2, 1::2] = gr
    return B


class K4Ensemble:
    def __init__(self, nc, n_chains, consts, temp, ratio, seed=0, use_wall=True, target=0.3, sampling_width=0.05,
                 x0=None, chain_offset=0, async_refresh=True, groups=2):
        self.nc, self.n, self.seed, self.chain_offset = nc, n_chains, int(seed), int(chain_offset)
        self.cfg = Config()
        self.cfg.nc, self.cfg.use_wall = nc, int(use_wall)
        for i, v in enumerate(consts):
            self.cfg.consts[i] = float(v)
        self.cfg.temp, self.cfg.target, self.cfg.ratio = float(temp), float(target), float(ratio)
        self.cfg.groups = int(groups)
        self.L = Layout()
        lib().k4o_layout_for(nc, ctypes.byref(self.L))
        x0 = np.zeros(1 + 2 * nc) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
        self.shift = x0.copy()
        self.state = np.zeros((n_chains, self.L.WORDS))
        for ch in range(n_chains):
            lib().k4o_init(ctypes.byref(self.cfg), self.state[ch].ctypes.data, x0.ctypes.data, float(sampling_width))
        self.n_measure, self.step = 1, 0
        self.ztab = np.ascontiguousarray(normal_table())
        self.cov_c = np.eye(nc, dtype=np.complex128)
        self.cov_a = 1.0
        self.mom_n, self.mom_a, self.mom_a2 = 0.0, 0.0, 0.0
        self.mom_c = np.zeros(nc, dtype=np.complex128)
        self.mom_cc = np.zeros((nc, nc), dtype=np.complex128)
        self.async_refresh = bool(async_refresh)
        # factors: the one the step kernel reads now, the one it adopts at the next step launch after a further measure
        self.B_now = bf16_from_double(embed_factor(self.cov_c))
        self.s_a_now = 1.0
        self._in_flight = None      # (B, s_a) of the refresh launched at the last measure
        self._ready = None          # ... of the one before: adopted by the next step launch

    # ---- stream
    def normals(self, step):
        """BF16 operand values [chains, K] of one step (exactly what the kernel writes)."""
        K = 2 * self.nc
        zb = np.zeros((self.n, K), dtype=np.float32)
        for ch in range(self.n):
            lib().k4o_normals(self.seed, self.chain_offset + ch, step, K, self.ztab.ctypes.data, zb[ch].ctypes.data)
        return zb

    def scalars(self, step):
        za, u = np.zeros(self.n), np.zeros(self.n)
        a, b = ctypes.c_double(), ctypes.c_double()
        for ch in range(self.n):
            lib().k4o_scalars(self.seed, self.chain_offset + ch, step, self.ztab.ctypes.data, ctypes.byref(a), ctypes.byref(b))
            za[ch], u[ch] = a.value, b.value
        return za, u

    def delta(self, zb):
        """Increments of every chain from BF16 operand values zb [chains, K] with the factor in use; (delta, scale)."""
        N = 2 * self.nc
        B = np.ascontiguousarray(self.B_now, dtype=np.float32)
        d = np.zeros((self.n, N), dtype=np.float32)
        sc = np.zeros((self.n, N))
        zb = np.ascontiguousarray(zb, dtype=np.float32)
        for ch in range(self.n):
            lib().k4o_delta(N, B.ctypes.data, zb[ch].ctypes.data, d[ch].ctypes.data, sc[ch].ctypes.data)
        return d, sc

    # ---- schedule (engine_shared.py: step() adopts the refresh launched at least one measure ago)
    def begin_step_launch(self):
        if self._ready is not None:
            self.B_now, self.s_a_now = self._ready
            self._ready = None

    def end_step_launch(self):
        if self._in_flight is not None:
            self._ready, self._in_flight = self._in_flight, None

    def step_injected(self, delta, za, u):
        """One step of every chain with injected increments [chains, N] (float32), normals za and uniforms u."""
        delta = np.ascontiguousarray(delta, dtype=np.float32)
        acc = np.zeros(self.n, dtype=bool)
        for ch in range(self.n):
            acc[ch] = bool(lib().k4o_step(ctypes.byref(self.cfg), self.state[ch].ctypes.data, delta[ch].ctypes.data,
                                          float(za[ch]), float(u[ch]), float(self.s_a_now), self.n_measure))
        self.step += 1
        return acc

    def measure(self):
        self.n_measure += 1
        n, nc, L = self.n_measure, self.nc, self.L
        for ch in range(self.n):
            lib().k4o_measure(ctypes.byref(self.cfg), self.state[ch].ctypes.data, n)
        # pooled moments about the shift (me_k4_moments) and, from the 50th measure on (n > 50: ME:389,396), the factor
        a = self.state[:, L.X] - self.shift[0]
        c = (self.state[:, L.X + 1:L.X + 1 + nc] - self.shift[1:1 + nc]) \
            + 1j * (self.state[:, L.X + 1 + nc:L.X + 1 + 2 * nc] - self.shift[1 + nc:])
        self.mom_n += self.n
        self.mom_a += a.sum()
        self.mom_a2 += (a * a).sum()
        self.mom_c += c.sum(axis=0)
        self.mom_cc += c.T @ c.conj()
        if n > 50:
            N = self.mom_n
            small = (self.state[:, L.SIG].sum() / self.n) ** 2 / n
            self.cov_c = (self.mom_cc - np.outer(self.mom_c, self.mom_c.conj()) / N) / (N - 1) + small * np.eye(nc)
            self.cov_a = (self.mom_a2 - self.mom_a * self.mom_a / N) / (N - 1) + small
            new = (bf16_from_double(embed_factor(self.cov_c)), float(np.sqrt(self.cov_a)))
            if self.async_refresh:
                if self._in_flight is not None:
                    self._ready, self._in_flight = self._in_flight, None
                self._in_flight = new
            else:
                self.B_now, self.s_a_now = new
