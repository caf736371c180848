"""ctypes front-end of the C oracle (oracle/me_oracle.c).  TEST INFRASTRUCTURE ONLY — see the header of
me_oracle.c.  The product package never imports this module."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "me_oracle.c")
LIB = os.path.join(HERE, "libme_oracle.so")

ENERGY_IDS = {"x2": 0, "xy_well": 1, "mixed_well": 2, "cylinder": 3}
E_CALLBACK = 100
INJECT, PHILOX, XOSHIRO = 0, 1, 2

_ENERGY_CB = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int)
_REJECT_CB = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_int, ctypes.c_int)


class Config(ctypes.Structure):
    _fields_ = [("n_r", ctypes.c_int), ("n_c", ctypes.c_int), ("energy_id", ctypes.c_int),
                ("use_reject", ctypes.c_int), ("consts", ctypes.c_double * 16), ("temp", ctypes.c_double),
                ("target", ctypes.c_double), ("ratio", ctypes.c_double), ("m", ctypes.c_int),
                ("energy_cb", _ENERGY_CB), ("reject_cb", _REJECT_CB), ("frozen", ctypes.c_int)]


class Offsets(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int) for k in ("X", "E", "SIG", "MEAN", "COVR", "COVC", "OBSM", "FACR", "FACC",
                                             "NACC", "STATUS", "WORDS")]


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.meo_run.restype = ctypes.c_int
        _lib.meo_run_group.restype = ctypes.c_int
        _lib.meo_uniform.restype = ctypes.c_double
        _lib.meo_uniform.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int]
    return _lib


def layout(n_r, n_c):
    o = Offsets()
    lib().meo_layout(n_r, n_c, ctypes.byref(o))
    return o


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


class CChain:
    """One chain of the C oracle.  ``energy`` is a built-in name (with ``consts``) or a python callable
    ``f(x[d]) -> float`` over the flat parameter vector [real..., Re c..., Im c...]."""

    def __init__(self, n_r, n_c, energy, consts=(), temp=0.0, sampling_width=0.05, target_acceptance=0.3,
                 ratio=None, x0=None, cov_r=None, cov_c=None, use_reject=False, reject=None):
        from oracle.py_port import adaptation_constants
        self.n_r, self.n_c, self.d = n_r, n_c, n_r + 2 * n_c
        self.cfg = Config()
        self.cfg.n_r, self.cfg.n_c = n_r, n_c
        self._keep = []
        if isinstance(energy, str):
            self.cfg.energy_id = ENERGY_IDS[energy]
        else:
            self.cfg.energy_id = E_CALLBACK
            d = self.d
            cb = _ENERGY_CB(lambda xp, a, b: float(energy(np.ctypeslib.as_array(xp, shape=(d,)))))
            self._keep.append(cb)
            self.cfg.energy_cb = cb
        if reject is not None:
            d = self.d
            rcb = _REJECT_CB(lambda xp, a, b: int(bool(reject(np.ctypeslib.as_array(xp, shape=(d,))))))
            self._keep.append(rcb)
            self.cfg.reject_cb = rcb
            use_reject = True
        self.cfg.use_reject = int(use_reject)
        for i, v in enumerate(consts):
            self.cfg.consts[i] = float(v)
        self.cfg.temp = float(temp)
        self.cfg.target = float(target_acceptance)
        self.cfg.m = n_r + n_c
        self.cfg.ratio = float(ratio if ratio is not None else adaptation_constants(n_r, n_c, target_acceptance)[2])
        self.off = layout(n_r, n_c)
        self.state = np.zeros(self.off.WORDS, dtype=np.float64)
        self.n_measure = ctypes.c_int64(1)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        assert x0.shape == (self.d,)
        cr = None if cov_r is None else np.ascontiguousarray(cov_r, dtype=np.float64)
        cre = cim = None
        if cov_c is not None:
            cre = np.ascontiguousarray(np.real(cov_c), dtype=np.float64)
            cim = np.ascontiguousarray(np.imag(cov_c), dtype=np.float64)
        widths = np.atleast_1d(np.asarray(sampling_width, dtype=np.float64))
        lib().meo_init(ctypes.byref(self.cfg), _dp(self.state), _dp(x0), ctypes.c_double(widths[0]),
                       _dp(cr), _dp(cre), _dp(cim))
        if widths.size == 2:                       # sampling_width=[sigma_real, sigma_complex] (ME:93-95)
            self.state[self.off.SIG], self.state[self.off.SIG + 1] = widths[0], widths[1]

    def run(self, n_blocks, spm, do_measure=True, delta=None, u=None, seed=0, chain_id=0, step0=0, want_ts=False,
            group=0, generator="philox"):
        S = n_blocks * spm
        acc = np.zeros(max(S, 1), dtype=np.uint8)
        ts = np.zeros((max(n_blocks, 1), self.d + 3)) if want_ts else None
        if delta is not None:
            delta = np.ascontiguousarray(delta, dtype=np.float64)
            u = np.ascontiguousarray(u, dtype=np.float64)
            mode = INJECT
        else:
            mode = XOSHIRO if generator == "xoshiro" else PHILOX
        rc = lib().meo_run_group(ctypes.byref(self.cfg), _dp(self.state), mode, ctypes.c_int64(n_blocks),
                                 ctypes.c_int64(spm), int(do_measure), ctypes.byref(self.n_measure), _dp(delta), _dp(u),
                                 ctypes.c_uint64(seed), ctypes.c_uint64(chain_id), ctypes.c_uint64(step0),
                                 acc.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)), _dp(ts), int(group))
        assert rc == 0
        return acc[:S].astype(bool), ts

    # ---- unpacked views of the state
    @property
    def x(self):
        return self.state[self.off.X:self.off.X + self.d]

    @property
    def energy(self):
        return self.state[self.off.E]

    @property
    def sigma(self):
        return self.state[self.off.SIG:self.off.SIG + 2]

    @property
    def mean(self):
        return self.state[self.off.MEAN:self.off.MEAN + self.d]

    @property
    def obs_mean(self):
        return self.state[self.off.OBSM:self.off.OBSM + 2 * self.n_r + self.n_c]

    @property
    def cov_r(self):
        return unpack_sym(self.state[self.off.COVR:], self.n_r)

    @property
    def cov_c(self):
        return unpack_herm(self.state[self.off.COVC:], self.n_c)

    @property
    def fac_r(self):
        return unpack_sym(self.state[self.off.FACR:], self.n_r, lower_only=True)

    @property
    def fac_c(self):
        return unpack_herm(self.state[self.off.FACC:], self.n_c, lower_only=True)


def unpack_sym(w, n, lower_only=False):
    out = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            out[i, j] = w[i * (i + 1) // 2 + j]
            if not lower_only:
                out[j, i] = out[i, j]
    return out


def unpack_herm(w, n, lower_only=False):
    out = np.zeros((n, n), dtype=np.complex128)
    dg = n * (n - 1)
    for i in range(n):
        out[i, i] = w[dg + i]
        for j in range(i):
            p = 2 * (i * (i - 1) // 2 + j)
            out[i, j] = w[p] + 1j * w[p + 1]
            if not lower_only:
                out[j, i] = np.conj(out[i, j])
    return out
