"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol include/me_b200.h
declares, its layout function agrees with the oracle's, and user functors compile through NVRTC."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT

HEADER = os.path.join(ROOT, "include", "me_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    from metropolisengine_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        ge.build()
    return _lib


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(me_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    L = lib.load()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), "libme_b200.so does not export %s" % n
        assert n in lib.SIGNATURES, "python binding lacks %s" % n
    assert sorted(lib.SIGNATURES) == names
    assert L.me_abi_version() == lib.ME_ABI_VERSION


def test_binding_structs_match_header_field_order(lib):
    text = open(HEADER).read()
    for cname, struct in (("me_config", lib.MeConfig), ("me_layout", lib.MeLayout), ("me_buffers", lib.MeBuffers),
                          ("me_k4_config", lib.MeK4Config)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), text, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = re.findall(r"\b(\w+)(?:\[\d+\])?;", body)
        assert fields == [f[0] for f in struct._fields_], cname


@pytest.mark.parametrize("nr,nc", [(1, 0), (2, 0), (2, 1), (3, 4), (0, 2), (1, 8), (1, 64)])
def test_state_layout_matches_oracle_layout(lib, nr, nc):
    from oracle import c_oracle as co
    lay = lib.layout(nr, nc)
    o = co.layout(nr, nc)
    for k in ("X", "E", "SIG", "MEAN", "COVR", "COVC", "OBSM", "FACR", "FACC", "NACC", "STATUS", "WORDS"):
        assert getattr(lay, k) == getattr(o, k), k
    d = nr + 2 * nc
    assert lay.D == d and lay.TS_COLS == d + (3 if (nr and nc) else 2)   # mixed engines record both group widths
    pw = d + d * (d + 1) // 2 + 2 * nr + nc
    assert lay.POOL_WORDS == (pw if pw <= 600 else 0)
    # unique persistent words 8 / 16 / 78 for C1-C3 (+ factors, accept count, status here)
    with pytest.raises(ValueError):
        lib.layout(0, 0)


def test_create_validates_like_the_reference(lib):
    L = lib.load()
    h = ctypes.c_void_p()
    cfg = lib.MeConfig(0, 0, 1, 0, 0.1, 0.3, 1.0, 0, 0, 0)
    assert L.me_create(ctypes.byref(cfg), ctypes.byref(h)) == lib.ME_ERR_INVALID      # ME:37-39
    assert b"ME:37" in L.me_last_error(None)
    cfg = lib.MeConfig(1, 0, 1, 0, -1.0, 0.3, 1.0, 0, 0, 0)
    assert L.me_create(ctypes.byref(cfg), ctypes.byref(h)) == lib.ME_ERR_INVALID      # ME:92
    cfg = lib.MeConfig(2, 0, 65536, 0, 0.1, 0.3, 3.49, 0, 0, 0)
    assert L.me_create(ctypes.byref(cfg), ctypes.byref(h)) == lib.ME_OK
    grid, block = ctypes.c_int32(), ctypes.c_int32()
    assert L.me_launch_dims(h, ctypes.byref(grid), ctypes.byref(block)) == lib.ME_OK
    assert grid.value * block.value >= 65536 and block.value % 32 == 0
    # nothing bound yet: the run entry points refuse instead of touching memory
    assert L.me_run(h, 1, 1, 0, None, 0, None) == lib.ME_ERR_STATE
    n, s = ctypes.c_int64(), ctypes.c_uint64()
    assert L.me_get_counters(h, ctypes.byref(n), ctypes.byref(s)) == lib.ME_OK and n.value == 1 and s.value == 0
    assert L.me_destroy(h) == lib.ME_OK


def test_user_functor_compiles_with_nvrtc_without_a_gpu(lib):
    src = """
__device__ double me_user_energy(const double* x, const double* cr, const double* ci, const double* k) {
    double s = k[0] * x[0] * x[0];
    for (int j = 0; j < ME_NC; j++) s += cr[j] * cr[j] + ci[j] * ci[j];
    return s;
}
__device__ bool me_user_reject(const double* x, const double* cr, const double* ci, const double* k) {
    return x[0] > 3.0;
}
"""
    for strict in (False, True):
        ok, log = lib.check_energy_source(src, 1, 3, use_reject=True, strict=strict)
        assert ok, log
    # small shape (D <= 4): compiled under the 160-register cap of the ahead-of-time small kernels
    src2 = "__device__ double me_user_energy(const double* x, const double* cr, const double* ci, const double* k) { return k[0] * (x[0] * x[0] + x[1] * x[1]); }"
    ok, log = lib.check_energy_source(src2, 2, 0, use_reject=False, strict=False)
    assert ok, log
    ok, log = lib.check_energy_source("__device__ double me_user_energy(const double* x) { return 0; }", 1, 0)
    assert not ok and "me_user_energy" in log
    # shapes that are not instantiated ahead of time compile at run time too (built-in functor, arbitrary shape)
    ok, log = lib.check_energy_source(None, 5, 2)
    assert ok, log


def test_engine_fails_loudly_without_cuda():
    import torch
    import metropolisengine_b200 as me
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        me.MetropolisEngine("x2", initial_real_params=[0.0], temp=.01)


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: the product may mention it in comments but never import, include, link
    or load it."""
    pkg = os.path.join(ROOT, "metropolisengine_b200")
    pat = re.compile(r"(^\s*(import|from)\s+oracle\b)|(#\s*include\s*[\"<][^\">]*oracle)|(libme_oracle)|(c_oracle)|(py_port)",
                     re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), "%s uses the oracle" % f


def test_set_group_validation(lib):
    """me_set_group (include/me_b200.h; ME:209-239 group steps, ME:168-207 magnitude / phase halves): argument checks
    need no GPU.  Groups 3 / 4 are available on every shape, large parameter spaces (D > 32) included."""
    L = lib.load()
    h = ctypes.c_void_p()
    cfg = lib.MeConfig(2, 0, 64, 0, 0.1, 0.3, 3.49, 0, 0, 0)                # all-real: no complex group of any kind
    assert L.me_create(ctypes.byref(cfg), ctypes.byref(h)) == lib.ME_OK
    assert L.me_set_group(h, 1) == lib.ME_OK
    for g in (2, 3, 4):
        assert L.me_set_group(h, g) == lib.ME_ERR_INVALID
    assert L.me_set_group(h, 5) == lib.ME_ERR_INVALID and L.me_set_group(h, -1) == lib.ME_ERR_INVALID
    assert L.me_destroy(h) == lib.ME_OK
    for n_r, n_c in ((3, 4), (1, 16), (0, 20)):                              # fused mixed, runtime-shape mixed / complex
        cfg = lib.MeConfig(n_r, n_c, 64, 0, 0.1, 0.3, 1.0, 0, 0, 0)
        assert L.me_create(ctypes.byref(cfg), ctypes.byref(h)) == lib.ME_OK
        for g in (0, 2, 3, 4):
            assert L.me_set_group(h, g) == lib.ME_OK, (n_r, n_c, g)
        assert L.me_set_group(h, 1) == (lib.ME_OK if n_r else lib.ME_ERR_INVALID)
        assert L.me_destroy(h) == lib.ME_OK


def test_reference_import_name_resolves_to_this_package():
    """`import metropolisengine as me` (reference README.md:15,33) is a one-file shim over metropolisengine_b200."""
    import metropolisengine
    import metropolisengine_b200
    assert metropolisengine.MetropolisEngine is metropolisengine_b200.MetropolisEngine
    assert metropolisengine.__file__.startswith(ROOT)


@pytest.mark.parametrize("nr,nc", [(27, 0), (0, 14)])
def test_large_fused_shapes_fit_the_static_shared_memory_budget(lib, nr, nc):
    """Fused shapes whose pooled-moment staging approaches 48 KB (POOLW 448..600) must still compile in the
    throughput build: the log table then stays in global memory instead of a 16 KB shared copy.  Compile-only (NVRTC,
    no GPU); the two shapes are the ones ptxas used to reject with 'uses too much shared data'."""
    lay = lib.layout(nr, nc)
    assert 447 < lay.POOL_WORDS <= 600
    ok, log = lib.check_energy_source(None, nr, nc, use_reject=False, strict=False)
    assert ok, log[-400:]


def test_user_functor_compiles_for_runtime_shapes(lib):
    """A user CUDA functor on a shape beyond the register-resident fused kernels (D > 32) is compiled around the
    runtime-shape kernels (me_generic.cuh) — same contract as for the fused kernels.  Compile-only (NVRTC, no GPU)."""
    src = """
__device__ double me_user_energy(const double* x, const double* cr, const double* ci, const double* k) {
    double e = k[0] * x[0] * x[0];
    for (int j = 0; j < ME_NC; j++) e += (1.0 + 0.01 * j) * (cr[j] * cr[j] + ci[j] * ci[j]);
    return e;
}
__device__ bool me_user_reject(const double* x, const double* cr, const double* ci, const double* k) { return fabs(x[0]) >= 1.0; }
"""
    ok, log = lib.check_energy_source(src, 1, 20, use_reject=True)
    assert ok, log
    ok, log = lib.check_energy_source(src, 1, 20, use_reject=False)          # the wall is optional
    assert ok, log
    ok, log = lib.check_energy_source(src.replace("k[0] *", "k[0]] *"), 1, 20, use_reject=True)
    assert not ok and "user_energy.cu" in log


def test_collective_entry_points_validate_their_arguments(lib):
    """The pooled-statistics collective and its peer-window set-up reject null handles / buffers without touching a GPU."""
    L = lib.load()
    null = ctypes.c_void_p(None)
    assert L.me_reduce_stats(null, null, 0, null) == lib.ME_ERR_INVALID
    assert L.me_accumulate_stats(null, null, null, null, null) == lib.ME_ERR_INVALID
    assert L.me_comm_allreduce(null, null, 0, null) == lib.ME_ERR_INVALID
    buf = (ctypes.c_ubyte * 64)()
    assert L.me_comm_peer_init(null, 1024, buf) == lib.ME_ERR_INVALID
    assert L.me_comm_peer_connect(null, buf) == lib.ME_ERR_INVALID
    assert L.me_comm_peer_enable(null, 1) == lib.ME_ERR_INVALID
    assert L.me_k4_accumulate_moments(null, null, null, null, null) == lib.ME_ERR_INVALID
