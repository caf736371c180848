"""Run under torchrun (one rank per GPU): a sharded ensemble must equal the single-GPU ensemble bit for bit, and the
pooled statistics (the path's one all-reduce, NCCL) must agree on every rank.  Exit code 0 = ok."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import metropolisengine_b200 as me  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 6000 + 7          # deliberately not divisible by the world size
    kw = dict(initial_real_params=np.zeros(3), initial_complex_params=np.zeros(4, dtype=complex), temp=.1, seed=77)
    eng = me.MetropolisEngine(("mixed_well", 1.0, -1.0, 0.5, 1.0), n_chains=n, distributed=True, **kw)
    eng.run(70, 5)
    ps = eng.pooled_statistics()
    mean_attr = eng.real_mean                     # pooled over all ranks (all-reduce inside)
    cov_attr = eng.covariance_matrix_complex
    # gather the shards on rank 0
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([eng.n_chains], dtype=torch.int64, device="cuda"))
    sizes = [int(s.item()) for s in sizes]
    assert sum(sizes) == n
    pad = max(sizes)
    buf = torch.zeros((eng.state.shape[0], pad), dtype=torch.float64, device="cuda")
    buf[:, :eng.n_chains] = eng.state
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    ok = True
    if rank == 0:
        full = torch.cat([p[:, :s] for p, s in zip(parts, sizes)], dim=1)
        ref = me.MetropolisEngine(("mixed_well", 1.0, -1.0, 0.5, 1.0), n_chains=n, **kw)
        ref.run(70, 5)
        ok = ok and torch.equal(full, ref.state)
        rps = ref.pooled_statistics()
        for k in ("mean_real", "cov_real", "cov_complex", "observables_mean"):
            ok = ok and np.allclose(ps[k], rps[k], rtol=1e-11, atol=1e-13)
        ok = ok and ps["count"] == rps["count"] == 70 * n
        ok = ok and np.allclose(mean_attr, ref.real_mean, rtol=1e-12) and np.allclose(cov_attr, ref.covariance_matrix_complex, rtol=1e-12)
    # the same collective from inside a CUDA graph: [fused run -> reduce -> NCCL all-reduce (library) -> accumulate]
    kw2 = dict(initial_real_params=np.zeros(2), temp=.1, seed=78, record=False)
    g = me.MetropolisEngine(("xy_well", 1.0), n_chains=n, distributed=True, **kw2)
    for _ in range(4):
        g.run_graphed(25, 4)
    g.run_graphed(25, 4, launches=3)     # eager first use
    g.run_graphed(25, 4, launches=3)     # capture: three rounds per graph, all-reduce of round i beside the launch of round i + 1
    g.run_graphed(25, 4, launches=3)     # replay
    gps = g.pooled_statistics()
    if rank == 0:
        ref2 = me.MetropolisEngine(("xy_well", 1.0), n_chains=n, **kw2)
        for _ in range(13):
            ref2.run(25, 4)
        r2 = ref2.pooled_statistics()
        ok = ok and gps["count"] == r2["count"] == 13 * 25 * n
        for k in ("mean_real", "cov_real", "observables_mean"):
            ok = ok and np.allclose(gps[k], r2[k], rtol=1e-11, atol=1e-13)
    # the library's collective itself (me_comm_allreduce): one-shot sum over the NVLink peer windows when the ranks could open
    # each other's windows (else NCCL) — exact small-integer sums, back to back without host synchronisation, odd sizes,
    # and replayed from a CUDA graph
    import ctypes
    from metropolisengine_b200 import _lib, parallel
    lib = _lib.load()
    comm = parallel.library_comm(lib, local)
    peer_on = parallel.peer_windows_on(comm)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for size in (1, 13, 601, 8328 * 2, 32768):
        base = torch.arange(size, dtype=torch.float64, device="cuda")
        for rep in range(40):
            buf = base * (rank + 1) + rep
            assert lib.me_comm_allreduce(comm, ctypes.c_void_p(buf.data_ptr()), size, st) == 0
            want = base * (world * (world + 1) // 2) + rep * world
            ok = ok and torch.equal(buf, want)
    gbuf = torch.zeros(601, dtype=torch.float64, device="cuda")
    src = torch.arange(601, dtype=torch.float64, device="cuda") * (rank + 1)
    torch.cuda.synchronize()
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        gbuf.copy_(src)
        gs = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        assert lib.me_comm_allreduce(comm, ctypes.c_void_p(gbuf.data_ptr()), 601, gs) == 0
    for rep in range(5):
        cg.replay()
        torch.cuda.synchronize()
        ok = ok and torch.equal(gbuf, torch.arange(601, dtype=torch.float64, device="cuda") * (world * (world + 1) // 2))
    # shared-covariance path: the pooled covariance (all-reduced moments -> identical Cholesky on every rank) must be the
    # same on all ranks bit for bit, and the sampler must be healthy
    k4 = me.SharedCovarianceEngine(energy_consts=(10.0, -1.0, 0.05, 1.0), temp=.1, n_chains=128 * 6 * world, seed=3,
                                   record=False, n_complex=16, distributed=True, sampling_width=0.02)
    k4.run(56, 4)
    k4.synchronize_refresh()
    torch.cuda.synchronize()
    cov = torch.view_as_real(k4._cov_c).clone()
    clo, chi = cov.clone(), cov.clone()
    dist.all_reduce(clo, op=dist.ReduceOp.MIN)
    dist.all_reduce(chi, op=dist.ReduceOp.MAX)
    ok = ok and torch.equal(clo, chi) and bool(torch.isfinite(cov).all()) and int(k4._psd_status.item()) == 0
    ok = ok and 0.05 < k4.acceptance_rate < 0.9
    if rank == 0:
        print("peer windows (one-shot NVLink all-reduce):", "on" if peer_on else "off (NCCL)")
    # every rank must hold the same pooled numbers
    v = torch.tensor(np.concatenate([ps["mean_real"], np.diag(ps["cov_real"])]), device="cuda")
    lo, hi = v.clone(), v.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ok = ok and torch.equal(lo, hi)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("dist_invariance world=%d: %s" % (world, "OK" if flag.item() else "FAILED"))
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
