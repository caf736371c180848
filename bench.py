#!/usr/bin/env python
"""bench.py — ensemble chain-steps/sec of the step_all()/measure() hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3]

Workload (N=1 and per rank for N>1, weak scaling): BASELINE.json configs[1] — demo/toymodel_xypotentialwell:
2 real params, E = x^2+y^2, T = 0.1, 65,536 chains x 10^5 steps, measure every 10 steps (DXY:39-44), FP64.
One bench "step" = ONE pass of that whole job: 10^4 x (10 x step_all() + measure()) for every chain, i.e.
6.5536e9 chain-steps and 10^4 time-series rows per chain (21 GB written to HBM per pass, larger than L2).

value  : chain-steps/s, device-timed (CUDA events on the launching stream, barrier + synchronize on both sides,
         max over ranks), state resident in HBM.
e2e    : the same metric through the public Python API with HOST buffers: per step the initial parameters of all
         chains are copied from pinned host memory (H2D), the job runs, and the final per-chain state, the pooled
         statistics and chain 0's time series are read back to the host (D2H), all inside the timed region.
roofline: the step kernel is FP64-pipe bound: `achieved` = algorithmic flops (20 flop per
         chain-step for the xy-well, special functions not counted) / kernel time, `peak` = FP64 FMA throughput
         measured live by me_probe_fp64 (MEASURED_PEAKS.json has no FP64 figure); the HBM side (time-series bytes
         / kernel time vs the measured copy bandwidth) is reported next to it.
cpu_baseline / --impl reference: the numpy port of the reference's own loop (oracle/py_port.py, pinned
         bit-for-bit against the reference) on the host cores, one chain per process.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic work per chain-step / bytes per chain-measure (x[D], E, sigma stored in FP64).
# fp64_inst / wide_inst: FP64 and IMAD.WIDE instructions per chain-step in the SASS of the step loop
# (tests/scripts/sass_loop.py on the shipped library) — the inputs of the pipe-level roofline below.
WORKLOADS = {
    "c1": dict(name="README x^2: 1 real param, T=0.01, measure every step", energy=("x2",), n_r=1, n_c=0, temp=0.01,
               chains=65536, measures=10000, spm=1, flop=10, sf=3, fp64_inst=52.5, wide_inst=18.0),
    "c2": dict(name="demo/toymodel_xypotentialwell: 2 real params, E=x^2+y^2, T=0.1, 65,536 chains x 1e5 steps, "
                    "measure every 10", energy=("xy_well", 1.0), n_r=2, n_c=0, temp=0.1, chains=65536,
               measures=10000, spm=10, flop=20, sf=5, fp64_inst=59.0, wide_inst=18.0),
    "c3": dict(name="mixed 3 real + 4 complex (bounded demo-style well), T=0.1, 262,144 chains, measure every 10",
               energy=("mixed_well", 1.0, -1.0, 0.5, 1.0), n_r=3, n_c=4, temp=0.1, chains=262144, measures=100, spm=10,
               flop=170, sf=23, fp64_inst=328.0, wide_inst=94.0),
}


def _ncu_traffic_per_chain_measure():
    """DRAM bytes per chain-measure of the step kernel from the committed ncu --set full capture
    (profiles/r01_ncu_c2_k_run_v5.csv: a launch of 300 measures x 65,536 chains)."""
    try:
        rd = wr = None
        for line in open(os.path.join(ROOT, "profiles", "r01_ncu_c2_k_run_v5.csv")):
            f = line.strip().split(",")
            if f[0] == "dram__bytes_read.sum":
                rd = float(f[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[1]]
            if f[0] == "dram__bytes_write.sum":
                wr = float(f[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[1]]
        return (rd + wr) / (300.0 * 65536.0)
    except Exception:
        return None


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


# ------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
            except Exception:
                pass
            self.h = None
            if uuid:
                for cand in ("GPU-" + uuid, uuid):
                    try:
                        self.h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if hasattr(cand, "encode") else cand)
                        break
                    except Exception:
                        continue
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:                               # pragma: no cover
            self.nv, self.err = None, repr(e)

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------ CPU baseline
def _cpu_worker(args):
    """One process = one chain of the reference's loop (numpy port), timed for `seconds` after a warm-up."""
    wl_key, seed, seconds, fixed_steps = args
    os.environ["OMP_NUM_THREADS"] = "1"
    import random
    import numpy as np
    from oracle import energies as en
    from oracle.py_port import PortChain
    wl = WORKLOADS[wl_key]
    np.random.seed(seed)
    random.seed(seed)
    if wl_key == "c1":
        ch = PortChain(en.x2, initial_real_params=[0.0], temp=wl["temp"])
    elif wl_key == "c2":
        ch = PortChain(en.xy_well, initial_real_params=np.array([0., 0.]), temp=wl["temp"])
    else:
        ch = PortChain(en.mixed_3r4c_bounded, initial_real_params=np.zeros(3), initial_complex_params=np.zeros(4, dtype=complex),
                       temp=wl["temp"])
    spm = wl["spm"]
    for _ in range(20):                       # warm-up
        for _ in range(spm):
            ch.step()
        ch.measure()
    steps = 0
    series = []                                # first coordinate after every step (for the ESS estimate)
    first = (lambda: float(ch.real_params[0])) if ch.n_r else (lambda: float(ch.complex_params[0].real))
    t0 = time.perf_counter()
    if fixed_steps:
        for _ in range(fixed_steps // spm):
            for _ in range(spm):
                ch.step()
                series.append(first())
            ch.measure()
        steps = (fixed_steps // spm) * spm
    else:
        while time.perf_counter() - t0 < seconds:
            for _ in range(spm):
                ch.step()
                series.append(first())
            ch.measure()
            steps += spm
    dt = time.perf_counter() - t0
    from oracle.py_port import statistical_inefficiency
    g = statistical_inefficiency(series[len(series) // 5:]) if len(series) > 2000 else float("nan")
    return steps, dt, g


def cpu_baseline(wl_key, seconds=10.0, fixed_steps=0, procs=None):
    import multiprocessing as mp
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(wl_key, s, seconds, fixed_steps) for s in range(procs)])
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    gs = sorted(r[2] for r in res if r[2] == r[2])
    cpu_baseline.last_g = gs[len(gs) // 2] if gs else float("nan")
    return steps, busy, wall, procs


def c_oracle_rate(wl_key, steps=2_000_000):
    """Single-core throughput of the C restatement (Philox mode) — a much stronger CPU figure than the python loop."""
    import numpy as np
    from oracle import c_oracle as co
    wl = WORKLOADS[wl_key]
    n_r, n_c = wl["n_r"], wl["n_c"]
    o = co.CChain(n_r, n_c, wl["energy"][0], consts=wl["energy"][1:], temp=wl["temp"], x0=np.zeros(n_r + 2 * n_c))
    spm = wl["spm"]
    o.run(100, spm, True, seed=1, chain_id=0)
    t0 = time.perf_counter()
    o.run(steps // spm, spm, True, seed=1, chain_id=0, step0=100 * spm)
    return (steps // spm) * spm / (time.perf_counter() - t0)


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    procs = os.cpu_count() or 1
    per_step = 100 * wl["spm"] if args.workload != "c1" else 1000    # chain-steps per process per bench step
    for _ in range(args.warmup):
        cpu_baseline(args.workload, fixed_steps=per_step, procs=procs)
    total, t_sum = 0, 0.0
    for _ in range(args.steps):
        steps, busy, wall, _ = cpu_baseline(args.workload, fixed_steps=per_step, procs=procs)
        total += steps
        t_sum += busy
    value = total / t_sum
    sample = "%d processes x %d chain-steps of the %s schedule per bench step (one chain per process, " \
             "OMP_NUM_THREADS=1)" % (procs, per_step, args.workload)
    line = {
        "impl": "reference", "metric": "ensemble chain-steps/sec", "value": value, "unit": "chain-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_sum / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "arm": "numpy port of the reference loop (oracle/py_port.py), host cores"},
        "cpu_baseline": {"value": value, "unit": "chain-steps/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import ctypes
    import metropolisengine_b200 as me
    from metropolisengine_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the hot path is CUDA-only (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    wl = WORKLOADS[args.workload]
    chains, M, spm = wl["chains"], wl["measures"], wl["spm"]
    if args.measures:
        M = args.measures
    n_r, n_c = wl["n_r"], wl["n_c"]
    d = n_r + 2 * n_c
    ts_cols = d + (3 if (n_r and n_c) else 2)
    ts_bytes = M * ts_cols * chains * 8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    kw = dict(temp=wl["temp"], n_chains=chains * world, seed=2024, distributed=(world > 1), device=dev,
              ts_chunk_bytes=ts_bytes)
    if n_r:
        kw["initial_real_params"] = np.zeros(n_r)
    if n_c:
        kw["initial_complex_params"] = np.zeros(n_c, dtype=complex)
    eng = me.MetropolisEngine(wl["energy"], **kw)
    eng.reserve_rows(M)
    lay = eng._lay
    stream = torch.cuda.current_stream(dev)
    pooled_dev = torch.zeros(max(lay.POOL_WORDS, 1), dtype=torch.float64, device=dev)

    def one_pass():
        """One pass of the job on resident state: fused step/measure launch + pooled-moment reduction
        (+ its all-reduce across ranks, the path's only collective)."""
        eng.clear_time_series(keep_storage=True)
        eng.run(M, spm)
        eng._launch(eng._lib.me_pool_reduce(eng._h, ctypes.c_void_p(pooled_dev.data_ptr()), 1, eng._stream()))
        if world > 1:
            dist.all_reduce(pooled_dev)

    # ---- value: device-timed, inputs resident
    for _ in range(args.warmup):
        one_pass()
    barrier()
    launches0 = eng.launch_count
    ker_ev = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()          # the sampler's NVML start-up takes milliseconds and differs per process: line the ranks up
        ev0.record(stream) # again so that no rank's timed region contains another rank's start-up
        for _ in range(args.steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.clear_time_series(keep_storage=True)
            a.record(stream)
            eng.run(M, spm)
            b.record(stream)
            ker_ev.append((a, b))
            eng._launch(eng._lib.me_pool_reduce(eng._h, ctypes.c_void_p(pooled_dev.data_ptr()), 1, eng._stream()))
            if world > 1:
                dist.all_reduce(pooled_dev)
        ev1.record(stream)
        barrier()
    launches = eng.launch_count - launches0
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    chain_steps_per_pass = chains * world * M * spm
    value = chain_steps_per_pass * args.steps / (ms_total * 1e-3)
    ker_ms = sum(a.elapsed_time(b) for a, b in ker_ev) / len(ker_ev)
    # per-rank view of the timed region (a straggler GPU — e.g. one that is power-capped when all eight run FP64 flat
    # out — sets the job's time through the per-step all-reduce): kernel time and clocks of every rank
    rank_view = None
    if world > 1:
        mine = {"rank": rank, "kernel_ms": ker_ms, "clocks": clocks.summary()}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        rank_view = gathered

    # ---- e2e: public API with host buffers, copies inside the timed region
    pin_r = torch.zeros((chains, n_r), dtype=torch.float64).pin_memory() if n_r else None
    pin_c = torch.zeros((chains, n_c), dtype=torch.complex128).pin_memory() if n_c else None
    host_state = torch.empty((lay.WORDS, chains), dtype=torch.float64).pin_memory()
    host_ts0 = torch.empty((M, ts_cols), dtype=torch.float64).pin_memory()
    h2d = (pin_r.numel() * 8 if n_r else 0) + (pin_c.numel() * 16 if n_c else 0)
    d2h = host_state.numel() * 8 + host_ts0.numel() * 8 + (lay.POOL_WORDS + 1) * 8

    def e2e_pass():
        eng.reset(initial_real_params=pin_r, initial_complex_params=pin_c)        # H2D + re-initialisation
        eng.run(M, spm)
        ps = eng.pooled_statistics()                                               # reduce (+all-reduce) + D2H
        host_state.copy_(eng.state, non_blocking=True)                            # final per-chain state D2H
        host_ts0.copy_(eng.time_series()[:, :, 0], non_blocking=True)             # chain 0 series (df) D2H
        torch.cuda.synchronize(dev)
        return ps

    e2e_steps = max(1, min(args.steps, 3 if args.steps > 3 else args.steps))
    e2e_pass()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ps = e2e_pass()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = chain_steps_per_pass * e2e_steps / t.item()

    # ---- FP64 roofline denominator, measured live
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    probe_out = torch.empty(n_sm * 8 * 256, dtype=torch.float64, device=dev)
    flops = ctypes.c_int64()
    best = None
    for it in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        rc = eng._lib.me_probe_fp64(local_rank, 20000, ctypes.c_void_p(probe_out.data_ptr()), probe_out.numel(),
                                    eng._stream(), ctypes.byref(flops))
        b.record(stream)
        torch.cuda.synchronize(dev)
        assert rc == 0
        ms = a.elapsed_time(b)
        if it > 0:
            best = ms if best is None else min(best, ms)
    fp64_peak = flops.value / (best * 1e-3) / 1e12
    acc_rate = eng.acceptance_rate            # collective when sharded: every rank must call it
    tpm = _ncu_traffic_per_chain_measure() if args.workload == "c2" else None

    # ---- ESS: statistical inefficiency per STEP from a side ensemble measured at every step (device kernel)
    g_steps = None
    if rank == 0:
        kw2 = dict(kw)
        kw2.update(n_chains=4096, distributed=False, ts_chunk_bytes=4096 * ts_cols * 8 * 6000)
        side = me.MetropolisEngine(wl["energy"], **kw2)
        side.record = False
        side.run(300, 10)                     # sigma / covariance adaptation
        side.record = True
        side.run(6000, 1)
        gq = torch.stack([side.statistical_inefficiency(column=c, n_chains=1024, burn_in=0.2) for c in range(d)])
        g_steps = float(gq.max(dim=0).values.median().item())      # worst coordinate per chain, median over chains
        del side

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = _peaks() or {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    per_launch_steps = chains * M * spm
    achieved_tf = per_launch_steps * wl["flop"] / (ker_ms * 1e-3) / 1e12
    ts_gbs = ts_bytes / (ker_ms * 1e-3) / 1e9
    # Pipe-level roofline of the step kernel.  A warp-wide FP64 instruction holds the FP64 pipe of its SM
    # sub-partition for 2 cycles and an IMAD.WIDE (Philox round) the FMA-heavy pipe for ~3.5; the two contend
    # (tests/scripts/issue_mix.cu, profiles/r01_microbench_issue_mix.txt), so a warp-step costs at least
    # 2 * fp64_inst + 3.5 * wide_inst cycles of its sub-partition.  Launches are balanced over the sub-partitions
    # (several waves, or the work-queue time segmentation of me_device.cuh for ensembles of about one wave), so the
    # bound is that cost times the average number of warps per sub-partition.
    clk = clocks.summary()
    sm_hz = 1e6 * (clk.get("sm_mhz") or 1965)
    n_smsp = 4 * n_sm
    warps = -(-chains // 32)
    warps_busiest = warps / n_smsp
    cyc_step = 2.0 * wl["fp64_inst"] + 3.5 * wl["wide_inst"]
    pipe_bound_ms = 1e3 * warps_busiest * M * spm * cyc_step / sm_hz
    pipe = {"fp64_inst_per_chain_step": wl["fp64_inst"], "imad_wide_per_chain_step": wl["wide_inst"],
            "cycles_per_warp_step_lower_bound": cyc_step, "warps_per_subpartition": warps_busiest,
            "bound_ms": pipe_bound_ms, "frac": pipe_bound_ms / ker_ms,
            "how": "FP64-pipe + FMA-heavy-pipe cycles the SASS of the step loop needs (2 per FP64 instruction, 3.5 per "
                   "IMAD.WIDE; the pipes contend) x average warps per SM sub-partition / measured kernel time",
            "ncu": ("profiles/r01_ncu_c2_k_run_v5.csv: FP64 pipe 43 % + FMA-heavy pipe 45 % of elapsed cycles (the "
                    "FMA-heavy pipe also runs the IMAD.MOV / IMAD.SHL the compiler uses as moves), issue slots 60 % busy"
                    if args.workload == "c2" else None)}
    line = {
        "metric": "ensemble chain-steps/sec", "value": value, "unit": "chain-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "chains_per_gpu": chains, "steps_per_chain_per_bench_step": M * spm,
                   "measure_every": spm, "chain_steps_per_bench_step": chain_steps_per_pass,
                   "time_series_bytes_per_bench_step_per_gpu": ts_bytes,
                   "l2": "outputs (%.1f GB per step) exceed L2; state is register-resident" % (ts_bytes / 1e9),
                   "parallelism": "chains sharded over %d GPU(s); pooled-moment all-reduce per bench step" % world,
                   "launch": {"grid": eng._grid, "block": eng._block}},
        "e2e": {"value": e2e_value, "unit": "chain-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "what": "reset() from pinned host params, run(), pooled_statistics(), D2H of the "
                                            "final per-chain state and of chain 0's time series"},
        "gpu_launches": launches,
        "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": achieved_tf / fp64_peak,
                     "traffic": (tpm * chains * M) if tpm else None,
                     "traffic_note": "DRAM read+write bytes per launch: per chain-measure figure of the committed ncu "
                                     "--set full capture (profiles/r01_ncu_c2_k_run_v5.csv) x this launch's "
                                     "chain-measures; algorithmic %d B per chain-measure" % (8 * ts_cols),
                     "pipe": pipe,
                     "kernel": "me::k_run", "kernel_ms": ker_ms,
                     "algorithmic": "%d flop + %d special functions per chain-step (SURVEY.md §8d); special functions "
                                    "and Philox integer work are NOT counted in achieved" % (wl["flop"], wl["sf"]),
                     "peak_source": "me_probe_fp64 (FP64 FMA, measured in this run)",
                     "hbm": {"achieved": ts_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ts_gbs / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"}},
        "clocks": clocks.summary(),
        "ranks": rank_view,
        "check": {"acceptance_rate": acc_rate, "pooled_var_x0": float(ps["cov_real"][0, 0]) if n_r else None},
        "ess": {"g_steps": g_steps, "ess_per_sec": value / g_steps if g_steps else None,
                "how": "statistical inefficiency per step (pymbar's definition, worst coordinate, median of 1024 "
                       "chains, 6000 consecutive steps after adaptation), device kernel me_statistical_inefficiency"},
    }
    # ---- CPU baseline on this box's host cores (bounded sample)
    if not args.no_cpu:
        steps, busy, wall, procs = cpu_baseline(args.workload, seconds=args.cpu_seconds)
        line["cpu_baseline"] = {
            "value": steps / busy, "unit": "chain-steps/s", "cores": procs, "kind": "port",
            "sample": "%d processes x %.0f s of the same schedule, one chain each (numpy port of the reference loop, "
                      "oracle/py_port.py, OMP_NUM_THREADS=1)" % (procs, args.cpu_seconds),
            "c_oracle_1core": c_oracle_rate(args.workload),
            "g_steps": cpu_baseline.last_g,
            "ess_per_sec": (steps / busy) / cpu_baseline.last_g if cpu_baseline.last_g == cpu_baseline.last_g else None,
        }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_c4(args):
    """BASELINE config 4: cylinder-style field, 1 real + 64 complex parameters, shared proposal covariance,
    32,768 chains per GPU; tcgen05 path (csrc/me_k4.cu).  One bench step = 100 x (10 x step_all() + measure())
    including the pooled-covariance update at every measure."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import metropolisengine_b200 as me

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    chains, M, spm = 32768, (args.measures or 100), 10
    eng = me.SharedCovarianceEngine(energy_consts=(10.0, -1.0, 0.05, 1.0), temp=.1, n_chains=chains * world, seed=2024,
                                    record=False, distributed=(world > 1), device=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    eng.run(60, spm)                      # past the 50th measure: the shared covariance is live
    for _ in range(args.warmup):
        eng.run(M, spm)
    barrier()
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()          # see run_ours: the sampler's start-up must not leak into another rank's timed region
        ev0.record(stream)
        for _ in range(args.steps):
            eng.run(M, spm)
        ev1.record(stream)
        barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    value = chains * world * M * spm * args.steps / (ms * 1e-3)
    # the step kernel alone
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    eng.step(200)
    b.record(stream)
    torch.cuda.synchronize(dev)
    ker_ms = a.elapsed_time(b)
    acc = eng.acceptance_rate
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peaks = _peaks() or {}
    tf_peak = peaks.get("bf16_tflops", 1590.0)
    tf = chains * 200 * 2.0 * 128 * 128 / (ker_ms * 1e-3) / 1e12
    line = {
        "metric": "ensemble chain-steps/sec", "value": value, "unit": "chain-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64 state/energy/accept, bf16 proposal contraction (fp32 accumulate)",
        "data": "synthetic",
        "config": {"workload": "cylinder-style Fourier-mode field: 1 real + 64 complex, shared proposal covariance, "
                               "32,768 chains per GPU, measure + covariance adaptation every 10 steps",
                   "chains_per_gpu": chains, "measures_per_bench_step": M, "measure_every": spm},
        "gpu_launches": eng.launch_count - l0 - 1,
        "roofline": {"bound": "tensor", "achieved": tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf / tf_peak,
                     "traffic": None, "kernel": "k4_steps", "kernel_ms": ker_ms / 200,
                     "algorithmic": "2*128*128 flop per chain-step in the L.Z contraction (SURVEY.md §8d C4); the kernel "
                                    "is bound by in-kernel Gaussian generation (Philox + Box-Muller), not by the "
                                    "tensor pipe",
                     "step_kernel_only_chain_steps_per_s": chains * 200 / (ker_ms * 1e-3)},
        "clocks": clocks.summary(), "check": {"acceptance_rate": acc},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c4"])
    ap.add_argument("--measures", type=int, default=0, help="override measures per bench step (debug)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                                   # timing rule: W >= 3
    if args.workload == "c4":
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the c4 reference arm is not wired; use --workload c2"}))
            return 0
        return run_c4(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
