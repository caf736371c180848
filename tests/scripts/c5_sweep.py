"""the C2 and C3 ensembles at 10^6, 10^7 and 10^8 chains in total, sharded over the ranks of one
box (contiguous global chain ranges, Philox counters carry the global chain id), with the pooled-moment reduction and
its all-reduce after every launch.  Prints one line per (workload, chains): device-timed chain-steps/s (max over
ranks) and the pooled ensemble statistics with 15 digits; the statistics after the third launch cover the same
samples at every GPU count, so runs at 1 / 2 / 4 / 8 GPUs can be compared for GPU-count invariance.  Time series are not recorded (10^8 chains of rows do not fit and are not the point here).

  python tests/scripts/c5_sweep.py                                       # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      tests/scripts/c5_sweep.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import metropolisengine_b200 as me

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

SPM, MPL = 10, 10                 # steps per measure, measures per launch (one all-reduce per launch)
WORKLOADS = {
    "c2": dict(energy=("xy_well", 1.0), n_r=2, n_c=0, temp=0.1, rate=1.6e11),
    "c3": dict(energy=("mixed_well", 1.0, -1.0, 0.5, 1.0), n_r=3, n_c=4, temp=0.1, rate=2.6e10),
}
# "graph": the launches of a configuration go into CUDA graphs of 10 rounds [run -> reduction -> all-reduce], each round's
# collective beside the next round's stepping launch (run_graphed(launches=10)); the pooled statistics are read once at the
# end.  Default: eager launches with a host read-back of the pooled statistics after every launch (the round-1 protocol).
GRAPH = "graph" in sys.argv[1:]
only = [a for a in sys.argv[1:] if a in WORKLOADS] or list(WORKLOADS)
counts = [int(float(a)) for a in sys.argv[1:] if a not in WORKLOADS and a != "graph"] or [10**6, 10**7, 10**8]


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


for name in only:
    wl = WORKLOADS[name]
    for total in counts:
        kw = dict(temp=wl["temp"], n_chains=total, seed=2024, distributed=(world > 1), device=dev, record=False)
        if wl["n_r"]:
            kw["initial_real_params"] = np.zeros(wl["n_r"])
        if wl["n_c"]:
            kw["initial_complex_params"] = np.zeros(wl["n_c"], dtype=complex)
        try:
            eng = me.MetropolisEngine(wl["energy"], **kw)
        except torch.OutOfMemoryError as e:
            if rank == 0:
                print(json.dumps({"workload": name, "chains": total, "n_gpus": world, "skipped": "does not fit: %s"
                                  % str(e).split("\n")[0][:80]}), flush=True)
            continue
        # about 0.4 s of timed work per configuration at the single-GPU rate, at least 3 launches
        launches = max(3, int(0.4 * wl["rate"] * world / (total * SPM * MPL)))
        launches = min(launches, 200)
        eng.run(60, SPM)                     # past the covariance-adaptation threshold (n > 50)
        eng.pooled_statistics()              # first use: communicator and peer windows are set up outside the timed region
        eng.reset_pooled_statistics()
        stream = torch.cuda.current_stream(dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pooled = None
        if GRAPH:
            launches = max(10, launches // 10 * 10)
            eng.run_graphed(MPL, SPM, 10)             # eager first use
            eng.run_graphed(MPL, SPM, 10)             # capture + first replay
            eng.reset_pooled_statistics()
        barrier()
        ev0.record(stream)
        if GRAPH:
            for i in range(launches // 10):
                eng.run_graphed(MPL, SPM, 10)
            ev1.record(stream)
            pooled = pooled3 = eng.pooled_statistics()
        else:
            for i in range(launches):
                eng.run(MPL, SPM)
                pooled = eng.pooled_statistics()          # device reduction + the path's one all-reduce
                if i == 2:
                    pooled3 = pooled                      # same sample set whatever the GPU count: the invariance check
            ev1.record(stream)
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        eng.check_status()
        if rank == 0:
            mean = np.concatenate([np.asarray(pooled["mean_real"], dtype=np.float64),
                                   np.asarray(pooled["mean_complex"], dtype=np.complex128).view(np.float64)])
            cov = pooled["cov_real"]
            print(json.dumps({
                "workload": name, "chains": total, "n_gpus": world, "launches": launches, "mode": "graph" if GRAPH else "eager",
                "steps_per_launch": SPM * MPL, "ms": round(ms, 3),
                "chain_steps_per_s": total * launches * SPM * MPL / (ms * 1e-3),
                "pooled_samples": float(pooled.get("count", float("nan"))),
                "pooled_mean": ["%.15e" % v for v in mean[:4]],
                "pooled_cov_real_diag": ["%.15e" % v for v in np.diag(cov)],
                "after_3_launches": {"samples": float(pooled3["count"]),
                                     "mean_real": ["%.15e" % v for v in pooled3["mean_real"]],
                                     "cov_real_diag": ["%.15e" % v for v in np.diag(pooled3["cov_real"])]},
            }), flush=True)
        del eng
        torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
