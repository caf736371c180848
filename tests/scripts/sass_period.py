"""Instruction accounting of one MEASURE PERIOD of a fused kernel in libme_b200.so (CPU-side, cuobjdump): what a warp
executes per `spm` steps + one measure — the quantity the issue-port model of bench.py needs for workloads that measure
often (C1 measures after every step, so its measure block weighs as much as its step).

usage: python tests/scripts/sass_period.py <mangled-name-substring> <steps per measure> [--json]

Structure of the SASS of run_body (me_device.cuh): the measure loop is the outermost backward branch that encloses the
step loop; inside it sit (a) the per-measure prologue (Robbins-Monro gains), (b) the step loop — steps in pairs for shapes
with D <= 4, single steps otherwise —, (c) for paired shapes the odd trailing step behind a forward branch, (d) the measure
block.  Rarely taken fallback spans (exact Metropolis threshold: a forward branch over a span that regenerates Philox bits
and ends in a DSETP) are left out, as in sass_loop.py.  A period of `spm` steps executes
    (a) + (d) + spm x (single-step loop)                       D > 4
    (a) + (d) + (spm // 2) x (pair loop) + (spm % 2) x (c)     D <= 4
Weights of the issue-port model: 1 cycle per instruction, +1 per FP64 instruction, +2.5 per IMAD.WIDE, +15 per DMMA
(mma.m8n8k4.f64 = 256 FMA on 16 FP64 lanes per sub-partition)."""
import json
import re
import subprocess
import sys

FP64 = ("DFMA", "DADD", "DMUL", "DSETP")


def parse(key):
    lib = "metropolisengine_b200/lib/libme_b200.so"
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    body = next(f for f in out.split("Function : ") if f.split("\n")[0].find(key) >= 0)
    ins = []
    for l in body.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def op_of(t):
    toks = t.split()
    return toks[1] if toks[0].startswith("@") else toks[0]


def drop_fallbacks(span):
    keep, i = [], 0
    while i < len(span):
        a, t = span[i]
        keep.append((a, t))
        m = re.search(r"^@!?P\d BRA(?:\.U)?\s+0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) > a:
            tgt = int(m.group(1), 16)
            sub = [(x, y) for x, y in span if a < x < tgt]
            if sum("IMAD.WIDE" in y for _, y in sub) >= 5 and any("DSETP" in y for _, y in sub):
                i += len(sub)
        i += 1
    return keep


def count(span):
    span = drop_fallbacks(span)
    n = len(span)
    f = sum(op_of(t).split(".")[0] in FP64 for _, t in span)
    w = sum("IMAD.WIDE" in t for _, t in span)
    d = sum(op_of(t).startswith("DMMA") for _, t in span)
    return dict(inst=n, fp64=f, wide=w, dmma=d, cycles=n + f + 2.5 * w + 15 * d)


def main():
    key, spm = sys.argv[1], int(sys.argv[2])
    ins = parse(key)
    back = []
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(?:P\d,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            back.append((int(m.group(1), 16), a))
    rng = lambda lo, hi: [(x, y) for x, y in ins if lo <= x <= hi]
    wide_in = lambda lo, hi: sum("IMAD.WIDE" in y for _, y in rng(lo, hi))
    # step loop: the shortest backward-branch body holding a full Philox call
    step = min((b for b in back if wide_in(*b) >= 14), key=lambda b: b[1] - b[0])
    # measure loop: the shortest backward-branch body strictly enclosing the step loop
    outer = min((b for b in back if b[0] < step[0] and b[1] > step[1]), key=lambda b: b[1] - b[0])
    paired = "k_run_small" in key                 # shapes with D <= 4 run their steps in pairs (run_body: DEEP)
    pro = rng(outer[0], step[0] - 1)
    rest = rng(step[1] + 1, outer[1])
    odd = []
    if paired:
        # the odd trailing step: a uniform forward branch right after the pair loop skips it when spm is even
        for i, (a, t) in enumerate(rest[:8]):
            m = re.search(r"BRA(?:\.U)?\s+U?P\d,\s*0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) > a:
                tgt = int(m.group(1), 16)
                odd = [(x, y) for x, y in rest if a < x < tgt]
                rest = [(x, y) for x, y in rest if not (a < x < tgt)]
                break
    c_pro, c_step, c_odd, c_meas = count(pro), count(rng(*step)), count(odd), count(rest)
    per = {}
    for k in ("inst", "fp64", "wide", "dmma", "cycles"):
        if paired:
            per[k] = c_pro[k] + c_meas[k] + (spm // 2) * c_step[k] + (spm % 2) * c_odd[k]
        else:
            per[k] = c_pro[k] + c_meas[k] + spm * c_step[k]
    res = dict(kernel=key, spm=spm, paired=paired, prologue=c_pro, step_loop=c_step, odd_step=c_odd, measure=c_meas,
               period=per, per_step={k: per[k] / spm for k in per})
    if "--json" in sys.argv:
        print(json.dumps(res))
    else:
        print("measure loop 0x%x..0x%x, step loop 0x%x..0x%x (%s)" % (outer + step + ("steps in pairs" if paired else "single steps",)))
        for name in ("prologue", "step_loop", "odd_step", "measure", "period", "per_step"):
            print("%-10s" % name, res[name])


main()
