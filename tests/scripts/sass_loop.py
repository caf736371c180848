"""Instruction histogram of the innermost hot loop of a kernel in libme_b200.so (CPU-side, cuobjdump).

usage: python tests/scripts/sass_loop.py <mangled-name-substring> [--dump]
The hot loop is the shortest backward-branch body that holds a full Philox call (>= 14 IMAD.WIDE: Philox4x32-7)."""
import collections, re, subprocess, sys

def main():
    key = sys.argv[1]
    lib = "metropolisengine_b200/lib/libme_b200.so"
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = out.split("Function : ")
    body = next(f for f in funcs if f.split("\n")[0].find(key) >= 0)
    ins = []
    for l in body.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(?:P\d,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            loop = [(x, y) for x, y in ins if lo <= x <= a]
            nw = sum("IMAD.WIDE" in y for _, y in loop)
            if nw >= 14 and (best is None or len(loop) < len(best[1])):
                best = (nw, loop)
    loop = best[1]
    # The exact-threshold fallback of the Metropolis test (me_device.cuh, finish_step: taken in ~2 steps of 10^5) sits
    # inside the loop behind a forward branch; it regenerates the accept bits (Philox) and ends in a DSETP.  The counts
    # that matter for the pipe model are those of the path a step normally executes, so such spans are left out.
    skipped = 0
    if "--all" not in sys.argv:
        keep, i = [], 0
        while i < len(loop):
            a, t = loop[i]
            keep.append((a, t))
            m = re.search(r"^@!?P\d BRA(?:\.U)?\s+0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) > a:
                tgt = int(m.group(1), 16)
                span = [(x, y) for x, y in loop if a < x < tgt]
                if sum("IMAD.WIDE" in y for _, y in span) >= 5 and any(y.split()[0].startswith("DSETP") or " DSETP" in y for _, y in span):
                    skipped += len(span)
                    i += len(span)
            i += 1
        loop = keep
    c = collections.Counter()
    for a, t in loop:
        toks = t.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        c[op.split(".")[0]] += 1
    fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DADD", "DMUL", "DSETP"))
    wide = sum("IMAD.WIDE" in t for _, t in loop)
    print(f"loop 0x{loop[0][0]:x}..0x{loop[-1][0]:x}: {len(loop)} instructions on the main path ({skipped} in rarely taken "
          f"fallback spans left out), {fp64} FP64, {wide} IMAD.WIDE")
    print(" ".join(f"{k}:{v}" for k, v in c.most_common()))
    if "--dump" in sys.argv:
        for a, t in loop:
            print(f"{a:05x} {t}")

main()
