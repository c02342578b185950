"""SASS opcode census of libpfc_b200.so: per kernel, how many tcgen05 / TMEM / TMA / cluster instructions the compiled
sm_100a code contains (cuobjdump -sass; runs on the CPU-only build box).   python tools/sass_summary.py > profiles/rNN_sass_summary.txt
UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit -> mbarrier, LDTM = tcgen05.ld (TMEM -> registers),
UTCATOMSWS = tcgen05.alloc / dealloc, UTMALDG / UTMASTG = TMA bulk-tensor load / store, UTMACCTL.PF = tensor-map prefetch,
SYNCS = mbarrier ops, UCGABAR_ARV = cluster barrier, LDGSTS = cp.async, HMMA = legacy mma.sync (none expected),
DFMA = fp64 (scorer)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "face-recognition-pytorch_b200", "libpfc_b200.so")
OPS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UTMALDG", "UTMASTG", "UTMACCTL", "SYNCS", "UCGABAR_ARV", "LDGSTS",
       "HMMA", "DFMA", "MUFU.EX2", "REDG", "ATOMG", "ATOMS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = {}
    cur, counts, total = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            total[cur] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[T\d]+\s+)?([A-Z0-9_.]+)", line)
        if m:
            total[cur] += 1
            op = m.group(1)
            for o in OPS:
                if op == o or op.startswith(o + ".") or op.startswith(o + "_"):
                    counts[cur][o] += 1
    names = list(counts)
    try:
        dem = subprocess.run(["c++filt"] + names, stdout=subprocess.PIPE, text=True, check=True).stdout.splitlines()
        demangle = dict(zip(names, dem))
    except Exception:
        pass
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  ({len(names)} kernels, arch sm_100a)")
    print(f"{'kernel':78s} {'instr':>6s}  " + " ".join(f"{o:>8s}" for o in OPS))
    for n in names:
        d = re.sub(r"\(.*", "", demangle.get(n, n)).replace("pfc::", "").replace("void ", "")
        c = counts[n]
        print(f"{d[:78]:78s} {total[n]:6d}  " + " ".join(f"{c[o] or '.':>8}" for o in OPS))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print(f"{'TOTAL':78s} {sum(total.values()):6d}  " + " ".join(f"{tot[o] or '.':>8}" for o in OPS))


if __name__ == "__main__":
    main()
