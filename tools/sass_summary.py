"""Per-kernel counts of the SASS instructions that prove the Blackwell-native path (tcgen05 MMA, TMA loads/stores,
TMEM loads/stores) in the built library -> profiles/sass_summary.txt.

    python tools/sass_summary.py [out.txt]
Command: cuobjdump -sass pytorch_stable_diffusion_b200/csrc/libsdb200.so
Mnemonics (B200_PROFILING.md): UTCHMMA / UTCQMMA... = tcgen05.mma, UTMALDG = TMA load, UTMASTG = TMA store,
LDTM / STTM = tcgen05.ld / tcgen05.st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops,
MUFU.EX2 = ex2.approx."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorch_stable_diffusion_b200", "csrc", "libsdb200.so")
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "MUFU.EX2", "MUFU.TANH", "HMMA", "FFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    name, counts, order = None, collections.defaultdict(collections.Counter), []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            name = name.replace("void ", "")
            if name not in order:
                order.append(name)
            continue
        if name is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[name]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[name][k] += 1
    lines = ["# cuobjdump -sass pytorch_stable_diffusion_b200/csrc/libsdb200.so | tools/sass_summary.py",
             "# static instruction counts per kernel (sm_100a). UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor "
             "load / store,", "# LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, SYNCS = mbarrier, "
             "UBLKCP = cp.async.bulk", f"{'kernel':58s} {'instrs':>7s} " + " ".join(f"{k:>9s}" for k in KEYS)]
    for n in order:
        c = counts[n]
        lines.append(f"{n[:58]:58s} {c['_total']:7d} " + " ".join(f"{c[k]:9d}" for k in KEYS))
    text = "\n".join(lines) + "\n"
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_summary.txt")
    open(dst, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
