"""Static evidence of what the shipped library contains: for every kernel of libprb200.so the
registers / shared memory / spills (cuobjdump --dump-resource-usage) and counts of the SASS
mnemonics that matter for this path (bulk-copy engine, mbarrier, reductions, shared-memory byte
probes).  No GPU needed.  Usage: python profiles/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "picturedrocks_b200", "libprb200.so")
MNEMONICS = ("UBLKCP", "SYNCS", "RED.", "REDG", "ATOMS", "ATOMG", "LDS.U8", "LDS", "STS", "LDG", "STG",
             "REDUX", "VOTE", "SHFL", "DADD", "BAR.SYNC", "HMMA", "UTCMMA", "LDL", "STL")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout
    return dict(zip(names, out.splitlines()))


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"(prb::)?([A-Za-z0-9_]+)(<.*>)?\(", name)
    if not m:
        return name[:70]
    t = m.group(3) or ""
    return m.group(2) + (t if len(t) <= 60 else t[:57] + "...>")


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = dict(re.findall(r"(\w+):(\d+)", line))
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    total = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if cur and m:
            op = m.group(1)
            total[cur] += 1
            for k in MNEMONICS:
                if op.startswith(k):
                    counts[cur][k] += 1
    names = demangle(sorted(usage))
    arch = re.findall(r"arch = (sm_\w+)", subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True,
                                                        text=True).stdout + sass)
    print("# %s" % os.path.relpath(LIB, ROOT))
    print("# cubin architectures: %s" % ", ".join(sorted(set(arch))))
    print("# %d kernels; columns: registers, static shared bytes, local (stack) bytes, SASS instructions, "
          "then the mnemonic counts that are not zero" % len(usage))
    for k in sorted(usage, key=lambda n: short(names[n])):
        u = usage[k]
        c = counts[k]
        extra = " ".join("%s=%d" % (m, c[m]) for m in MNEMONICS if c[m])
        print("%-78s reg %3s smem %6s stack %4s sass %6d  %s" % (
            short(names[k]), u.get("REG"), u.get("SHARED"), u.get("STACK"), total[k], extra))
    spills = [short(names[k]) for k in usage if int(usage[k].get("STACK", 0)) > 0]
    print("# kernels with a stack frame (spills or local arrays): %s" % (", ".join(spills) or "none"))


if __name__ == "__main__":
    sys.exit(main())
