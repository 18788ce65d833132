"""Static SASS summary of libspade_b200.so for profiles/ (run in the dev container):
registers, stack / spills, and the counts of memory, synchronisation and FP64 instructions per
kernel (cuobjdump -sass / -res-usage).

    python tools/sass_summary.py profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pyspade_b200", "libspade_b200.so")
KEYS = ["LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG", "RED", "LDGSTS", "UBLKCP", "SYNCS", "SHFL", "LDL", "STL",
        "DFMA", "DMUL", "DADD", "MUFU", "BAR", "WARPSYNC"]


def demangle(name):
    return subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip()


def main(dst):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    usage_txt = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage, cur = {}, None
    for line in usage_txt.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[demangle(cur)] = line.strip()
            cur = None
    rows = []
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        name = demangle(f.split("\n", 1)[0].strip())
        if "cub::" in name or "thrust::" in name:
            continue
        ops, n = collections.Counter(), 0
        for line in f.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                n += 1
                ops[m.group(2).split(".")[0]] += 1
        rows.append((name, n, ops))
    with open(dst, "w") as fh:
        fh.write("# cuobjdump -sass / -res-usage of pyspade_b200/libspade_b200.so (sm_100a): static instruction counts\n"
                 "# LDL / STL = local memory (spills, stack); UBLKCP = cp.async.bulk (TMA); SYNCS = mbarrier ops;\n"
                 "# LDGSTS = cp.async.  CUB / thrust kernels (layout scan, fallback member sort) omitted.\n")
        for name, n, ops in sorted(rows):
            short = re.sub(r"\((spade::)?(RunArgs|FinCtx|StageArgs|const|int|long|double|unsigned).*", "", name)
            fh.write(f"{short}\n    {usage.get(name, '')}\n    instructions={n} "
                     + " ".join(f"{k}={ops[k]}" for k in KEYS if ops[k]) + "\n")


if __name__ == "__main__":
    main(sys.argv[1])
