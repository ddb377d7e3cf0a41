"""Per-kernel SASS opcode histogram of libsrl_b200.so (run here, no GPU needed):

    python profiles/sass_histogram.py > profiles/r2_sass_histogram.txt

So that claims like "the scan kernels move tiles with TMA and hand them over through mbarriers" or "the pair loss kernel
gathers with 256-bit loads" can be checked without rebuilding: UTMALDG / UTMASTG = TMA tile loads / stores, SYNCS = mbarrier
operations, LDGSTS = cp.async, LDG.*256 / STG.*256 = 256-bit global accesses, ACQBULK / PREEXIT = programmatic dependent
launch (griddepcontrol.wait / launch_dependents), DADD / DMUL / DFMA = float64 pipe, MUFU = special-function unit."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "srl_b200", "libsrl_b200.so")
KEYS = ["UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "LDGSTS", "LDG.256", "STG.256", "LDG", "STG", "LDS", "STS", "ACQBULK", "PREEXIT",
        "DADD", "DMUL", "DFMA", "F2F", "MUFU", "SHFL", "BAR", "ATOM", "RED", "MEMBAR", "LDL", "STL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::|srl::|<unnamed>::", "", name)
            name = re.sub(r"\(.*", "", name)[:100]
            per[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name is not None:
            op = m.group(1)
            c = per[name]
            c["total"] += 1
            base = op.split(".")[0]
            c[base] += 1
            if base in ("LDG", "STG") and ".256" in op:
                c[base + ".256"] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: static SASS instruction counts per kernel (sm_100a)")
    print("# " + " ".join(["kernel", "total"] + KEYS))
    for k, c in per.items():
        cells = [f"{key}={c[key]}" for key in KEYS if c[key]]
        print(f"{k}: total={c['total']} " + " ".join(cells))
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("# whole library: " + " ".join(f"{key}={tot[key]}" for key in KEYS if tot[key]))


if __name__ == "__main__":
    main()
