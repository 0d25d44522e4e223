#!/usr/bin/env python
"""Static resource table of every kernel in libmole_b200.so from the ptxas -v log that build.sh keeps
(mole_b200/csrc/_obj/ptxas_mole_api.log): registers, stack frame, spill stores / loads, static shared memory.
No GPU needed.  Usage: python tools/ptxas_report.py [substring of the kernel name] [> profiles/rNN_ptxas_resources.txt]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG = os.path.join(ROOT, "mole_b200", "csrc", "_obj", "ptxas_mole_api.log")

ENTRY = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'")
PROPS = re.compile(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads")
USED = re.compile(r"Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?")


def parse(path=LOG):
    rows, cur = [], None
    for line in open(path):
        m = ENTRY.search(line)
        if m:
            cur = {"name": m.group(1), "stack": 0, "spill_st": 0, "spill_ld": 0, "regs": None, "smem": 0}
            rows.append(cur)
            continue
        if cur is None:
            continue
        m = PROPS.search(line)
        if m and cur["regs"] is None:
            cur["stack"], cur["spill_st"], cur["spill_ld"] = map(int, m.groups())
            continue
        m = USED.search(line)
        if m and cur["regs"] is None:
            cur["regs"] = int(m.group(1))
            cur["smem"] = int(m.group(2) or 0)
    names = "\n".join(r["name"] for r in rows)
    dem = subprocess.run(["c++filt"], input=names, capture_output=True, text=True).stdout.splitlines()
    for r, d in zip(rows, dem):
        r["demangled"] = d
    return rows


def main():
    rows = [r for r in parse() if len(sys.argv) < 2 or sys.argv[1] in r["demangled"]]
    print("# ptxas -v resources per kernel (sm_100a, flags of mole_b200/csrc/build.sh); %d kernels" % len(rows))
    print("%-5s %-6s %-8s %-8s %-7s %s" % ("regs", "stack", "spill_st", "spill_ld", "smem", "kernel"))
    for r in sorted(rows, key=lambda r: (-r["spill_st"], r["demangled"])):
        print("%-5d %-6d %-8d %-8d %-7d %s" % (r["regs"], r["stack"], r["spill_st"], r["spill_ld"], r["smem"], r["demangled"]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
