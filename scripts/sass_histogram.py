"""Opcode histogram of one kernel from `cuobjdump -sass` (which SASS the hot kernel really consists of).

    python scripts/sass_histogram.py <object-or-library> <substring of the mangled kernel name> [--per-node LANES]
"""
import collections
import re
import subprocess
import sys


def histogram(path, needle):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    name, counts = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name and needle in name:
            counts.setdefault(name, collections.Counter())[m.group(1)] += 1
    return counts


if __name__ == "__main__":
    path, needle = sys.argv[1], sys.argv[2]
    for name, c in histogram(path, needle).items():
        total = sum(c.values())
        print(f"{name}: {total} instructions")
        print("  " + ", ".join(f"{op} {n}" for op, n in c.most_common(40)))
