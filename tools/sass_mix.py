#!/usr/bin/env python3
"""Instruction mix of every kernel in a cubin / executable / shared library (cuobjdump -sass), one JSON line per kernel.
IMAD.WIDE is split by its addend: RZ (pure product, full rate) vs a register pair (accumulating form) vs .X (carry-in).
usage: python tools/sass_mix.py <binary> [substring-filter]"""
import collections
import json
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    name, rows = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name:
                yield name, rows
            name, rows = m.group(1), []
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
        if m and name:
            rows.append(m.group(1).strip())
    if name:
        yield name, rows


def mix(rows):
    c = collections.Counter()
    for r in rows:
        r = re.sub(r"^@!?U?P\d+\s+", "", r)
        op = r.split()[0]
        if op.startswith("IMAD.WIDE"):
            if ".X" in op:
                op = "IMAD.WIDE.X"
            elif r.rstrip().endswith("RZ"):
                op = "IMAD.WIDE(RZ)"
            else:
                op = "IMAD.WIDE(acc)"
        c[op] += 1
    return c


if __name__ == "__main__":
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    for name, rows in kernels(sys.argv[1]):
        if flt not in name:
            continue
        c = mix(rows)
        print(json.dumps({"kernel": name, "instructions": len(rows), "mix": dict(c.most_common(14))}))
