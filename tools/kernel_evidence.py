#!/usr/bin/env python3
"""Kernel evidence, regenerated from the built library:
  profiles/<round>_sass_histograms.jsonl : per-kernel SASS instruction histogram (cuobjdump -sass), IMAD.WIDE split by form,
                                            with the tcgen05 / TMA / LDGSTS mnemonics called out
  profiles/<round>_ptxas_table.tsv       : registers / spill bytes / stack / shared memory per kernel (ptxas -v log of the build)
usage: python tools/kernel_evidence.py [round-tag, default r02]"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sass_mix

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "halo2_scaffold_b200", "lib", "libh2b200.so")
log = os.path.join(ROOT, "halo2_scaffold_b200", "lib", "ptxas.log")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


rows = list(sass_mix.kernels(lib))
dm = demangle([n for n, _ in rows])
with open(os.path.join(ROOT, "profiles", tag + "_sass_histograms.jsonl"), "w") as f:
    for name, ins in rows:
        c = sass_mix.mix(ins)
        special = {k: v for k, v in c.items() if any(s in k for s in ("UTMA", "UTCMMA", "TCGEN", "LDGSTS", "UBLKCP", "SYNCS", "REDG", "ATOMG", "RED.", "ATOM"))}
        f.write(json.dumps({"kernel": re.sub(r"\(.*", "", dm[name]), "instructions": len(ins), "mix": dict(c.most_common(16)), "memory_and_async": special}) + "\n")

entries, cur = [], None
for line in open(log):
    m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", line)
    if m:
        cur = {"kernel": m.group(1)}
        entries.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        cur["stack"], cur["spill_st"], cur["spill_ld"] = m.groups()
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = m.group(1)
        s = re.search(r"(\d+) bytes smem", line)
        cur["smem"] = s.group(1) if s else "0"
dm = demangle([e["kernel"] for e in entries])
with open(os.path.join(ROOT, "profiles", tag + "_ptxas_table.tsv"), "w") as f:
    f.write("kernel\tregisters\tstack_bytes\tspill_store_bytes\tspill_load_bytes\tstatic_smem_bytes\n")
    for e in entries:
        f.write("%s\t%s\t%s\t%s\t%s\t%s\n" % (re.sub(r"\(.*", "", dm[e["kernel"]]), e.get("regs", "?"), e.get("stack", "0"), e.get("spill_st", "0"), e.get("spill_ld", "0"), e.get("smem", "0")))
print("wrote %d kernels (SASS), %d entries (ptxas)" % (len(rows), len(entries)))
