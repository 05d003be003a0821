#!/usr/bin/env python3
"""Attribute the per-SASS-instruction counters of an ncu report to CUDA source lines.

    python scripts/ncu_lines.py <lib.so> <mangled-kernel-substring> <report.ncu-rep> [top]

Uses nvdisasm's line info of the cubin embedded in the library and zips it, in
address order, with `ncu --page source --csv` of the first matching kernel."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    so, pat, rep = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and pat in l and l.rstrip().endswith(":"))
    # annotations come innermost first ("... inlined at ..."), then the enclosing call
    # sites; keep, per instruction, the innermost location that is not a CUDA header
    # or pk_math.cuh (so math helpers are charged to the phase that called them)
    inst, cur, group, fresh = [], None, [], True
    skip = ("pk_math.cuh",)
    for l in dis[start + 1:]:
        if l.startswith("//-----") or l.startswith(".text."):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                group, fresh = [], False
            group.append((os.path.basename(m.group(1)), int(m.group(2)), m.group(1)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            if not fresh:
                own = [g for g in group if "/csrc/" in g[2] and g[0] not in skip]
                pick = own[0] if own else (group[0] if group else None)
                cur = (pick[0], pick[1]) if pick else None
                fresh = True
            inst.append((int(m.group(1), 16), m.group(2).strip(), cur))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.split("\n")))
    hdr = next(r for r in rows if len(r) > 6 and r[0] == "Address")
    sass, k, take = [], 0, False
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name":
            if take:
                break
            take = pat_demangled(pat, r[1])
            continue
        if take and len(r) > 6 and r[0].startswith("0x"):
            sass.append(r)
    iex, ith, ismp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    print(f"# {len(inst)} SASS instructions in cubin, {len(sass)} in report")
    n = min(len(inst), len(sass))
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    ops = collections.defaultdict(int)
    for (off, txt, cur), r in zip(inst[:n], sass[:n]):
        a = agg[cur or ("?", 0)]
        a[0] += int(r[iex]); a[1] += int(r[ith]); a[2] += int(r[ismp]); a[3] += 1
        ops[txt.split()[0].split(".")[0]] += int(r[iex])
    tot = sum(a[0] for a in agg.values())
    thr = sum(a[1] for a in agg.values())
    print(f"# executed warp instructions: {tot}, avg active threads {thr / max(tot, 1):.1f}")
    byfile = collections.defaultdict(lambda: [0, 0, 0])
    for (f, l), a in agg.items():
        for i in range(3):
            byfile[f][i] += a[i]
    for f, a in sorted(byfile.items(), key=lambda x: -x[1][0]):
        print(f"{f:24s} {a[0]:10d} {100 * a[0] / tot:5.1f}%  threads {a[1] / max(a[0], 1):4.1f}  samples {a[2]}")
    print("# top lines: file line  warp-inst  share  avg-threads  stall-samples  static-sass")
    for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print(f"{f:20s} {l:5d} {a[0]:10d} {100 * a[0] / tot:5.1f}%  {a[1] / max(a[0], 1):4.1f}  {a[2]:5d}  {a[3]:5d}")
    print("# opcode mix")
    for o, c in sorted(ops.items(), key=lambda x: -x[1])[:25]:
        print(f"{o:10s} {c:10d} {100 * c / tot:5.1f}%")


def pat_demangled(pat, name):
    # mangled pattern like ik_chain_kernelILi6ELi1E -> demangled "ik_chain_kernel<(int)6, (int)1>"
    m = re.search(r"([A-Za-z_0-9]+?)I((?:Li\d+E)+)", pat)
    if not m:
        return pat in name
    args = re.findall(r"Li(\d+)E", m.group(2))
    return m.group(1).lstrip("0123456789") in name and ", ".join(f"(int){a}" for a in args) in name


if __name__ == "__main__":
    main()
