#!/usr/bin/env python3
"""Executed warp instructions of the sub-warp chain kernel per phase of pk_coop.cuh (limits, FK,
scan, error + Jlog, columns, box, QP loop, rounds ...), from scripts/ncu_lines.py output.
    python scripts/ncu_phases.py <lib.so> <mangled-kernel-substring> <report.ncu-rep> [warps]"""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
so, pat, rep = sys.argv[1:4]
out = subprocess.run([sys.executable, os.path.join(HERE, 'ncu_lines.py'), so, pat, rep, '2000'], capture_output=True, text=True).stdout
src = open(os.path.join(HERE, '..', 'pink_b200', 'csrc', 'pk_coop.cuh')).read().split('\n')
def find(s):
    for i,l in enumerate(src):
        if s in l: return i+1
    raise KeyError(s)
marks = [(find('limit check (pink/configuration.py'), 'limits'), (find('forward kinematics of the lane'), 'fk'), (find('segment factor of the suffix product'), 'segfactor'),
         (find('suffix product over the lanes: S_h'), 'scan'), (find('task error, Jlog6 (replicated)'), 'error+jlog'), (find('rows of -W Jlog6'), 'columns'),
         (find('diagonal part: posture rows'), 'diag'), (find('box rows, column norms'), 'box+norms'), (find('static PK_HD int solve_qp'), 'qp-init'),
         (find('for (; !to_rounds; ++it)'), 'qp-loop'), (find('static PK_HD int rounds'), 'rounds-gather'), (find('PK_HD void ik_step_coop'), 'epilogue')]
agg = collections.OrderedDict((n, [0, 0.0, 0]) for _, n in [(0,'pre')]+marks)
other = collections.Counter()
tot = 0
for l in out.split('\n'):
    m = re.match(r'(\S+)\s+(\d+)\s+(\d+)\s+([\d.]+)%\s+([\d.]+)\s+(\d+)\s+(\d+)', l)
    if not m: continue
    f, ln, n, thr, smp = m.group(1), int(m.group(2)), int(m.group(3)), float(m.group(5)), int(m.group(6))
    tot += n
    if f == 'pk_coop.cuh':
        name = 'pre'
        for a, nm in marks:
            if ln >= a: name = nm
        agg[name][0] += n; agg[name][1] += n * thr; agg[name][2] += smp
    else:
        other[f] += n
warps = float(sys.argv[4]) if len(sys.argv) > 4 else 2048
print("total", tot, "per warp", tot / warps)
for k, (n, t, s) in agg.items():
    print(f"{k:16s} {n:9d} {n / warps:8.1f}/warp  thr {t / max(n, 1):5.1f}  samples {s}")
for k, n in other.items(): print(f"{k:16s} {n:9d} {n / warps:8.1f}/warp")
