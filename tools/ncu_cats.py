#!/usr/bin/env python
"""Instruction share per source-line RANGE of an ncu source-page CSV.  usage: ncu_cats.py source.csv nwarps file:lo-hi=name ..."""
import csv, os, sys
rows = list(csv.reader(open(sys.argv[1])))
nwarps = float(sys.argv[2])
cats = []
for spec in sys.argv[3:]:
    loc, name = spec.split('=')
    f, rng = loc.split(':')
    lo, hi = rng.split('-')
    cats.append((f, int(lo), int(hi), name))
agg = {}; f = None; cur = None
def num(x):
    try: return float(x)
    except: return 0.0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': f = os.path.basename(r[1]); continue
    if r[0] in ('Function Name', 'Line No'): continue
    if r[0].isdigit(): cur = (f, int(r[0])); agg.setdefault(cur, [0, 0, 0]); continue
    if r[0] == '' and len(r) > 8 and r[2].startswith('0x') and cur:
        a = agg[cur]; a[0] += num(r[4]); a[1] += num(r[7]); a[2] += num(r[8])
tot = {}
for (f, l), a in agg.items():
    name = f
    for cf, lo, hi, nm in cats:
        if cf == f and lo <= l <= hi: name = nm; break
    t = tot.setdefault(name, [0, 0, 0])
    for k in range(3): t[k] += a[k]
ti = sum(t[1] for t in tot.values()); ts = sum(t[0] for t in tot.values())
print("total warp-instr %.0f = %.0f per warp" % (ti, ti / nwarps))
for n, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-28s %5.1f%% inst %5.1f%% smp  lanes %4.1f  %6.0f inst/warp" % (n, 100 * t[1] / ti, 100 * t[0] / max(ts, 1), t[2] / max(t[1], 1), t[1] / nwarps))
