#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` export: instructions executed,
stall samples, lane use.  SASS rows are summed under the source line they follow.
usage: ncu_lines.py source.csv [top] [--remap wrongname=path]   (ncu sometimes labels a header with the wrong file name)"""
import csv, sys, os
args = [a for a in sys.argv[1:] if not a.startswith('--remap=')]
remap = dict(a[8:].split('=') for a in sys.argv[1:] if a.startswith('--remap='))
rows = list(csv.reader(open(args[0])))
top = int(args[1]) if len(args) > 1 else 60
agg = {}
f = None
cur = None
def num(x):
    try: return float(x)
    except: return 0.0
for r in rows:
    if not r: continue
    if r[0] == 'File Path':
        f = r[1]; continue
    if r[0] in ('Function Name', 'Line No'): continue
    if r[0].isdigit():
        cur = (f, int(r[0])); agg.setdefault(cur, [0.0, 0.0, 0.0]); continue
    if r[0] == '' and len(r) > 8 and r[2].startswith('0x') and cur:
        a = agg[cur]; a[0] += num(r[4]); a[1] += num(r[7]); a[2] += num(r[8])
srccache = {}
def text(fn, ln):
    fn = remap.get(os.path.basename(fn), fn)
    if fn not in srccache:
        try: srccache[fn] = open(fn).read().split('\n')
        except Exception: srccache[fn] = []
    L = srccache[fn]
    return (os.path.basename(fn), L[ln-1].strip()[:105] if 0 < ln <= len(L) else '?')
ti = sum(a[1] for a in agg.values()); ts = sum(a[0] for a in agg.values()); tt = sum(a[2] for a in agg.values())
print("total warp-instr %.0f  samples %.0f  thread-instr %.0f  avg lanes %.1f" % (ti, ts, tt, tt/max(ti,1)))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    fn, tx = text(*k)
    print("%5.1f%% inst %5.1f%% smp lanes %4.1f  %s:%d  %s" % (100*a[1]/ti, 100*a[0]/ts, a[2]/max(a[1],1), fn, k[1], tx))
