"""Buckets the SASS of an `ncu --page source --csv --print-source cuda,sass` export by address range:
warp instructions per tile (pass the tile count) and stall samples per bucket.
usage: python ncu_regions.py export.csv n_tiles [bucket_bytes]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n_tiles = float(sys.argv[2]); bs = int(sys.argv[3], 0) if len(sys.argv) > 3 else 0x800
h = None; cur = None; fileN = None
def f(x):
    try: return float(x)
    except Exception: return 0.0
seen = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fileN = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': h = r; continue
    if h is None or len(r) != len(h): continue
    if r[2] == '-':
        cur = (fileN, int(r[0])); continue
    try: a = int(r[2], 16)
    except ValueError: continue
    d = dict(zip(h, r))
    seen.setdefault(a, (cur, r[3].strip(), f(d['# Samples']), f(d['Instructions Executed'])))
addrs = sorted(seen)
print(len(addrs), 'sass instrs; warp instr per tile', sum(seen[a][3] for a in addrs) / n_tiles, '; samples', sum(seen[a][2] for a in addrs))
b0 = addrs[0]; buckets = {}
for a in addrs:
    k = (a - b0) // bs
    cur, txt, s, i = seen[a]
    bu = buckets.setdefault(k, [0, 0, set()])
    bu[0] += i; bu[1] += s
    if cur[0].endswith('.cu'): bu[2].add(cur[1])
for k in sorted(buckets):
    i, s, ls = buckets[k]; ls = sorted(ls)
    print(hex(b0 + k * bs), f'inst/tile={i / n_tiles:8.1f} samples={int(s):6d} lines {ls[0] if ls else 0}-{ls[-1] if ls else 0}')
