"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line:
samples, instructions and the top stall reasons.  usage: python ncu_lines.py export.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
h = None
lines = {}
def f(x):
    try: return float(x)
    except Exception: return 0.0
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': h = r; continue
    if h is None or len(r) != len(h) or r[2] != '-': continue   # only the per-line summary rows
    d = dict(zip(h, r))
    # duplicated 'Source' header: first = cuda text
    key = (cur_file, int(r[0]))
    lines[key] = (r[1], d)
stall = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(f(d['# Samples']) for _, d in lines.values())
toti = sum(f(d['Instructions Executed']) for _, d in lines.values())
print('total samples', int(tot), 'warp instructions', int(toti))
for key, (src, d) in sorted(lines.items(), key=lambda kv: -f(kv[1][1]['# Samples']))[:top_n]:
    st = sorted(((f(d[c]), c) for c in stall), reverse=True)[:3]
    print(f'{key[0][:14]:14s}{key[1]:5d} {int(f(d["# Samples"])):7d} {f(d["# Samples"]) / tot:6.3f} inst={int(f(d["Instructions Executed"])):10d} ',
          ' '.join(f'{c[6:]}={int(v)}' for v, c in st), '|', src.strip()[:80])
print({c[6:]: int(sum(f(d[c]) for _, d in lines.values())) for c in stall})
