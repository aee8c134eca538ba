"""ncu DRAM counters -> profiles/r02_dram_traffic.json (what bench.py's `traffic` / `frac_dram` fields read).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:'k_mp_edge_tc3|k_aggregate_dets|k_det_prepare' -s 60 -c 36 --csv --log-file gpurun_out/r02_dram.csv <bench cmd>
    python profiles/make_dram_traffic.py gpurun_out/r02_dram.csv profiles/r02_dram_traffic.json "<bench cmd>"

Only steady-state launches count (window full: duration within 5 % of the longest launch of the kernel).  Association rows
per launch are derived from the edge kernel's own write traffic: it writes 256 B of state + 8 B of logit / score per row
and nothing else."""
import collections
import csv
import json
import subprocess
import sys

src, out = sys.argv[1], sys.argv[2]
cmd = sys.argv[3] if len(sys.argv) > 3 else ''
rows = list(csv.DictReader(l for l in open(src) if l.startswith('"')))
launch = collections.OrderedDict()
for r in rows:
    key = (int(r['ID']), r['Kernel Name'].split('(')[0].replace('<unnamed>::', ''))
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    if r['Metric Name'] == 'gpu__time_duration.sum':
        v *= {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3}.get(u, 1e-6)
    else:
        v *= {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1.0)
    launch.setdefault(key, {})[r['Metric Name']] = v
per = collections.defaultdict(list)
for (i, name), m in launch.items():
    per[name].append(m)


def steady(ms):
    top = max(m['gpu__time_duration.sum'] for m in ms)
    return [m for m in ms if m['gpu__time_duration.sum'] >= 0.95 * top]


edge = steady(per['k_mp_edge_tc3'])
mean = lambda ms, k: sum(m[k] for m in ms) / len(ms)
rows_per_launch = mean(edge, 'dram__bytes_write.sum') / 264.0
git = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip()
res = {'source': f'{src}: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none of `{cmd}`',
       'git': git, 'rows_per_launch': round(rows_per_launch),
       'rows_per_launch_how': 'write bytes of k_mp_edge_tc3 / 264 (256 B state + 8 B logit / score per association row)'}
for name, ms in per.items():
    st = steady(ms)
    rd, wr, t = mean(st, 'dram__bytes_read.sum'), mean(st, 'dram__bytes_write.sum'), mean(st, 'gpu__time_duration.sum')
    res[name] = {'kernel': name, 'steady_state_launches': len(st), 'dram_read_bytes': round(rd), 'dram_write_bytes': round(wr),
                 'ms_under_ncu': round(t, 4), 'dram_bytes_per_edge_row': round((rd + wr) / rows_per_launch, 1),
                 'dram_gbs_under_ncu': round((rd + wr) / t / 1e6, 1)}
json.dump(res, open(out, 'w'), indent=1)
print(json.dumps(res, indent=1))
