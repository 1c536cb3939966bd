"""Per-launch table of one EfficientNet forward from an ncu launch list (gpu__time_duration + dram bytes):
python tools/effnet_launch_table.py gpurun_out/launches_effnet_X.csv [--all]"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]; idx = {n: i for i, n in enumerate(h)}
per = {}
for r in rows[1:]:
    k = int(r[idx['ID']])
    m = r[idx['Metric Name']]; key = 'rd' if 'read' in m else ('wr' if 'write' in m else 't')
    per.setdefault(k, {})[key] = float(r[idx['Metric Value']])
    per[k]['grid'] = r[idx['Grid Size']]; per[k]['blk'] = r[idx['Block Size']]
    per[k]['name'] = r[idx['Kernel Name']].replace('void avexk::<unnamed>::', '').replace('avexk::<unnamed>::', '')[:26]
ks = sorted(per)
st = [k for k in ks if 'melspec_k' in per[k]['name']]
a = st[int(sys.argv[sys.argv.index("--fwd") + 1])] if "--fwd" in sys.argv else st[-1]
b = min([k for k in st if k > a] + [ks[-1] + 1])
tot = 0; byk = {}
for k in ks:
    if a - 1 <= k < b - 1:
        v = per[k]; t = v['t'] / 1e3; tot += t
        byk[v['name'][:12]] = byk.get(v['name'][:12], 0) + t
        if '--all' in sys.argv or t > 60:
            print(k, f"{v['name']:26s}", v['grid'], v['blk'], f"{t:8.1f}us rd {v['rd']/1e6:7.1f} wr {v['wr']/1e6:7.1f} MB  {(v['rd']+v['wr'])/v['t']:6.0f} GB/s")
print(round(tot, 1), {k: round(v, 1) for k, v in byk.items()})
