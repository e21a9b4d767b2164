"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per tracked frame: python scripts/launch_summary.py file.csv"""
import csv, collections, re, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines); hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for row in r:
    v = float(row[vi].replace(',', ''))
    v = v / 1000 if row[ui] == 'ns' else v * 1000 if row[ui] == 'ms' else v
    name = row[ki]
    m = re.search(r'(\w+)\s*(<[^(]*)?\(', name.replace('<unnamed>::', ''))
    short = m.group(1) if m else name[:40]
    if 'DeviceRadixSort' in name: short = 'cub_radix_sort_' + ('onesweep' if 'Onesweep' in name else 'hist' if 'Histogram' in name else 'sum')
    if 'DeviceScan' in name: short = 'cub_scan'
    seq.append((short, v))
idx = [i for i, (n, _) in enumerate(seq) if n == 'backproject_kernel']
print("launches", len(seq), "frame starts", idx)
fr = seq[idx[-2]:idx[-1]]
agg = collections.OrderedDict()
for n, v in fr:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in fr)
print("launches per frame", len(fr), "sum of kernel times us", round(tot, 1))
own = [n for n in agg if n.endswith('_kernel') and not n.startswith('cub') and 'elementwise' not in n and 'indices' not in n]
print("own launches", sum(agg[n][0] for n in own), "library launches", len(fr) - sum(agg[n][0] for n in own))
print(f"| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| `{n}` | {c} | {v:.1f} | {v / c:.2f} | {100 * v / tot:.1f}% |")
print("jtj per pass:", [round(v, 1) for n, v in fr if n == 'data_jtj_kernel'])
