#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches_X.csv profiles/X_launches.md
  python scripts/summarize_ncu.py full gpurun_out/prof_X.ncu-rep profiles/X_full.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

FULL_METRICS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__cycles_active.avg', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_bytes.sum',
    'smsp__inst_executed.sum',
]


def short(name):
    name = re.sub(r'\(.*', '', name)
    name = name.replace('void ', '').replace('<unnamed>::', '').replace('unnamed>::', '')
    return name.strip()


def launches(src, dst):
    rows = []
    with open(src) as f:
        lines = [l for l in f if not l.startswith('==')]
    rd = csv.DictReader(io.StringIO(''.join(lines)))
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1e-3)
        rows.append((short(r['Kernel Name']), v * scale))
    agg = collections.OrderedDict()
    for k, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(dst, 'w') as f:
        f.write('# ncu launch list summary (%s)\n\n' % src)
        f.write('per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes\n\n')
        f.write('| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n')
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('| %s | %d | %.1f | %.1f%% | %.2f |\n' % (k, n, us, 100 * us / tot, us / n))
        f.write('\ntotal %d launches, %.1f us\n' % (len(rows), tot))
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as f:
        f.write('# ncu --set full summary (%s)\n\n' % src)
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write('## %s  (ID %s)\n\n' % (short(d['Kernel Name']), d.get('ID', '?')))
            for m in FULL_METRICS:
                if m in d:
                    f.write('- %s = %s %s\n' % (m, d[m], units[hdr.index(m)]))
            f.write('\n')
    print(open(dst).read())


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
