"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/xxx_launches_summary.md
"""
import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline='') as fh:
        lines = [l for l in fh if not l.startswith('==')]
    rd = csv.reader(lines)
    hdr = next(rd)
    col = {h: i for i, h in enumerate(hdr)}
    tot = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if len(r) < len(hdr) or r[col['Metric Name']] != 'gpu__time_duration.sum':
            continue
        name = r[col['Kernel Name']]
        unit = r[col['Metric Unit']]
        v = float(r[col['Metric Value']].replace(',', ''))
        ms = v / 1e6 if unit in ('ns', 'nsecond') else v / 1e3 if unit in ('us', 'usecond') else v
        t = tot[name]
        t[0] += 1
        t[1] += ms
    ours = {k: v for k, v in tot.items() if 'rl::' in k or k.startswith(('gram', 'spmm', 'gemm', 'update', 'ew_', 'dots', 'syevj',
                                                                          'split', 'minmax', 'fill', 'sell', 'pack', 'colsum'))}
    total = sum(v[1] for v in ours.values())
    print('| kernel | launches | total ms | share |')
    print('|---|---:|---:|---:|')
    for k, (c, ms) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.3f | %.1f%% |' % (k.split('(')[0][:90], c, ms, 100 * ms / total if total else 0))
    print('\nTotal device time of libraleigh_b200.so kernels: %.1f ms over %d launches; other kernels in the list '
          '(torch: synthetic data, verification): %d launches.' %
          (total, sum(v[0] for v in ours.values()), sum(v[0] for k, v in tot.items() if k not in ours)))


if __name__ == '__main__':
    main()
