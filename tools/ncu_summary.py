"""Markdown table of the metrics we judge kernels by, from an `ncu --set full` report.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--title "..."] > profiles/xxx.md

Needs `ncu` on PATH (reads the report with `ncu -i ... --page raw --csv`)."""
import csv
import io
import subprocess
import sys

METRICS = [
    ('launch__grid_size', 'grid'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'L1/TEX throughput %'),
    ('lts__t_sectors_srcunit_tex_op_read.sum', 'L2->L1 read sectors'),
    ('l1tex__t_sector_hit_rate.pct', 'L1 sector hit rate %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'DMMA sub-pipe % (active)'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('sm__cycles_elapsed.avg', 'SM cycles elapsed'),
    ('sm__cycles_active.avg', 'SM cycles active (avg)'),
    ('sm__cycles_active.min', 'SM cycles active (min)'),
    ('sm__cycles_active.max', 'SM cycles active (max)'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait / issue'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math_pipe_throttle / issue'),
]


def main():
    rep = sys.argv[1]
    title = sys.argv[3] if len(sys.argv) > 3 and sys.argv[2] == '--title' else rep
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    names = [r[idx['Kernel Name']] for r in data]
    print('## %s\n' % title)
    print('| metric | ' + ' | '.join('launch %d' % (i + 1) for i in range(len(data))) + ' |')
    print('|---|' + '---|' * len(data))
    print('| kernel | ' + ' | '.join('`%s`' % n.split('(')[0].replace('void ', '')[:60] for n in names) + ' |')
    for key, label in METRICS:
        if key not in idx:
            continue
        u = units[idx[key]]
        vals = []
        for r in data:
            v = r[idx[key]]
            try:
                f = float(v)
                v = ('%d' % f) if f == int(f) and abs(f) < 1e15 else ('%.4g' % f)
            except ValueError:
                pass
            vals.append(v)
        print('| %s%s | ' % (label, (' [%s]' % u) if u else '') + ' | '.join(vals) + ' |')
    print()


if __name__ == '__main__':
    main()
