#!/usr/bin/env python3
"""Development aid: one-screen summary of an .ncu-rep (per kernel) and, with --lines KERNEL, its hottest source lines."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size',
        'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld_lookup_miss.sum']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    h = r[0]
    print('%-55s' % 'metric', *['%-22s' % row[h.index('Kernel Name')][:22] for row in r[2:]])
    for w in WANT:
        if w in h:
            i = h.index(w)
            print('%-55s' % (w[:48] + ' ' + r[1][i][:6]), *['%-22s' % row[i][:20] for row in r[2:]])


def lines(rep, kernel, top=40, by=1):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + kernel],
                         capture_output=True, text=True).stdout
    cur, agg, ti, ts, first = None, {}, 0, 0, None
    for x in csv.reader(out.splitlines()):
        if len(x) == 2 and x[0] == 'Function Name':
            if first is None:
                first = x[1]
            elif x[1] != first:
                break
        if len(x) == 2 and x[0] == 'File Path':
            cur = x[1].split('/')[-1]
            continue
        if len(x) < 10 or x[0] in ('Line No', ''):
            continue
        try:
            ln, s, i, t = int(x[0]), int(x[6]), int(x[7]), int(x[8])
        except ValueError:
            continue
        a = agg.setdefault((cur, ln, x[1].strip()[:105]), [0, 0, 0])
        a[0] += s; a[1] += i; a[2] += t
        ti += i; ts += s
    print(first, 'inst', ti, 'samples', ts)
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][by])[:top]:
        print(f'{k[0]}:{k[1]:4d} inst {100 * a[1] / ti:5.1f}% samp {100 * a[0] / max(ts, 1):5.1f}% thr {a[2] / max(a[1], 1):5.1f} | {k[2]}')


def md(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    h = r[0]
    names = [row[h.index('Kernel Name')].split('(')[0].replace('void ', '') for row in r[2:]]
    print('| metric | ' + ' | '.join(names) + ' |')
    print('|---|' + '---|' * len(names))
    for w in WANT:
        if w in h:
            i = h.index(w)
            vals = []
            for row in r[2:]:
                try:
                    vals.append('%.4g' % float(row[i]))
                except ValueError:
                    vals.append(row[i])
            print(f'| {w} ({r[1][i]}) | ' + ' | '.join(vals) + ' |')


if __name__ == '__main__':
    if '--md' in sys.argv:
        md(sys.argv[1])
        sys.exit(0)
    if '--lines' in sys.argv:
        k = sys.argv[sys.argv.index('--lines') + 1]
        by = 0 if '--by-samples' in sys.argv else 1
        lines(sys.argv[1], k, by=by)
    else:
        raw(sys.argv[1])
