#!/usr/bin/env python3
"""Summarise ncu outputs into profiles/ (launch list CSV -> per-kernel table; .ncu-rep raw page
-> key metrics).  usage: ncu_summary.py launches <csv> <out.txt> <title> | kernel <ncu-rep> <out.txt> <title>
| metrics <ncu-rep> <out.csv> <title>   (metric,unit,value lines of the first kernel; bench.py parses dram__bytes_*)"""
import collections
import csv
import subprocess
import sys

KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.sum',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed',
        'sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_elapsed.max']


def launches(path, out, title):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row['Kernel Name'].split('(')[0]
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}[row['Metric Unit']]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    o = ['# ' + title, '%-70s %8s %12s %8s' % ('kernel', 'launches', 'total_ms', 'share')]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        o.append('%-70s %8d %12.3f %7.1f%%' % (k[:70], v[0], v[1] / 1e6, 100 * v[1] / tot))
    o.append('%-70s %8d %12.3f' % ('TOTAL', sum(v[0] for v in agg.values()), tot / 1e6))
    open(out, 'w').write('\n'.join(o) + '\n')
    print('\n'.join(o))


def kernel(rep, out, title):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.split('\n')))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    o = ['# ' + title]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        for k in KEYS:
            if k in idx:
                o.append('  %-78s %-16s %s' % (k, units[idx[k]], r[idx[k]][:120]))
        o.append('  warp stall reasons (warps per issue-active cycle):')
        for i, h in enumerate(hdr):
            if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio'):
                v = float(r[i])
                if v > 0.03:
                    o.append('    %-40s %.3f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
        o.append('')
    open(out, 'w').write('\n'.join(o) + '\n')
    print('\n'.join(o))


CSV_KEYS = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
            'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
            'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
            'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max']


def metrics(rep, out, title):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.split('\n')))
    hdr, units, r = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}
    o = ['# metric, unit, value: ' + title]
    for k in CSV_KEYS:
        if k in idx:
            o.append('%s,%s,%s' % (k, units[idx[k]], r[idx[k]].replace(',', '')))
    open(out, 'w').write('\n'.join(o) + '\n')
    print('\n'.join(o))


if __name__ == '__main__':
    {'launches': launches, 'kernel': kernel, 'metrics': metrics}[sys.argv[1]](*sys.argv[2:5])
