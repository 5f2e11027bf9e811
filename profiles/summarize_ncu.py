#!/usr/bin/env python
"""Summarise ncu reports / launch lists into the small text files committed under profiles/.
usage: python profiles/summarize_ncu.py raw <report.ncu-rep> | launches <launches.csv>"""
import collections, csv, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('---', r[hdr.index('Kernel Name')], 'grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'  {w} = {r[i]} {units[i]}')
        br, bw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
        def tobytes(v, u):
            v = float(v.replace(',', ''))
            return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
        t = tobytes(r[br], units[br]) + tobytes(r[bw], units[bw])
        d = float(r[hdr.index('gpu__time_duration.sum')].replace(',', ''))
        du = units[hdr.index('gpu__time_duration.sum')]
        d *= {'ns': 1e-9, 'us': 1e-6, 'usecond': 1e-6, 'ms': 1e-3, 'msecond': 1e-3, 'nsecond': 1e-9, 'second': 1}[du]
        print(f'  traffic(dram read+write) = {t/1e6:.2f} MB ; dram GB/s under ncu = {t/d/1e9:.0f}')


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    print(f'{len(rows)-1} launches, total {tot/1e3:.1f} us (cold-cache, serialised under ncu: compare SHARES)')
    for k, v in agg.items():
        print(f'  {k[:70]:70s} n={len(v):3d} avg={sum(v)/len(v)/1e3:8.2f} us  share={sum(v)/tot:.3f}')


if __name__ == '__main__':
    {'raw': raw, 'launches': launches}[sys.argv[1]](sys.argv[2])
