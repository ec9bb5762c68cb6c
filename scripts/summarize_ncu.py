"""Summarise ncu outputs into small text files for profiles/.

  launch list : ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv <cmd>
                python scripts/summarize_ncu.py launches launches.csv > profiles/<round>_launches.txt
  full capture: ncu -i X.ncu-rep --page raw --csv > raw.csv
                python scripts/summarize_ncu.py raw raw.csv > profiles/<round>_<kernel>_raw.txt
"""
import collections
import csv
import sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_atom.sum',
        'smsp__average_warp_latency_per_inst_issued.ratio']


def launches(path):
    rows = list(csv.reader(open(path, errors='replace')))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d['Metric Name'] != 'gpu__time_duration.sum':
            continue
        v = float(d['Metric Value'].replace(',', ''))
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(d['Metric Unit'], 1e-3)
        a = agg.setdefault(d['Kernel Name'], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print('# kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)')
    print('# %-86s %6s %10s %7s' % ('kernel', 'count', 'avg_us', 'share'))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-88s %6d %10.2f %7.3f' % (k[:88], a[0], a[1] / a[0], a[1] / tot))
    print('# total %.1f us over %d launches' % (tot, sum(a[0] for a in agg.values())))


def raw(path):
    rows = list(csv.reader(open(path, errors='replace')))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('## %s   grid %s block %s' % (d['Kernel Name'], d.get('Grid Size', '?'), d.get('Block Size', '?')))
        for k in hdr:
            if k in KEEP or 'issue_stalled' in k and k.endswith('per_issue_active.ratio'):
                print('  %-92s %16s %s' % (k, d[k], units[hdr.index(k)]))


if __name__ == '__main__':
    {'launches': launches, 'raw': raw}[sys.argv[1]](sys.argv[2])
