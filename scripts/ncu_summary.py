"""Headline metrics of every kernel in `ncu -i X.ncu-rep --page raw --csv` output read from stdin."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get('Kernel Name', '')[:100])
    for k in keys:
        if k in d:
            print('   ', k, d[k], rows[1][hdr.index(k)])
    for k in hdr:
        if 'issue_stalled' in k and 'ratio' in k and 'not_issued' not in k:
            try:
                v = float(d[k])
                if v > 0.25:
                    print('    stall', k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), round(v, 2))
            except ValueError:
                pass
