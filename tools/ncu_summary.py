"""Summarise an .ncu-rep (run `ncu -i X.ncu-rep --page raw --csv > X.csv` first) into the metrics we track."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
for d in data:
    print('=====')
    for w in want:
        if w in idx:
            print(f'  {w}: {d[idx[w]]} {units[idx[w]]}')
    st = sorted(((float(d[idx[h]].replace(",", "")) if d[idx[h]] not in ("", "n/a") else 0.0, h) for h in stall), reverse=True)[:6]
    for v, h in st:
        print(f'  stall {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")}: {v:.2f}')
