"""Summarise an `ncu --page raw --csv` export: key metrics + stall reasons per kernel.
    python tools/ncu_summary.py report.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
for r in rows[2:]:
    print('====', r[hdr.index('Kernel Name')][:70])
    for n in KEYS:
        if n in hdr:
            print(f"  {n:75s} {r[hdr.index(n)]}")
    st = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
          if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('_not_issued') and r[i] not in ('', 'n/a')]
    tot = sum(v for v, _ in st) or 1.0
    for v, h in sorted(st, reverse=True)[:9]:
        print(f"     {100 * v / tot:6.2f}%  {h.replace('smsp__pcsamp_warps_issue_stalled_', '')}")
