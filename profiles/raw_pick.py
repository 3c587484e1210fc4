"""Print the metrics that matter from an `ncu --page raw --csv` dump."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    print('----', r[idx['Kernel Name']][:60])
    for w in want:
        if w in idx:
            print(f"  {w:72s} {r[idx[w]]:>18s} {units[idx[w]]}")
    st = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
    vals = sorted(((float(r[idx[h]]) if r[idx[h]] not in ('', 'n/a') else 0, h) for h in st), reverse=True)[:7]
    for v, h in vals:
        print(f"  stall {v:7.2f}  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")
