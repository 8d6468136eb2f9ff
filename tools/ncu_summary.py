"""Print the metrics we track from an `ncu --page raw --csv` export. usage: python tools/ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.per_cycle_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get('Kernel Name', ''), d.get('launch__grid_size', ''))
    for h, u in zip(hdr, units):
        if h in want or ('issue_stalled' in h and 'per_issue_active' in h and float(d[h] or 0) > 0.05):
            print(f"  {h} [{u}] = {d[h]}")
