"""Print the key metrics of an .ncu-rep (raw page): usage python profiles/ncu_keys.py file.ncu-rep [kernel-substring]"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit', 'smsp__issue_active.avg.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'lts__t_sector_hit_rate',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        '_per_issue_active.ratio', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct', 'l1tex__t_sector_hit_rate', 'launch__grid_size',
        'lts__t_sectors_srcunit_tex_op_write', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum', 'sm__inst_executed_pipe_fp64', 'smsp__thread_inst_executed_per_inst_executed']
for r in rows[2:]:
    name = r[h.index('Kernel Name')]
    if len(sys.argv) > 2 and sys.argv[2] not in name:
        continue
    print('==', name)
    for i, k in enumerate(h):
        if any(x in k for x in keys) and 'pcsamp' not in k:
            try:
                if float(r[i]) == 0:
                    continue
            except ValueError:
                pass
            print('  %-90s %-10s %s' % (k, rows[1][i], r[i]))
