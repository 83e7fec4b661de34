import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
keys=['gpu__time_duration.sum','sm__cycles_elapsed.avg','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fp64.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_fmaheavy.sum','sm__inst_executed_pipe_fmalite.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    name=r[hdr.index('Kernel Name')]
    print('==',name[:80])
    for i,h in enumerate(hdr):
        if h in keys or h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') or h.startswith('smsp__average_warp') and 'per_issue_active' in h:
            print('  %-90s %-12s %s'%(h,units[i],r[i]))
