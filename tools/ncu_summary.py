import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__cycles_active.avg.pct_of_peak_sustained_elapsed','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors.sum.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:70])
    for k in keys:
        if k in hdr:
            i=hdr.index(k); print(f'  {k} = {r[i]} {units[i]}')
    st={k.replace('smsp__pcsamp_warps_issue_stalled_',''):int(r[i]) for i,k in enumerate(hdr) if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('_not_issued')}
    tot=sum(st.values()) or 1
    print('  stalls:', ', '.join(f'{k}={100*v/tot:.0f}%' for k,v in sorted(st.items(), key=lambda x:-x[1])[:7]))
