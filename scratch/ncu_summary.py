import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg','l1tex__data_pipe_lsu_wavefronts_mem_shared.avg','sm__cycles_active.avg','launch__grid_size','launch__occupancy_limit_registers','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']
def col(name):
    for i,h in enumerate(hdr):
        if h==name or h.endswith('.'+name): return i
    return None
idx=[col(w) for w in want]
ki=hdr.index('Kernel Name')
for r in rows[2:]:
    print('---', r[ki][:70])
    for w,i in zip(want,idx):
        if i is not None: print(f'  {w}: {r[i]} {units[i]}')
