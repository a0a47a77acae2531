mkdir -p gpurun_out
M=smsp__inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,launch__grid_size,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,lts__t_sectors_op_red.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active
for r in 0 1; do
PL_KIND=heat PL_NEL=200 timeout 600 ncu --metrics $M -k regex:k_march_hex -s 55 -c 1 --csv --log-file gpurun_out/own_rank$r.csv python scripts/part_local_time.py 2 $r > gpurun_out/own_rank$r.log 2>&1
tail -1 gpurun_out/own_rank$r.log | cut -c1-300
done
