mkdir -p gpurun_out
for r in 0 1; do
PL_KIND=${PLK:-elasticity} PL_NEL=${PLN:-128} timeout 600 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,launch__grid_size -k regex:k_march -s ${PLS:-3} -c 1 --csv --log-file gpurun_out/pl_rank$r.csv python scripts/part_local_time.py 2 $r > gpurun_out/pl_rank$r.log 2>&1
tail -1 gpurun_out/pl_rank$r.log
done
