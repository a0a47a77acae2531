# round-2 profile captures of the default bench command: launch list (durations + DRAM bytes) and one full capture of k_march_hex
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_march_hex -s 4 -c 1 -f -o /tmp/prof_c2 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_run.log 2>&1
tail -1 gpurun_out/ncu_run.log
(python profiles/ncu_summary.py /tmp/prof_c2.ncu-rep 30; python profiles/sass_hist.py /tmp/prof_c2.ncu-rep) > gpurun_out/r02_prof_c2_final.txt 2>&1
