mkdir -p gpurun_out
python scripts/part_local_time.py 2 1 2>&1 | tail -1
python scripts/part_local_time.py 8 0 2>&1 | tail -1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_march_hex -s 3 -c 1 -f -o /tmp/prof_p python scripts/part_local_time.py 2 1 > gpurun_out/ncu_run.log 2>&1
(python profiles/ncu_summary.py /tmp/prof_p.ncu-rep 25; python profiles/sass_hist.py /tmp/prof_p.ncu-rep) > gpurun_out/r02_prof_c2_march_cellmap.txt 2>&1
