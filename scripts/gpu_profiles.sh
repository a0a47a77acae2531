# round profile captures: launch list of the default bench command + one full capture per dominant kernel.
# Reports are summarised on the box (gpurun brings back at most 64 MiB) and removed.
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b.log 2>&1
for cfg in "c2 k_cell_scalar" "c5 k_cell_syrk" "c3 k_cell_syrk" "c4 k_cell_blocks"; do set -- $cfg
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o /tmp/prof_$1_final python bench.py --config $1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_run.log 2>&1
  tail -1 gpurun_out/ncu_run.log
  (python profiles/ncu_summary.py /tmp/prof_$1_final.ncu-rep; echo; echo "regions between barriers:"; python profiles/sass_regions.py /tmp/prof_$1_final.ncu-rep) > gpurun_out/r01_prof_$1_final.txt 2>&1
  rm -f /tmp/prof_$1_final.ncu-rep
done
