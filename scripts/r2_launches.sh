mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/b.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_c2.csv')) if len(r)>10]
hdr=rows[0]; 
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
from collections import OrderedDict
d=OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii], r[ki][:60]),{})[r[mi]]=r[vi]
for (i,k),m in list(d.items())[-16:]:
    print(i,k,m)
PY
