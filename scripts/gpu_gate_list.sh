# launch list of the gate-stage kernels only (device leg of bench.py, 12 scans)
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"forest_|pat_table|grid_build|live_scan|tree_off" -c 108 --csv --log-file gpurun_out/gate_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/gl.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/gate_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]),r[ki].split('(')[0]),{})[r[mi]]=float(r[vi].replace(',',''))
ids=sorted(d)
for k in ids[-9:]:
    v=d[k]; print(k, {m.split('__')[1][:14]:round(x,1) for m,x in v.items()})
PY
