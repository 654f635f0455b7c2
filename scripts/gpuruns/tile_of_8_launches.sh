mkdir -p gpurun_out
MA_PART=0,8 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r3h_launches_tile.csv python scripts/prof_eval.py c3 1.0 3 > gpurun_out/r3h_ncu.log 2>&1
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/r3h_launches_tile.csv') if not l.startswith('==')]
seq=[]
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
    seq.append((row['Kernel Name'][:60], v, row.get('Grid Size','')))
# last evaluation: from the last k_gather_w on
idx=[i for i,(n,_,_) in enumerate(seq) if n.startswith('k_gather_w')]
tot=0
for n,v,g in seq[idx[-1]:]:
    print(f"{n:60s} {v:8.1f} us  grid {g}"); tot+=v
print("sum", tot)
PY
