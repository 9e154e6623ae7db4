cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python profiles/experiments/r2_elem_select_launches.py > gpurun_out/r2_exp35_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_exp35_launches.csv python profiles/experiments/r2_elem_select_launches.py > gpurun_out/r2_exp35_ncu.log 2>&1
echo rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_exp35_launches.csv') if l.startswith('"'))]
h=rows[0]; k=h.index("Kernel Name"); m=h.index("Metric Name"); v=h.index("Metric Value"); i=h.index("ID")
cur={}
for r in rows[1:]:
    cur.setdefault(r[i], [r[k][:70], {}])[1][r[m]]=r[v]
for id_,(name,d) in list(cur.items())[-24:]:
    print(id_, name, d)
PY
