python bench.py --profile micro > /dev/null 2>&1 && ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:sample_pdf -c 80 --csv --log-file gpurun_out/pdf_inst.csv python bench.py --profile micro > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/pdf_inst.csv")) if len(r)>10]
h=rows[0]; ik=h.index("Kernel Name"); im=h.index("Metric Name"); iv=h.index("Metric Value"); iid=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault(r[iid],{})[r[im]]=float(r[iv]); d[r[iid]]["k"]=r[ik][:60]
agg=collections.OrderedDict()
for k,v in d.items():
    key=(v["k"], int(v["smsp__inst_executed.sum"]))
    agg.setdefault(key, []).append(v["gpu__time_duration.sum"])
for (k,inst),ts in agg.items(): print(k, "inst/ray", round(inst/2**20,1), "launches", len(ts), "us", round(sum(ts)/len(ts)/1e3,1))
PY
