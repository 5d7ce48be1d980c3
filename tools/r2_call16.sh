for K in 32 16; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_small_launches_k$K.csv python tools/small_fit_one.py $K tensor 12 > gpurun_out/r2_small_ncu_k$K.log 2>&1; echo "rc=$?"
python - <<PY
import csv, io, collections
rows=[r for r in csv.reader(io.StringIO("".join(l for l in open('gpurun_out/r2_small_launches_k$K.csv') if l.startswith('"'))))]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
agg=collections.OrderedDict()
for r in rows[1:]:
    name=r[ix["Kernel Name"]][:60]+" grid "+r[ix["Grid Size"]]; t=float(r[ix["Metric Value"]].replace(",",""))
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=t
print("K=$K")
for name,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:16]:
    print(f"{name:95s} {n:5d} launches  avg {t/n/1e3:9.1f} us  total {t/1e3:10.1f} us")
PY
done
