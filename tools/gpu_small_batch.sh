for b in 100 256; do
  python tools/time_host_call.py $b 2>&1 | tail -5
  ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_b$b.csv python tools/time_host_call.py $b > /dev/null 2>&1
  python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_b$b.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:44]].append(float(r[vi].replace(",","")))
    except: pass
for k,v in agg.items(): print("  B=$b", k, len(v), round(sum(v)/len(v)/1e3,2), "us")
PY
done
