# C3 / C4 bench lines + launch list of C4 (per-kernel durations)
tag=$1
python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err
python bench.py --workload c3 --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_c4.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import json,csv,collections
for w in ("c4","c3"):
    d=json.loads(open("gpurun_out/${tag}_bench_%s.json"%w).read().strip().splitlines()[-1])
    print(w,"value",d["value"],"e2e",d["e2e"]["value"],"frac",d["roofline"]["frac"],"kernel_us",d["roofline"]["kernel_us_per_launch"])
rows=[r for r in csv.reader(open("gpurun_out/${tag}_launches_c4.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:40]].append(float(r[vi].replace(",","")))
    except: pass
for k,v in agg.items(): print(k, len(v), sum(v)/len(v))
PY
