nproc
for n in 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-cpu-baseline --steps 150 --warmup 5 > gpurun_out/r2b_bench_c1_${n}gpu.json 2> gpurun_out/r2b_bench_c1_${n}gpu.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2b_bench_c1_${n}gpu.json") if l.startswith("{")][-1])
print("N=$n value",d["value"],"e2e",d["e2e"]["value"],"loop",{k:v for k,v in (d["e2e"].get("sampler_loop") or {}).items() if k!="note"})
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --no-cpu-baseline --walkers 65536 --steps 20 --warmup 3 > gpurun_out/r2b_bench_c1_65536_8gpu.json 2> gpurun_out/r2b_bench_c1_65536_8gpu.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2b_bench_c1_65536_8gpu.json") if l.startswith("{")][-1])
print("65536 N=8 value",d["value"],"e2e",d["e2e"]["value"],"loop",{k:v for k,v in (d["e2e"].get("sampler_loop") or {}).items() if k!="note"})
PY
