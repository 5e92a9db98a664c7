run() {  # n, tag, extra args
  n=$1; tag=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --no-cpu-baseline "$@" > gpurun_out/r2c_bench_${tag}_${n}gpu.json 2> gpurun_out/r2c_bench_${tag}_${n}gpu.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2c_bench_${tag}_${n}gpu.json") if l.startswith("{")][-1])
print("$tag N=$n value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"nccl",(d.get("with_nccl_gather") or {}).get("value"),"k_us",d["roofline"]["kernel_us_per_launch"],"loop",{k:v for k,v in (d["e2e"].get("sampler_loop") or {}).items() if k!="note"})
PY
}
python -m pytest tests -m gpu -q -k "multi" > gpurun_out/r2c_pytest_gpu_multi_8gpu.log 2>&1; tail -2 gpurun_out/r2c_pytest_gpu_multi_8gpu.log
run 8 c1 --steps 150 --warmup 5
run 4 c1 --steps 150 --warmup 5
run 8 c1_65536 --walkers 65536 --steps 20 --warmup 3
