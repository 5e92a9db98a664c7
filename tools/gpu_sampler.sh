# native sampler loop: GPU tests, the example run with breakdown, a short bench line
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "library_sampler or seeded" > gpurun_out/$1_pytest_sampler.log 2>&1; echo "pytest rc $?" >> gpurun_out/$1_pytest_sampler.log
tail -4 gpurun_out/$1_pytest_sampler.log
python examples/run_example.py > gpurun_out/$1_example_run.txt 2>&1; tail -12 gpurun_out/$1_example_run.txt
PSFMC_NATIVE_SAMPLER=0 python examples/run_example.py 2>&1 | grep "model setup" 
python tools/example_breakdown.py > gpurun_out/$1_example_breakdown.txt 2>&1; head -30 gpurun_out/$1_example_breakdown.txt
python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/$1_bench_n1.json 2> gpurun_out/$1_bench_n1.err; tail -2 gpurun_out/$1_bench_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/$1_bench_n1.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"loop",d["e2e"]["sampler_loop"])
PY
python tools/time_host_call.py 125 2>&1 | tail -8
