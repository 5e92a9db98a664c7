# all GPU tests, the example run (+ breakdown), a short bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/$1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/$1_pytest.log
tail -3 gpurun_out/$1_pytest.log
python examples/run_example.py > gpurun_out/$1_example_run.txt 2>&1; grep "model setup" gpurun_out/$1_example_run.txt
python tools/example_breakdown.py > gpurun_out/$1_example_breakdown.txt 2>&1; sed -n 5,22p gpurun_out/$1_example_breakdown.txt
python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/$1_bench_n1.json 2> gpurun_out/$1_bench_n1.err; tail -2 gpurun_out/$1_bench_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/$1_bench_n1.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"priors",d["e2e"]["with_python_priors"],"map",d["e2e"]["pool_map"],d["e2e"]["pool_map_inside"],"kernel_us",d["roofline"]["kernel_us_per_launch"],"frac",d["roofline"]["frac"])
print("loop",d["e2e"]["sampler_loop"])
PY
