python -m pytest tests -m gpu -x -q > gpurun_out/$1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/$1_pytest.log
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/$1_bench_n1.json 2> gpurun_out/$1_bench_n1.err
tail -3 gpurun_out/$1_pytest.log
python - <<PY
import json
d=json.loads(open("gpurun_out/$1_bench_n1.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"raw",d["e2e"]["without_fp64_rescue"],"priors",d["e2e"]["with_python_priors"],"map",d["e2e"]["pool_map"],d["e2e"]["pool_map_inside"],"kernel_us",d["roofline"]["kernel_us_per_launch"],"frac",d["roofline"]["frac"],"share",d["roofline"]["kernel_share_of_step"])
PY
