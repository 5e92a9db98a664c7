# kernel variants built on the box (PSFMC_NVCC_EXTRA): C1 kernel time and C4 rate each
tag=$1; shift
run() {
  python bench.py --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_$1_c1.json 2>/dev/null
  python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_$1_c4.json 2>/dev/null
  python - <<PY
import json
a=json.loads(open("gpurun_out/${tag}_$1_c1.json").read().strip().splitlines()[-1])
b=json.loads(open("gpurun_out/${tag}_$1_c4.json").read().strip().splitlines()[-1])
print("$1: C1 value",a["value"],"kernel_us",a["roofline"]["kernel_us_per_launch"],"e2e",a["e2e"]["value"],"| C4 value",b["value"],"kernel_us",b["roofline"]["kernel_us_per_launch"])
PY
}
run asbuilt
for flags in "$@"; do
  PSFMC_NVCC_EXTRA="$flags" python -c "import __graft_entry__ as g; g.build(force=True)" > /dev/null 2>&1
  run $(echo "$flags" | tr -c 'A-Za-z0-9' '_')
done
