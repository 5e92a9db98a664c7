tag=$1; shift
for flags in "$@"; do
  PSFMC_NVCC_EXTRA="$flags" python -c "import __graft_entry__ as g; g.build(force=True)" > /dev/null 2>&1
  name=$(echo "$flags" | tr -c 'A-Za-z0-9' '_')
  python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_${name}.json").read().strip().splitlines()[-1])
print("$flags", "value",d["value"],"raw e2e",d["e2e"]["without_fp64_rescue"])
PY
done
