for mb in 512 74 148 37; do
  PSFMC_CHUNK_MB=$mb python bench.py --workload c4 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/c4_chunk_$mb.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/c4_chunk_$mb.json").read().strip().splitlines()[-1])
print("chunk MB $mb value",d["value"],"e2e",d["e2e"]["value"],"raw",d["e2e"]["without_fp64_rescue"])
PY
done
