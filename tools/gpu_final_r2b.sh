# Round-2 (second half) measurement suite on one B200: everything lands in gpurun_out/r2b_*.
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2b_pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2b_bench_c1_reference_arm.json 2> gpurun_out/r2b_bench_ref.err
python bench.py --steps 300 --warmup 5 > gpurun_out/r2b_bench_c1.json 2> gpurun_out/r2b_bench_c1.err
python bench.py --workload c3 --steps 100 --warmup 5 > gpurun_out/r2b_bench_c3.json 2> gpurun_out/r2b_bench_c3.err
python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/r2b_bench_c4.json 2> gpurun_out/r2b_bench_c4.err
python bench.py --walkers 200 --steps 2000 --warmup 20 --no-cpu-baseline > gpurun_out/r2b_bench_c1_200walkers.json 2>/dev/null
python bench.py --walkers 65536 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_c1_65536walkers.json 2>/dev/null
( for w in 250 4096 16384; do python tools/time_sampler_loop.py $w 100 ball; done; python tools/time_sampler_loop.py 4096 100 prior ) > gpurun_out/r2b_sampler_loop.txt 2>&1
python examples/run_example.py > gpurun_out/r2b_example_run.txt 2>&1
python tools/example_breakdown.py > gpurun_out/r2b_example_breakdown.txt 2>&1
( python tools/time_host_call.py 2048; python tools/time_host_call.py 125; python tools/time_pool_path.py ) > gpurun_out/r2b_host_call_anatomy.txt 2>&1
python tools/tolerance_audit.py > gpurun_out/r2b_fp32_tolerance_audit.json 2> gpurun_out/r2b_tolerance_audit.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_fused_c1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:fused_lnlike -c 1 --launch-skip 8 -o gpurun_out/r2b_prof_fused python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -3 gpurun_out/r2b_pytest_gpu.log
ls -la gpurun_out/r2b_*
