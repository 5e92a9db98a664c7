# Round-2 closing lines on one B200 (after the device loop; kernels as in gpu_final_r2b.sh):
# everything lands in gpurun_out/r2c_*. The N-GPU lines: tools/gpu_run8.sh under gpurun --gpus 8,
# the N = 2 profile of the sharded loop: tools/time_sampler_loop_sharded.py under torchrun.
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2c_pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2c_bench_c1_reference_arm.json 2>/dev/null
python bench.py --steps 300 --warmup 5 > gpurun_out/r2c_bench_c1.json 2> gpurun_out/r2c_bench_c1.err
python bench.py --walkers 65536 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_c1_65536walkers.json 2>/dev/null
python bench.py --walkers 200 --steps 1000 --warmup 20 --no-cpu-baseline > gpurun_out/r2c_bench_c1_200walkers.json 2>/dev/null
( for w in 250 1000 4096 16384 65536; do python tools/time_sampler_loop.py $w 100 ball; done
  python tools/time_sampler_loop.py 4096 100 prior
  PSFMC_DEVICE_LOOP=0 python tools/time_sampler_loop.py 4096 100 ball
  PSFMC_DEVICE_LOOP=1 python tools/time_sampler_loop.py 250 100 ball ) > gpurun_out/r2c_sampler_loop.txt 2>&1
python examples/run_example.py > gpurun_out/r2c_example_run.txt 2>&1
