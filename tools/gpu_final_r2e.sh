# What the commits after the round's last GPU session still owe a B200 (none of them touches a
# kernel; everything was verified on the CPU emulator / against the unmodified reference):
#   * tests/test_gpu_parity.py::test_gpu_wide_box_fuzz has never run on hardware;
#   * the bulk MT19937 stream of the sampler loops (host code) is measured on the build
#     container's host only -- its effect on the sampler loops, N = 1 and sharded, is open;
#   * BatchPool / ShardedPool probe their first ensemble once per pool (one extra host call
#     in the warm-up of bench.py's sampler_loop / pool_map legs): no effect on any timed region
#     is expected; the N-GPU lines should reproduce profiles/r2c_* and r2d_*.
# Everything lands in gpurun_out/r2e_*.  One GPU:   gpurun --timeout 1500 -- bash tools/gpu_final_r2e.sh
# N GPUs (N = 2, 4, 8):  gpurun --gpus N -- 'python -m torch.distributed.run --nnodes=1 \
#     --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus N \
#     --steps 300 --warmup 5 > gpurun_out/r2e_bench_c1_${N}gpu.json'
set -x
python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2e_pytest_gpu.log
python bench.py --steps 300 --warmup 5 > gpurun_out/r2e_bench_c1.json 2> gpurun_out/r2e_bench_c1.err
( for w in 250 1000 4096 16384; do python tools/time_sampler_loop.py $w 100 ball; done
  PSFMC_DEVICE_LOOP=0 python tools/time_sampler_loop.py 4096 100 ball ) > gpurun_out/r2e_sampler_loop.txt 2>&1
python examples/run_example.py > gpurun_out/r2e_example_run.txt 2>&1
