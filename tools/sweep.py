#!/usr/bin/env python
"""
Ensemble-size sweep (BASELINE.json configs[4]): lnL evaluations/s for ensembles of
200 ... 65536 walkers on the 128^2 (C1 model), 256^2 (C3) and 512^2 (C4) frames, one GPU,
device-resident (CUDA events) and end to end through the C ABI with host buffers.
Each emcee iteration is two half-ensemble batches, as in bench.py.

    python tools/sweep.py > profiles/rN_sweep.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(workload, walkers, seconds=1.0):
    import torch
    import bench
    from psfmc_b200 import MultiComponentModel
    from psfmc_b200.synthetic import draw_walkers_fast
    model = MultiComponentModel(bench.build_components(workload), precision='fp32',
                                devices=[0])
    engine = model.engine
    ndim, half = model.num_params, walkers // 2
    base = draw_walkers_fast(model, min(walkers, 4096), seed=walkers)
    thetas = np.ascontiguousarray(np.resize(base, (walkers, ndim)))
    dev = torch.device('cuda', 0)
    th_dev = torch.from_numpy(thetas).to(dev)
    th_pin = torch.from_numpy(thetas).pin_memory()
    lnl_dev = torch.empty(walkers, dtype=torch.float64, device=dev)
    lnl_pin = torch.empty(walkers, dtype=torch.float64).pin_memory()
    stream = torch.cuda.current_stream(dev)

    def step_device():
        for h in range(2):
            engine.lnlike_device(th_dev.data_ptr() + h * half * ndim * 8, half, ndim,
                                 lnl_dev.data_ptr() + h * half * 8, stream=stream.cuda_stream)

    def step_host():
        th, out = th_pin.numpy(), lnl_pin.numpy()
        for h in range(2):
            engine.lnlike(th[h * half:(h + 1) * half], out=out[h * half:(h + 1) * half])

    for _ in range(3):
        step_device()
        step_host()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    step_device()
    e1.record(stream)
    e1.synchronize()
    steps = int(max(3, min(2000, seconds * 1e3 / max(e0.elapsed_time(e1), 1e-3))))
    e0.record(stream)
    for _ in range(steps):
        step_device()
    e1.record(stream)
    e1.synchronize()
    dev_rate = walkers * steps / (e0.elapsed_time(e1) * 1e-3)
    t0 = time.perf_counter()
    for _ in range(steps):
        step_host()
    host_rate = walkers * steps / (time.perf_counter() - t0)
    info = engine.info()
    model.engine.close()
    return {'workload': workload, 'frame': list(engine.shape), 'walkers': walkers,
            'batch_per_launch': half, 'steps': steps,
            'device_evals_per_s': round(dev_rate, 1), 'e2e_evals_per_s': round(host_rate, 1),
            'engine_path': {1: 'fused', 2: 'fused-cluster4', 3: 'tiled-4x4'}.get(info['path'], 'staged'),
            'fp64_rescued_walkers': int(info['rescued_total'])}


def main():
    import __graft_entry__ as entry
    entry.build()
    rows = []
    for workload in ('c1', 'c3', 'c4'):
        for walkers in (200, 512, 1024, 4096, 16384, 65536):
            rows.append(measure(workload, walkers))
            print(json.dumps(rows[-1]), file=sys.stderr)
    print(json.dumps({'sweep': rows, 'n_gpus': 1, 'unit': 'evals/s'}, indent=1))


if __name__ == '__main__':
    main()
