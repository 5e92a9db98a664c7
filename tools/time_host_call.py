"""Host-timed anatomy of one 2048-walker C1 call: the blocking C-ABI call with pinned
host buffers against the same kernels on device-resident theta (launch + sync only),
and the device time of the kernels (CUDA events).
Usage: python tools/time_host_call.py [walkers per call = 2048]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_components  # noqa: E402
from psfmc_b200 import MultiComponentModel  # noqa: E402
from psfmc_b200.synthetic import draw_walkers_fast  # noqa: E402


def main():
    half = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    calls = 300 if half >= 1024 else 1000
    model = MultiComponentModel(build_components('c1'), precision='fp32', devices=[0],
                                fp64_rescue=os.environ.get('PSFMC_GRAPH_ALWAYS') == '1')
    engine = model.engine
    th = draw_walkers_fast(model, half, seed=5)
    ndim = th.shape[1]
    print('walkers per call', half)
    pin = torch.from_numpy(th).pin_memory()
    out = torch.empty(half, dtype=torch.float64).pin_memory()
    dev = torch.from_numpy(th).cuda()
    lnl_dev = torch.empty(half, dtype=torch.float64, device='cuda')
    stream = torch.cuda.current_stream()

    def host_call():
        engine.lnlike(pin.numpy(), out=out.numpy())

    def device_call():
        engine.lnlike_device(dev.data_ptr(), half, ndim, lnl_dev.data_ptr(),
                             stream=stream.cuda_stream)
        stream.synchronize()

    def h2d_only():
        dev.copy_(pin, non_blocking=True)
        stream.synchronize()

    for name, func in (('C-ABI call, pinned host buffers', host_call),
                       ('device-resident theta + sync', device_call),
                       ('H2D copy of theta + sync', h2d_only)):
        for _ in range(20):
            func()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(calls):
            func()
        print('{:36s} {:8.1f} us'.format(name, (time.perf_counter() - t0) / calls * 1e6))
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(calls):
        engine.lnlike_device(dev.data_ptr(), half, ndim, lnl_dev.data_ptr(),
                             stream=stream.cuda_stream)
    e1.record(stream)
    e1.synchronize()
    print('{:36s} {:8.1f} us'.format('kernels back to back (events)',
                                     e0.elapsed_time(e1) / calls * 1e3))


if __name__ == '__main__':
    main()
