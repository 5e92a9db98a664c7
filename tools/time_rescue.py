"""Cost of the float64 repeat of non-finite float32 walkers (C1, 2048-walker calls):
time per psfmc_lnlike_batch call for a batch without / with one / with four such
walkers, graph path (conditional node) against the plain launch path + host repeat.
Usage: python tools/time_rescue.py [calls]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from bench import build_components  # noqa: E402
from psfmc_b200 import MultiComponentModel  # noqa: E402
from psfmc_b200.synthetic import draw_walkers_fast  # noqa: E402
from conftest import HIGH_DYNAMIC_RANGE_THETAS  # noqa: E402


def main():
    calls = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    half = 2048
    rows = {}
    for mode in ('graph', 'plain'):
        if mode == 'plain':
            os.environ['PSFMC_NO_GRAPH'] = '1'
        else:
            os.environ.pop('PSFMC_NO_GRAPH', None)
        model = MultiComponentModel(build_components('c1'), precision='fp32', devices=[0])
        engine = model.engine
        thetas = draw_walkers_fast(model, 3 * half, seed=5)
        lnl = engine.lnlike(thetas)
        clean = thetas[np.isfinite(lnl)][:half]
        assert len(clean) == half
        hdr = np.array(HIGH_DYNAMIC_RANGE_THETAS)
        batches = {'0 non-finite': clean.copy()}
        one = clean.copy()
        one[100] = hdr[0]
        batches['1 non-finite'] = one
        four = clean.copy()
        four[[7, 300, 1200, 2000]] = hdr[[0, 1, 0, 1]]
        batches['4 non-finite'] = four
        out = torch.empty(half, dtype=torch.float64).pin_memory().numpy()
        for name, batch in batches.items():
            pin = torch.from_numpy(batch).pin_memory().numpy()
            for _ in range(10):
                engine.lnlike(pin, out=out)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(calls):
                engine.lnlike(pin, out=out)
            dt = (time.perf_counter() - t0) / calls
            rows[(mode, name)] = dt * 1e6
            print('{:6s} {:14s} {:8.1f} us/call  ({} non-finite left)'.format(
                mode, name, dt * 1e6, int(np.sum(~np.isfinite(out)))), flush=True)
        print(mode, engine.info())
        engine.close()


if __name__ == '__main__':
    main()
