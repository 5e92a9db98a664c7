#!/usr/bin/env python
"""Where the wall time of psfMC's own example run goes (examples/run_example.py: burn 200 +
200 iterations x 250 walkers, trace database, posterior images): cProfile over
model_galaxy_mcmc, cumulative seconds of the package's own functions and of the C-ABI calls.
    python tools/example_breakdown.py > profiles/rN_example_breakdown.txt"""
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from psfmc_b200 import MultiComponentModel, model_galaxy_mcmc
    model_file = os.path.join(ROOT, 'examples', 'model_J0005-0006.py')
    outdir = tempfile.mkdtemp(prefix='psfmc_example_')
    model = MultiComponentModel(model_file)
    # warm-up (library load, first-batch prior validation, graph capture)
    model_galaxy_mcmc(model, output_name=os.path.join(outdir, 'warm'), burn=2,
                      iterations=4, chains=250, seed=2, verbose=False)
    model = MultiComponentModel(model_file)
    prof = cProfile.Profile()
    start = time.perf_counter()
    prof.enable()
    model_galaxy_mcmc(model, output_name=os.path.join(outdir, 'out'), burn=200,
                      iterations=200, chains=250, seed=1, verbose=False)
    prof.disable()
    elapsed = time.perf_counter() - start
    evals = 250 * 402
    print('{} posterior evaluations + outputs in {:.3f} s under cProfile ({:.0f} evals/s)'
          .format(evals, elapsed, evals / elapsed))
    start = time.perf_counter()
    model = MultiComponentModel(model_file)
    setup = time.perf_counter() - start
    start = time.perf_counter()
    model_galaxy_mcmc(model, output_name=os.path.join(outdir, 'plain'), burn=200,
                      iterations=200, chains=250, seed=1, verbose=False)
    plain = time.perf_counter() - start
    print('the same without the profiler: {:.3f} s ({:.0f} evals/s); model setup {:.2f} s'
          .format(plain, evals / plain, setup))
    stream = io.StringIO()
    stats = pstats.Stats(prof, stream=stream).sort_stats('cumulative')
    stats.print_stats(45)
    keep = []
    for line in stream.getvalue().splitlines():
        if ('psfmc_b200' in line or 'ncalls' in line or '_lib' in line or 'numpy' in line
                or 'method' in line or 'built-in' in line):
            keep.append(line.replace(ROOT + '/', ''))
    print('\n'.join(keep[:48]))


if __name__ == '__main__':
    main()
