#!/usr/bin/env python
"""
The psfMC example run (BASELINE.json configs[0]: J0005-0006 quasar + host, burn 200,
200 retained iterations, 250 walkers as in the reference's examples/run_example.py)
on the GPU engine: sampling, FITS trace database, posterior images.

    python examples/run_example.py [--chains 250] [--burn 200] [--iterations 200]
                                   [--outdir /tmp/psfmc_example] [--seed 1]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from psfmc_b200 import MultiComponentModel, model_galaxy_mcmc      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--chains', type=int, default=250)
    ap.add_argument('--burn', type=int, default=200)
    ap.add_argument('--iterations', type=int, default=200)
    ap.add_argument('--outdir', default='/tmp/psfmc_example')
    ap.add_argument('--seed', type=int, default=1)
    args = ap.parse_args()
    os.makedirs(args.outdir, exist_ok=True)
    model_file = os.path.join(ROOT, 'examples', 'model_J0005-0006.py')
    output = os.path.join(args.outdir, 'out_J0005-0006')
    for name in os.listdir(args.outdir):
        if name.startswith('out_J0005-0006'):
            os.remove(os.path.join(args.outdir, name))
    start = time.perf_counter()
    model = MultiComponentModel(model_file)
    setup = time.perf_counter() - start
    start = time.perf_counter()
    database = model_galaxy_mcmc(model, output_name=output, burn=args.burn,
                                 iterations=args.iterations, chains=args.chains,
                                 seed=args.seed, verbose=False)
    elapsed = time.perf_counter() - start
    evals = args.chains * (args.burn + args.iterations + 2)
    print('model setup {:.2f} s; {} posterior evaluations + FITS outputs in {:.2f} s '
          '({:.0f} evals/s end to end incl. Python priors, sampler and I/O)'.format(
              setup, evals, elapsed, evals / elapsed))
    print('acceptance fraction {:.3f}, converged {}'.format(
        database.meta['MCACCEPT'], database.meta['MCCONVRG']))
    for name in ('0_Sky_adu', '1_PointSource_mag', '2_Sersic_mag', '2_Sersic_index',
                 '2_Sersic_reff'):
        col = np.asarray(database[name])
        print('  {:20s} median {:10.4f}  16-84% [{:.4f}, {:.4f}]'.format(
            name, np.median(col), np.percentile(col, 16), np.percentile(col, 84)))
    print('outputs:', sorted(os.listdir(args.outdir)))


if __name__ == '__main__':
    main()
