#!/usr/bin/env python
"""Smallest run of every kernel for compute-sanitizer: a handful of C1 walkers through
the fused path (float32) and the staged paths (float32 + float64), plus blob images.
    compute-sanitizer --tool racecheck python tools/sanitize_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import load_golden, model_from_file   # noqa: E402

golden = load_golden('c1_golden.json')
thetas = np.array(golden['theta'][:6])
os.environ['PSFMC_FUSED_CTAS'] = '2'          # 3 walkers per CTA: exercises the walker loop
fused = model_from_file('j0005/model_c1.py', 'fp32')
print('fused', fused.engine.info()['path'], fused.log_likelihood_batch(thetas))
os.environ['PSFMC_FORCE_STAGED'] = '1'
staged = model_from_file('j0005/model_c1.py', 'fp32')
print('staged fp32', staged.log_likelihood_batch(thetas[:3]))
staged64 = model_from_file('j0005/model_c1.py', 'fp64')
print('staged fp64', staged64.log_likelihood_batch(thetas[:2]))
imgs = staged64.engine.render(thetas[:1])
print('images', sorted(imgs))
