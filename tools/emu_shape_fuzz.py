#!/usr/bin/env python
"""
Random frame shapes on the CPU emulator (tests/emu) against the oracle: heights 8 ... 139,
even widths 8 ... 138, PSF stamps from 1 x 1 up to the frame (at most 64 x 64) -- the
zero-padded transform frame, the fold back and the kernel origin (psfMC/utils.py:9-32) for
whatever shape a user's cut-out has. tests/conftest.py: arbitrary_frame_model (asymmetric
PSF, a point source in the frame corner whose wings wrap around, bad pixels, a NaN).

    python tools/emu_shape_fuzz.py [seed] [n_shapes]

Test tool: imports the oracle; no GPU needed.
"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import numpy as np
import conftest
from psfmc_b200.synthetic import draw_walkers_fast
rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
worst64 = worst32 = 0
for it in range(n):
    H = int(rng.randint(8, 140)); W = int(2 * rng.randint(4, 70))
    ph = int(rng.randint(1, min(H, 64) + 1)); pw = int(rng.randint(1, min(W, 64) + 1))
    t0 = time.time()
    try:
        m64 = conftest.arbitrary_frame_model(H, W, ph, pw, precision='fp64', library=conftest.EMU_LIB)
    except Exception as e:
        print((H, W, ph, pw), 'create raised', type(e).__name__, str(e)[:150]); continue
    th = draw_walkers_fast(m64, 2, seed=it)
    orc = conftest.oracle_from_model(m64)
    with np.errstate(all='ignore'):
        exp = orc.lnlike_batch(th)
    got = m64.log_likelihood_batch(th)
    m32 = conftest.arbitrary_frame_model(H, W, ph, pw, precision='fp32', library=conftest.EMU_LIB, fp64_rescue=True)
    got32 = m32.log_likelihood_batch(th)
    b = conftest.fp32_bounds(m32, th, orc)
    fin = np.isfinite(exp)
    ok_f = np.array_equal(np.isfinite(got), fin) and np.array_equal(np.isfinite(got32), fin)
    r64 = np.max(np.abs(got[fin]-exp[fin])/np.abs(exp[fin])) if fin.any() else 0
    r32 = np.max(np.abs(got32[fin]-exp[fin])/b[fin]) if fin.any() else 0
    worst64 = max(worst64, r64); worst32 = max(worst32, r32)
    flag = '' if (ok_f and r64 < 1e-10 and r32 < 1) else '   <<<<<<<<'
    print((H, W, ph, pw), 'path', m64.engine.info()['path'], m32.engine.info()['path'], 'fin ok', ok_f, 'rel64 %.2g  err32/bound %.2g  (%.1fs)%s' % (r64, r32, time.time()-t0, flag))
    m64.engine.close(); m32.engine.close()
print('worst', worst64, worst32)
