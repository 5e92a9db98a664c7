import sys, numpy as np
sys.path.insert(0,'/root/repo')
import bench
from psfmc_b200 import MultiComponentModel
m32 = MultiComponentModel(bench.build_components('c1'), precision='fp32', fp64_rescue=False)
m64 = MultiComponentModel(bench.build_components('c1'), precision='fp64')
tot=0
for s in range(4):
    th = bench.ensemble(m32, 4096, s)
    a = m32.log_likelihood_batch(th); b = m64.log_likelihood_batch(th)
    bad = np.flatnonzero(~np.isfinite(a))
    for i in bad:
        t = th[i]
        def d(x): return abs(x-round(x))
        print('set',s,'row',i,'fp64',b[i],'idx',round(t[5],3),round(t[12],3),'centre offs',round(np.hypot(d(t[9]),d(t[10])),4),round(np.hypot(d(t[16]),d(t[17])),4))
    tot+=len(bad)
print('total nonfinite fp32', tot, 'of', 4*4096)
