import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from radiobear_b200 import _lib
ctx=_lib.get_context(0)
rng=np.random.default_rng(1)
x=10**rng.uniform(-12,30,1<<20)*rng.choice([1.0],1<<20)
for nw in (0,1,2):
    y=ctx.probe_rcp(x,nw)
    rel=np.abs(y*x-1.0)
    print('newton',nw,'max rel err',rel.max(),'mean',rel.mean())
print('fp64 peak', ctx.fp64_peak_tflops(20000))
