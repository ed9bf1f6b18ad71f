import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fhmcanalysis_b200 import engine, synth
n=1001; lnpi=synth.two_peak_lnpi(n); N=np.arange(n,dtype=float)
def timed(fn, reps=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
for rec in (3,2):
    dh=engine.DeviceHistogram(lnpi,N,1.0,0.0,smooth=10,sel=["N",N*N]); dh.use_recurrence=rec; dh.ensure_hull()
    for S in (60000, 80000, 100000, 131072, 160000, 262144):
        mu=dh._dev_array(np.linspace(-0.03,0.03,S)); out=dh.sweep(mu,pmax=4)
        print(json.dumps({"rec":rec,"S":S,"us":1e3*timed(lambda: dh.sweep(mu,pmax=4,out=out))}), flush=True)
