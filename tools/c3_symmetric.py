import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qwtdh_b200 as qw
n=4096
t=np.linspace(0.1,512.0,256); b=np.linspace(0.5/4096,2.5/4096,256)
dt=torch.tensor(t,device="cuda"); db=torch.tensor(b,device="cuda")
for flags in (8192,0):
    prob=qw.Problem(n=n,topology="complete",schedule="lin",gamma_on="laplacian",n_steps=2048,flags=flags)
    for rep in range(3):
        torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); p,nrm=qw.rk4_heatmap(prob,dt,db); e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)
    print(flags, qw.plan(prob)["path"], "%.3f ms"%ms, "%.3e pts/s"%(65536/ms*1e3))
    if flags: pm=p.clone()
print("max |dp|", float((pm-p).abs().max()), "argmax equal", int(pm.argmax())==int(p.argmax()))
