import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import pic_b200
L=50.0
for (N,M) in ((5000,250),(10000,500),(1000,100)):
  for B in (1, 16, 148):
    rng=np.random.RandomState(0)
    x=rng.uniform(0,L,size=(B,N)); v=rng.normal(size=(B,N))
    for th in (256,512,1024):
        eng=pic_b200.Engine(N,M,L,0.05,n_envs=B,mode="resident")
        try: eng.set_tuning(th,0,-1)
        except Exception as e: print("skip",N,B,th); eng.close(); continue
        eng.set_state(x,v)
        eng.step_mesh(None,200); eng.sync()
        t0=time.perf_counter(); eng.step_mesh(None,2000); eng.sync()
        print("N=%5d B=%3d threads=%4d  %.2f us/step"%(N,B,th,(time.perf_counter()-t0)/2000*1e6), flush=True)
        eng.close()
