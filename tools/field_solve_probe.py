import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
import pic_b200
L=50.0; B=4096; M=250
for N in (32, 512, 1250, 2500, 5000):
    rng=np.random.RandomState(0)
    x=rng.uniform(0,L,size=(B,N)); v=rng.normal(size=(B,N))
    eng=pic_b200.Engine(N,M,L,0.05,n_envs=B,mode="resident",max_mode=3)
    act=pic_b200.E_field(L,M,3); eng.set_actuator_basis(act.basis_cos,act.basis_sin)
    eng.set_tuning(512,0,-1)
    eng.set_state(x,v)
    c=torch.as_tensor(rng.uniform(-1,1,size=(20,B,6)),device="cuda")
    eng.step_coeffs_device(c.data_ptr(),20); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): eng.step_coeffs_device(c.data_ptr(),20)
    e1.record(); torch.cuda.synchronize()
    print(N, "%.4f ms/step"%(e0.elapsed_time(e1)/60), eng.launch_info()); eng.close()
