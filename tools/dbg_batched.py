import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200
L=50.0; B=4096
def run(tag, sampler, T, tune=None):
    bp = pic_b200.Engine(5000, 250, L, 0.05, n_envs=B, mode="resident", deposit="split32", max_mode=3)
    act = pic_b200.E_field(L, 250, 3); bp.set_actuator_basis(act.basis_cos, act.basis_sin)
    if tune: bp.set_tuning(*tune)
    if sampler: bp.sample_state("bump-on-tail", seed=7, n_global=5000)
    else:
        rng=np.random.RandomState(0)
        bp.set_state(rng.uniform(0,L,(B,5000)), rng.normal(size=(B,5000)))
    coeffs = torch.rand(T, B, 6, dtype=torch.float64, device="cuda")*2-1
    bp.step_coeffs_device(coeffs.data_ptr(), T); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): bp.step_coeffs_device(coeffs.data_ptr(), T)
    e1.record(); torch.cuda.synchronize()
    print(tag, bp.launch_info(), "ms/step %.4f"%(e0.elapsed_time(e1)/(5*T)), "flags", bp.error_flags(), flush=True)
    bp.close()
run("sampler T10 default", True, 10)
run("random  T10 default", False, 10)
run("sampler T20 default", True, 20)
run("sampler T10 tuned512", True, 10, (512,0,-1))
run("random  T20 tuned512", False, 20, (512,0,-1))
