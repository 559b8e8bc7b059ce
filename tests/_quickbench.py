import sys, json, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import barbay_b200 as bb
K = int(os.environ.get("QK", "8"))
model, da, _ = bb.synth.config(2)
eng = bb.Engine(da, model, n_samples=K, dtype="f32", seed=1, device=0)
eng.init_params(1); eng.set_optimizer(os.environ.get("QOPT", "decayed"))
eng.step(20); eng.sync()
tot, p1, p2 = eng.time_steps(200)
n=200
print(json.dumps({"K": K, "acc_kb": os.environ.get("BB_P1_ACC_KB"), "step_us": tot/n*1e3, "p1_us": p1/n*1e3, "p2_us": p2/n*1e3, "units_per_s": 5e6*K*n/(tot*1e-3), "alg_GBs_step": eng.algorithmic_bytes_per_step/(tot/n*1e-3)/1e9}))
