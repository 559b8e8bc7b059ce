"""Development helper (GPU box): where the wall time of one advi()-equivalent call goes."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import barbay_b200 as bb
model, da, _ = bb.synth.config(2)
out = {}
for rep in range(2):
    t = [time.perf_counter()]
    eng = bb.Engine(da, model, n_samples=8, dtype="f32", seed=1, device=0); t.append(time.perf_counter())
    eng.init_params(1); eng.sync(); t.append(time.perf_counter())
    eng.set_optimizer("decayed"); eng.sync(); t.append(time.perf_counter())
    eng.step(2000); eng.sync(); t.append(time.perf_counter())
    tr = eng.step(1, elbo_trace=True); t.append(time.perf_counter())
    m, s = eng.get_posterior(); t.append(time.perf_counter())
    eng.close(); t.append(time.perf_counter())
    names = ["create", "init_params", "set_optimizer", "step2000", "elbo_step", "get_posterior", "close"]
    out[rep] = {n: round((t[i + 1] - t[i]) * 1e3, 1) for i, n in enumerate(names)}
print(json.dumps(out))
