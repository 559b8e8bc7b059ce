"""Development helper (GPU box with N GPUs): the single-call multi-GPU path -- ONE handle (bb_desc.n_devices = N) in ONE
process driving N GPUs, what `BarBay.vi.advi(...; n_devices=N)` uses -- on cfg2, K = 8: step rate (host wall clock around
the blocking bb_step call), whole-call time and agreement with the single-GPU posterior.  Usage: _ndev_bench.py N"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import barbay_b200 as bb

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K, steps = 8, 2000
model, da, _ = bb.synth.config(2)
units = np.asarray(da.bc_count).size * K
out = {"n_devices": N, "K": K, "steps": steps}
post = {}
for nd in (1, N):
    t0 = time.time()
    eng = bb.Engine(da, model, n_samples=K, dtype="f32", seed=1, device=0, n_devices=nd)
    t_create = time.time() - t0
    eng.init_params(1); eng.set_optimizer("decayed")
    eng.step(50); eng.sync()
    t0 = time.time(); eng.step(steps); eng.sync(); dt = time.time() - t0
    t0 = time.time(); m, s = eng.get_posterior(); t_post = time.time() - t0
    post[nd] = (m, s)
    out[f"n{nd}"] = {"create_s": t_create, "step_us": dt / steps * 1e6, "units_per_s": units * steps / dt,
                     "get_posterior_s": t_post, "plane": eng.data_plane()}
    eng.close()
# fp32, 2050 steps: compare robustly (a few latents sit on sign-like AdaGrad steps)
d = np.abs(post[N][0] - post[1][0]) / (np.abs(post[1][0]) + 1.0)
out["posterior_mean_rel_median"] = float(np.median(d)); out["posterior_mean_rel_p999"] = float(np.quantile(d, 0.999))
out["speedup"] = out["n1"]["step_us"] / out[f"n{N}"]["step_us"]
print(json.dumps(out))
