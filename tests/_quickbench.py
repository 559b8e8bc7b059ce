import sys, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import barbay_b200 as bb
K = int(os.environ.get("QK", "8"))
cfg = int(os.environ.get("QCFG", "2"))
model, da, _ = bb.synth.config(cfg, scale=float(os.environ.get('QSCALE', '1.0')))
eng = bb.Engine(da, model, n_samples=K, dtype=os.environ.get("QDT", "f32"), seed=1, device=0)
eng.init_params(1); eng.set_optimizer(os.environ.get("QOPT", "decayed"))
eng.step(20); eng.sync()
n = int(os.environ.get("QN", "200"))
s = torch.cuda.Stream(); eng.set_stream(s.cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
eng.step(5); torch.cuda.synchronize()
e0.record(s); eng.step(n); e1.record(s); torch.cuda.synchronize()
plain = e0.elapsed_time(e1) / n * 1e3
tot, p1, p2 = eng.time_steps(n)
R = np.asarray(da.bc_count); units = R.size * K
print(json.dumps({"plane": eng.data_plane(), "persist": eng.persist_stats(), "cfg": cfg, "K": K, "opt": os.environ.get("QOPT", "decayed"), "nofuse": os.environ.get("BB_NO_FUSE"),
                  "step_us": plain, "bracketed_step_us": tot/n*1e3, "p1_us": p1/n*1e3, "p2_us": p2/n*1e3,
                  "units_per_s": units/(plain*1e-6), "alg_GBs_step": eng.algorithmic_bytes_per_step/(plain*1e-6)/1e9}))
