"""Development helper (GPU box): a short run of the production step for ncu captures.
   QK / QCFG / QDT / QOPT as in _quickbench.py; QSTEPS steps after set-up."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import barbay_b200 as bb
K = int(os.environ.get("QK", "8")); cfg = int(os.environ.get("QCFG", "2"))
model, da, _ = bb.synth.config(cfg, scale=float(os.environ.get("QSCALE", "1.0")))
eng = bb.Engine(da, model, n_samples=K, dtype=os.environ.get("QDT", "f32"), seed=1, device=0)
eng.init_params(1); eng.set_optimizer(os.environ.get("QOPT", "decayed"))
eng.step(int(os.environ.get("QSTEPS", "12"))); eng.sync()
print("ok", eng.data_plane())
