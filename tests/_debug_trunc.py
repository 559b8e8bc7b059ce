"""Development helper (GPU box): fp32 step-kernel trajectories against the oracle (relative to the largest latent),
per model / K / optimiser / engine mode -- the numbers behind the tolerances of tests/test_gpu_stepk.py and
test_gpu_parity.py::test_truncated_adagrad_unstaged_ring_fp32_k8."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import barbay_b200 as bb
from helpers import load_fixture, oracle_problem, rel_err
from oracle import advi_ref
n_steps = int(os.environ.get("DN", "9"))
MODES = {"persist": {}, "persist4": {"BB_PERSIST": "4"}, "pair": {"BB_PERSIST": "0"}, "round1": {"BB_NO_STEPK": "1"},
         "round1_unfused": {"BB_NO_STEPK": "1", "BB_NO_FUSE": "1"}}
for model in ("fitness_normal", "multienv_fitness_normal"):
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df, **cols)
    for K in (2, 8, 3):
        for opt in ("decayed", "truncated"):
            out = []
            tr = None
            for mode, env in MODES.items():
                for k in ("BB_PERSIST", "BB_NO_STEPK", "BB_NO_FUSE"):
                    os.environ.pop(k, None)
                os.environ["BB_STEPK_ODD"] = "1"
                os.environ.update(env)
                eng = bb.Engine(da, model, n_samples=K, dtype="f32", seed=1234)
                eng.init_params(5)
                mu0, om0 = eng.get_params()
                if opt == "truncated":
                    eng.set_optimizer("truncated", eta=0.1, tau=1.0, n=3); ro = advi_ref.TruncatedADAGrad(0.1, 1.0, 3)
                else:
                    eng.set_optimizer("decayed", eta=0.1, pre=1.0, post=0.9); ro = advi_ref.DecayedADAGrad(0.1, 1.0, 0.9)
                if tr is None:
                    tr = advi_ref.advi_run(model, oracle_problem(da, model), n_steps, K, ro, mu0, om0, seed=1234)
                eng.step(n_steps)
                mu, om = eng.get_params()
                out.append(f"{mode} {rel_err(mu, tr.mu):.1e}/{rel_err(om, tr.omega):.1e}")
                eng.close()
            print(model[:8], K, opt[:5], " | ".join(out))
