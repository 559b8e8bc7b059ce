"""Development helper (GPU box): fp32 trajectories of every model family against the oracle, both optimisers,
one call of n steps (relative to the largest latent; worst latent's reference index and group)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import barbay_b200 as bb
from helpers import load_fixture, oracle_problem, rel_err
from oracle import advi_ref
n_steps, K = int(os.environ.get("DN", "12")), int(os.environ.get("DK", "4"))
for model in ("fitness_normal", "replicate_fitness_normal", "multienv_fitness_normal", "genotype_fitness_normal"):
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df, **cols)
    for opt in ("decayed", "truncated"):
        eng = bb.Engine(da, model, n_samples=K, dtype="f32", seed=1234)
        eng.init_params(5)
        mu0, om0 = eng.get_params()
        if opt == "truncated":
            eng.set_optimizer("truncated", eta=0.1, tau=1.0, n=3); ro = advi_ref.TruncatedADAGrad(0.1, 1.0, 3)
        else:
            eng.set_optimizer("decayed", eta=0.1, pre=1.0, post=0.9); ro = advi_ref.DecayedADAGrad(0.1, 1.0, 0.9)
        tr = advi_ref.advi_run(model, oracle_problem(da, model), n_steps, K, ro, mu0, om0, seed=1234)
        eng.step(n_steps)
        mu, om = eng.get_params()
        dm, do = np.abs(mu - tr.mu), np.abs(om - tr.omega)
        names = eng.var_names() if hasattr(eng, "var_names") else None
        print(f"{model[:10]:10s} {opt[:5]} D={mu.size} rel mu {rel_err(mu, tr.mu):.1e} om {rel_err(om, tr.omega):.1e} | worst mu idx {int(dm.argmax())} ({dm.max():.1e}) om idx {int(do.argmax())} ({do.max():.1e})")
        eng.close()
