"""Development helper (GPU box): where do NaNs first appear in the documented naive-prior workflow?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import barbay_b200 as bb
from helpers import load_fixture
df, _ = load_fixture("fitness_normal")
pri = bb.stats.prior_matrices(bb.stats.naive_prior(df.copy()))
da = bb.utils.data_to_arrays(df)
for tag, env, dtype, K, opt in [("default", {}, "f32", 1, "truncated"), ("nofuse", {"BB_NO_FUSE": "1"}, "f32", 1, "truncated"),
                                ("f64", {}, "f64", 1, "truncated"), ("K2", {}, "f32", 2, "truncated"),
                                ("decayed", {}, "f32", 1, "decayed"), ("vecprior", {"VEC": "1"}, "f32", 1, "truncated")]:
    for k in ("BB_NO_FUSE", "VEC"):
        os.environ.pop(k, None)
    os.environ.update(env)
    eng = bb.Engine(da, "fitness_normal", None if env.get("VEC") else pri, n_samples=K, dtype=dtype, seed=11)
    eng.init_params(11)
    eng.set_optimizer(opt)
    first = None
    for blk in range(30):
        eng.step(100)
        mu, om = eng.get_params()
        bad = ~np.isfinite(mu) | ~np.isfinite(om)
        if bad.any():
            first = (blk + 1) * 100
            idx = np.flatnonzero(bad)
            print(tag, "NaN by step", first, "count", bad.sum(), "first idx", idx[:8], "D", eng.D, eng.data_plane())
            break
    if first is None:
        print(tag, "ok: finite after 3000 steps; max|mu|", np.abs(mu).max(), "max omega", om.max(), "min omega", om.min())
    eng.close()
