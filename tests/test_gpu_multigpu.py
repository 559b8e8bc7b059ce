"""GPU (needs >= 2 GPUs, skipped otherwise): sharded run == single-GPU run."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("model", ["fitness_normal", "replicate_fitness_normal", "multienv_fitness_normal",
                                   "genotype_fitness_normal"])
def test_two_gpu_run_matches_single_gpu(model):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_gpu_check.py"),
           model, "f64"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["all_owned"]
    # identical up to the reduction order of the partial sums (fp64)
    assert out["elbo_rel"] < 1e-10 and out["mean_rel"] < 1e-9 and out["std_rel"] < 1e-9, out
