"""CPU, world_size 2, gloo: the barcode-sharded two-pass algebra (what each GPU rank computes around
the per-step all-reduce) reproduces the oracle's full gradient.  Covers the N>1 host logic without a GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, z, R, T, N, M, out_path):
    import shard_algebra as sa
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = torch.from_numpy(sa.pass1_partials(z, T, N, M, rank, world))
    dist.all_reduce(part)                                   # the one collective of a step
    pri = {"s_pop": (0.0, 2.0), "lsig_pop": (0.0, 1.0), "s_bc": (0.0, 2.0), "lsig_bc": (0.0, 1.0), "lam": (3.0, 3.0)}
    g = torch.from_numpy(sa.gradients_from_sums(z, part.numpy(), R, T, N, M, rank, world, pri))
    dist.all_reduce(g)                                      # assemble owned pieces (what get_posterior's caller does)
    if rank == 0:
        np.save(out_path, g.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_partials_and_gradient_match_oracle(bb, tmp_path):
    from helpers import load_fixture, oracle_problem, plausible_latents
    from oracle import model_ref
    model = "fitness_normal"
    df, cols = load_fixture(model)
    da = bb.utils.data_to_arrays(df, **cols)
    lay = bb.model.var_groups(bb.model.fitness_normal, da.n_time, 1, da.n_neutral, da.n_bc)
    z = plausible_latents(lay, da, np.random.default_rng(0), 1)[0]
    R = np.asarray(da.bc_count).astype(np.float64)
    out = str(tmp_path / "grad.npy")
    port = 29531 + (os.getpid() % 200)
    mp.spawn(_worker, args=(2, port, z, R, R.shape[0], da.n_neutral, da.n_bc, out), nprocs=2, join=True)
    g = np.load(out)
    _, g_ref = model_ref.logjoint_and_grad(model, z, oracle_problem(da, model))
    assert np.max(np.abs(g - g_ref)) <= 1e-9 * np.max(np.abs(g_ref))


def test_shard_ranges_partition_everything():
    import shard_algebra as sa
    for N, M, world in [(5, 10, 2), (1000, 999000, 8), (3, 7, 4), (1, 1, 2)]:
        ns = [sa.shard_ranges(N, M, r, world) for r in range(world)]
        assert ns[0][0][0] == 0 and ns[-1][0][1] == N and ns[0][1][0] == 0 and ns[-1][1][1] == M
        for a, b in zip(ns[:-1], ns[1:]):
            assert a[0][1] == b[0][0] and a[1][1] == b[1][0]
