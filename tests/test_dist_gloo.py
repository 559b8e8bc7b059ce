"""CPU, world_size 2, gloo: the barcode-sharded two-pass algebra (what each GPU rank computes around
the per-step all-reduce) reproduces the oracle's full gradient.  Covers the N>1 host logic without a GPU."""
import os
import sys

import numpy as np
import pytest
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


# ---------------------------------------------------------------------------------------------------------------
# The PRODUCT's shard logic (csrc/bb_layout.cpp), driven through the C ABI on the CPU: bb_layout_probe runs
# build_layout without any CUDA call.
@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("model,spec", [
    ("fitness_normal", dict(n_neutral=37, n_bc=901, n_time=5)),
    ("replicate_fitness_normal", dict(n_neutral=20, n_bc=333, n_time=5, n_rep=3)),
    ("multienv_fitness_normal", dict(n_neutral=11, n_bc=257, n_time=6, envs=[1, 1, 2, 3, 2, 3])),
    ("genotype_fitness_normal", dict(n_neutral=16, n_bc=600, n_time=5, n_geno=23)),
    ("multienv_replicate_fitness_normal", dict(n_neutral=9, n_bc=130, n_time=6, n_rep=2, envs=[1, 2, 3, 1, 2, 3])),
    ("replicate_fitness_normal", dict(n_neutral=7, n_bc=90, n_time=[5, 4, 6])),
    ("multienv_replicate_fitness_normal", dict(n_neutral=9, n_bc=130, n_time=[6, 4, 6],
                                               envs=[[1, 2, 3, 1, 2, 3], [1, 1, 2, 3], [1, 3, 2, 1, 3, 2]])),
])
def test_product_shard_layout_owns_every_latent_exactly_once(model, spec, world):
    """Over the ranks of a `world`-way split every latent of the reference order is owned by exactly one shard
    (population latents: rank 0), neutral / mutant ranges tile the barcode axis, and -- genotype model -- no genotype
    group is cut (so theta_g needs no exchange)."""
    import barbay_b200 as bb
    da, _ = bb.synth.simulate(model, seed=5, **spec)
    total = None
    prev_n1 = prev_m1 = 0
    g_of = None
    if "genotype" in model:
        _, gidx = bb.model.indexin_unique(list(da.genotypes))
        order = np.argsort(np.asarray(gidx), kind="stable")          # the library's genotype-sorted mutant order
        g_of = np.asarray(gidx)[order]
    hyper_owned = 0
    for rank in range(world):
        pr = bb.Engine(da, model, n_samples=2, rank=rank, world=world, probe_only=True, corrected_ragged=True)
        info = pr.probe_info
        total = pr.owned.astype(np.int64) if total is None else total + pr.owned
        assert info["n0"] == prev_n1 and info["m0"] == prev_m1          # contiguous ranges, rank order
        prev_n1, prev_m1 = info["n1"], info["m1"]
        hyper_owned += info["H"]
        if g_of is not None and 0 < info["m0"] < da.n_bc:
            assert g_of[info["m0"]] != g_of[info["m0"] - 1]            # cut on a genotype boundary
    assert prev_n1 == da.n_neutral and prev_m1 == da.n_bc
    assert total.min() == 1 and total.max() == 1
    if "genotype" in model:
        assert hyper_owned == da.n_geno
    one = bb.Engine(da, model, n_samples=2, probe_only=True)
    assert one.owned.min() == 1 and one.owned.max() == 1 and one.D == total.size


def test_product_layout_probe_validation_messages():
    import barbay_b200 as bb
    rag, _ = bb.synth.simulate("replicate_fitness_normal", n_neutral=4, n_bc=9, n_time=[4, 3], seed=1)
    with pytest.raises(bb.BarBayError, match="single shard"):          # as-written pairing (the default) does not shard
        bb.Engine(rag, "replicate_fitness_normal", rank=0, world=2, probe_only=True)
    bb.Engine(rag, "replicate_fitness_normal", probe_only=True)
    bb.Engine(rag, "replicate_fitness_normal", rank=1, world=2, probe_only=True, corrected_ragged=True)
    # multienv x replicate with unequal T: one environment list per replicate (…replicates.jl:449-472)
    m5 = "multienv_replicate_fitness_normal"
    envs = [["a", "b", "a", "c"], ["a", "c", "b"]]
    rag5, _ = bb.synth.simulate(m5, n_neutral=4, n_bc=9, n_time=[4, 3], envs=envs, seed=1)
    pr = bb.Engine(rag5, m5, {"envs": envs}, probe_only=True)
    assert pr.owned.min() == 1 and pr.owned.max() == 1 and pr.D == 2 * 5 + 3 * 9 + 3 * 3 * 9 * 2 + 7 * 13
    with pytest.raises(bb.BarBayError, match="one environment list per replicate"):
        bb.Engine(rag5, m5, {"envs": ["a", "b", "a", "c"]}, probe_only=True)
    with pytest.raises(bb.BarBayError, match="for all replicates"):
        bb.Engine(rag5, m5, {"envs": [["a", "b", "a", "c"], ["a", "c"]]}, probe_only=True)
    da, _ = bb.synth.simulate("fitness_normal", n_neutral=4, n_bc=9, n_time=4, seed=1)
    with pytest.raises(bb.BarBayError, match="rank, world"):
        bb.Engine(da, "fitness_normal", rank=3, world=2, probe_only=True)
