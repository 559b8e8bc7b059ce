"""CPU: Philox4x32 known-answer vectors (Random123 kat_vectors, 10 rounds) and the statistics of the 7-round noise
lattice the kernels use (the CUDA draws are checked against this oracle in tests/test_gpu_parity.py)."""
import numpy as np

from helpers import load_fixture, oracle_problem
from oracle import philox_ref

KAT = [  # (counter, key, expected) -- Random123 tests/kat_vectors, philox4x32 10 rounds
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox4x32_known_answers():
    for ctr, key, exp in KAT:
        out = philox_ref.philox4x32(*[np.uint32(c) for c in ctr], key[0], key[1], rounds=10)
        assert tuple(int(x) for x in out) == exp
    # the lattice runs the same round function / key schedule 7 times (Random123's smallest Crush-resistant count)
    assert philox_ref.LATTICE_ROUNDS == 7
    a = philox_ref.philox4x32(np.uint32(1), np.uint32(2), np.uint32(3), np.uint32(4), 5, 6, rounds=7)
    b = philox_ref.philox4x32(np.uint32(1), np.uint32(2), np.uint32(3), np.uint32(4), 5, 6, rounds=10)
    assert tuple(map(int, a)) != tuple(map(int, b))


def test_lattice_distribution_ks_tails_and_independence():
    """10^7 draws of the 7-round lattice: Kolmogorov-Smirnov distance to N(0, 1), tail masses (the radius uniform has
    22 bits, so |n| <= sqrt(2 ln 2^23) = 5.65: documented truncation), and no correlation between the lanes of a
    counter, between consecutive columns, samples or steps."""
    from scipy import special, stats
    ncol = 1_250_000
    col = np.arange(ncol, dtype=np.uint32)
    n8 = philox_ref.normals8(col, np.uint32(philox_ref.STREAM_COLUMN << 24), np.uint32(3), np.uint32(17), 0x1234_5678_9ABC)
    x = n8.reshape(-1)
    n = x.size
    d = stats.kstest(x, "norm").statistic
    assert d < 1.63 / np.sqrt(n), d                      # 1 % critical value of the KS distance
    for thr in (2.0, 3.0, 4.0):
        p = special.erfc(thr / np.sqrt(2.0))
        got = np.mean(np.abs(x) > thr)
        assert abs(got - p) < 5.0 * np.sqrt(p * (1 - p) / n), (thr, got, p)
    assert np.abs(x).max() <= np.sqrt(2.0 * np.log(2.0 ** 23)) + 1e-12
    assert abs(x.mean()) < 5 / np.sqrt(n) and abs(x.var() - 1) < 5 * np.sqrt(2.0 / n) and abs((x ** 4).mean() - 3) < 0.02
    c = np.corrcoef(n8.T)                                # the eight lanes of one counter
    assert np.abs(c - np.eye(8)).max() < 5 / np.sqrt(ncol)
    assert abs(np.corrcoef(n8[:-1, 0], n8[1:, 0])[0, 1]) < 5 / np.sqrt(ncol)          # neighbouring columns
    other = philox_ref.normals8(col, np.uint32(philox_ref.STREAM_COLUMN << 24), np.uint32(4), np.uint32(17), 0x1234_5678_9ABC)
    later = philox_ref.normals8(col, np.uint32(philox_ref.STREAM_COLUMN << 24), np.uint32(3), np.uint32(18), 0x1234_5678_9ABC)
    assert abs(np.corrcoef(n8[:, 0], other[:, 0])[0, 1]) < 5 / np.sqrt(ncol)          # next MC sample
    assert abs(np.corrcoef(n8[:, 0], later[:, 0])[0, 1]) < 5 / np.sqrt(ncol)          # next step
    assert abs(np.corrcoef(n8[:, 0] ** 2, other[:, 0] ** 2)[0, 1]) < 5 / np.sqrt(ncol)


def test_uniforms_open_interval_and_box_muller_moments():
    x = np.array([0, 0xFFFFFFFF, 0x80000000, 0x3FF, 0xFFFFFC00], dtype=np.uint32)
    u, v = philox_ref.word_uniforms(x)
    assert u.min() > 0 and u.max() < 1 and v.min() > 0 and v.max() < 1
    # 1024 equispaced angles: trigonometric moments are exact (cos^2 -> 1/2, cos^4 -> 3/8, cos*sin -> 0)
    ang = 2 * np.pi * (np.arange(1024) + 0.5) / 1024
    assert abs(np.mean(np.cos(ang) ** 2) - 0.5) < 1e-14 and abs(np.mean(np.cos(ang) ** 4) - 0.375) < 1e-14
    assert abs(np.mean(np.cos(ang) * np.sin(ang))) < 1e-14
    idx = np.arange(200_000, dtype=np.uint32)
    n = philox_ref.lattice_normal(idx, np.zeros_like(idx), philox_ref.STREAM_COLUMN, 0, 0, 12345)
    assert abs(n.mean()) < 0.01 and abs(n.std() - 1) < 0.01
    assert abs(((n ** 4).mean()) - 3) < 0.1


def test_lattice_covers_every_latent_once(bb):
    for model in ("fitness_normal", "replicate_fitness_normal", "multienv_fitness_normal",
                  "genotype_fitness_normal"):
        df, cols = load_fixture(model)
        da = bb.utils.data_to_arrays(df, **cols)
        prob = oracle_problem(da, model)
        stream, entity, slot = philox_ref.lattice_coords(model, prob)
        keys = set(zip(stream.tolist(), entity.tolist(), slot.tolist()))
        assert len(keys) == stream.size            # no two latents share a draw
        eps = philox_ref.noise(model, prob, 2, 3, 7)
        assert eps.shape == (2, stream.size) and np.isfinite(eps).all()
        assert not np.allclose(eps[0], eps[1])
        assert not np.allclose(eps, philox_ref.noise(model, prob, 2, 4, 7))   # fresh draws every step
