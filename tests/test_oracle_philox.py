"""CPU: Philox4x32-10 known-answer vectors (Random123 kat_vectors) and lattice sanity."""
import numpy as np

from helpers import load_fixture, oracle_problem
from oracle import philox_ref

KAT = [  # (counter, key, expected) -- Random123 tests/kat_vectors, philox4x32 10 rounds
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox4x32_10_known_answers():
    for ctr, key, exp in KAT:
        out = philox_ref.philox4x32_10(*[np.uint32(c) for c in ctr], key[0], key[1])
        assert tuple(int(x) for x in out) == exp


def test_uniforms_open_interval_and_box_muller_moments():
    x = np.array([0, 0xFFFFFFFF, 0x80000000, 0x7FF, 0xFFFFF800], dtype=np.uint32)
    u, v = philox_ref.word_uniforms(x)
    assert u.min() > 0 and u.max() < 1 and v.min() > 0 and v.max() < 1
    # 2048 equispaced angles: trigonometric moments are exact (cos^2 -> 1/2, cos^4 -> 3/8, cos*sin -> 0)
    ang = 2 * np.pi * (np.arange(2048) + 0.5) / 2048
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
