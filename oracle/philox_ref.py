"""Oracle for the counter-based noise lattice (Philox4x32-7 + Box-Muller).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference draws eps from Julia's global Xoshiro RNG inside AdvancedVI
(``rand_and_logjac``; call site src/vi.jl:201) and is not reproducible draw
for draw.  The CUDA backend instead defines a stateless
noise lattice; this file restates that definition independently, from the
specification in DESIGN.md §"Noise lattice", so tests can compare the kernel's
draws and whole ADVI trajectories.

Philox4x32-R is the published algorithm of Salmon et al., "Parallel random
numbers: as easy as 1, 2, 3" (SC'11); the lattice uses R = 7 rounds (the smallest
Crush-resistant count reported there; 10 is the library's safety-margin default).
The known-answer vectors checked in tests/test_oracle_philox.py are the Random123
``kat_vectors`` for philox4x32-10: they pin the round function and key schedule of
this implementation, of which the 7-round lattice is the same loop run 7 times.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

STREAM_COLUMN = 0
STREAM_SHARED = 1
STREAM_HYPER = 2
STREAM_INIT = 3


LATTICE_ROUNDS = 7      # the noise lattice uses Philox4x32-7 (DESIGN.md "Noise lattice"); the KATs pin the 10-round function


def philox4x32(c0, c1, c2, c3, k0: int, k1: int, rounds: int = 10):
    """Vectorised Philox4x32-R.  Counters are uint32 arrays (broadcastable)."""
    c0, c1, c2, c3 = np.broadcast_arrays(
        np.asarray(c0, dtype=np.uint64), np.asarray(c1, dtype=np.uint64),
        np.asarray(c2, dtype=np.uint64), np.asarray(c3, dtype=np.uint64))
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0)
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def word_uniforms(x: np.ndarray):
    """One Philox word -> (radius uniform from the top 22 bits, angle uniform from the low 10 bits)."""
    u = ((x >> np.uint32(10)).astype(np.float64) + 0.5) / 4194304.0
    v = ((x & np.uint32(0x3FF)).astype(np.float64) + 0.5) / 1024.0
    return u, v


def box_muller(x: np.ndarray):
    """word -> (radius*cos(angle), radius*sin(angle))."""
    u, v = word_uniforms(x)
    radius = np.sqrt(-2.0 * np.log(u))
    angle = 2.0 * np.pi * v
    return radius * np.cos(angle), radius * np.sin(angle)


def normals8(c0, c1, c2, c3, seed: int):
    """Eight standard normals per counter: word w -> lanes 2w, 2w + 1."""
    xs = philox4x32(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF, LATTICE_ROUNDS)
    out = []
    for x in xs:
        a, b = box_muller(x)
        out += [a, b]
    return np.stack(out, axis=-1)


def lattice_normal(entity, slot, stream: int, k: int, step: int, seed: int) -> np.ndarray:
    """Normal attached to (stream, entity, slot) for MC sample k of ADVI step ``step``.

    counter = (entity_word, stream << 24 | call, k, step) with
      column stream : entity_word = column id, call = slot >> 3, lane = slot & 7
      other streams : entity_word = slot >> 3 (entity unused), call = 0, lane = slot & 7
    """
    entity = np.asarray(entity, dtype=np.uint32)
    slot = np.asarray(slot, dtype=np.uint32)
    if stream == STREAM_COLUMN:
        c0 = entity
        call = slot >> np.uint32(3)
    else:
        c0 = slot >> np.uint32(3)
        call = np.zeros_like(slot)
    c1 = (np.uint32(stream) << np.uint32(24)) | call
    lane = (slot & np.uint32(7)).astype(np.int64)
    n8 = normals8(c0, c1, np.uint32(k), np.uint32(step), seed)
    return np.take_along_axis(n8, lane[..., None], axis=-1)[..., 0]


# --------------------------------------------------------------------------
# Lattice coordinates of every latent, in reference (VarInfo) order
# --------------------------------------------------------------------------
def lattice_coords(model: str, prob: dict) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Returns (stream, entity, slot) arrays of length D in reference latent order.

    Column ids: replicate r, barcode b (neutrals first, 0-based) -> r * B + b.
    Column slots: log-lambda at time t -> t; then per environment e either
    (s, log-sigma) -> T_r + 2e + {0,1} or (theta-tilde, log-tau, log-sigma) ->
    T_r + 3e + {0,1,2} for hierarchical models.
    Shared (population) latents use their index in [s-bar..., log-sigma-bar...];
    hyper latents theta use their index in the reference theta vector.
    """
    from .model_ref import _indexin_unique

    R = prob["bc_count"]
    N, M = prob["n_neutral"], prob["n_bc"]
    B = N + M
    if isinstance(R, (list, tuple)):
        Ts = [int(np.asarray(r).shape[0]) for r in R]
    else:
        Ra = np.asarray(R)
        Ts = [Ra.shape[0]] * (Ra.shape[2] if Ra.ndim == 3 else 1)
    n_rep = len(Ts)
    hier = model in ("replicate_fitness_normal", "genotype_fitness_normal",
                     "multienv_replicate_fitness_normal")
    E = len(_indexin_unique(list(prob["envs"]))[0]) if "multienv" in model else 1
    n_st = sum(t - 1 for t in Ts)
    stream, entity, slot = [], [], []

    def add(st, en, sl):
        stream.append(np.full(len(sl), st, dtype=np.int64))
        entity.append(np.asarray(en, dtype=np.int64))
        slot.append(np.asarray(sl, dtype=np.int64))

    add(STREAM_SHARED, np.zeros(n_st), np.arange(n_st))                 # s-bar
    add(STREAM_SHARED, np.zeros(n_st), n_st + np.arange(n_st))          # log-sigma-bar
    per = 3 if hier else 2
    if hier:
        n_hyper = (len(_indexin_unique(list(prob["genotypes"]))[0])
                   if model == "genotype_fitness_normal" else E * M)
        add(STREAM_HYPER, np.zeros(n_hyper), np.arange(n_hyper))        # theta
    # barcode-level groups, reference index = e + E*(m + M*r)
    e_i, m_i, r_i = np.meshgrid(np.arange(E), np.arange(M), np.arange(n_rep), indexing="ij")
    e_f, m_f, r_f = (a.transpose(2, 1, 0).reshape(-1) for a in (e_i, m_i, r_i))
    T_of = np.asarray(Ts)[r_f]
    col = r_f * B + N + m_f
    for j in range(per):                                               # (s,lσ) or (θ̃,lτ,lσ)
        add(STREAM_COLUMN, col, T_of + per * e_f + j)
    for r in range(n_rep):                                             # log-lambda, (r, b, t) t fastest
        b_i, t_i = np.meshgrid(np.arange(B), np.arange(Ts[r]), indexing="ij")
        add(STREAM_COLUMN, (r * B + b_i).reshape(-1), t_i.reshape(-1))
    return np.concatenate(stream), np.concatenate(entity), np.concatenate(slot)


def noise(model: str, prob: dict, n_samples: int, step: int, seed: int) -> np.ndarray:
    """eps[k, i] for k < n_samples and every latent i in reference order."""
    stream, entity, slot = lattice_coords(model, prob)
    out = np.empty((n_samples, stream.size), dtype=np.float64)
    for st in (STREAM_COLUMN, STREAM_SHARED, STREAM_HYPER):
        sel = np.nonzero(stream == st)[0]
        if sel.size == 0:
            continue
        for k in range(n_samples):
            out[k, sel] = lattice_normal(entity[sel], slot[sel], st, k, step, seed)
    return out


def init_params(D: int, seed: int) -> tuple[np.ndarray, np.ndarray]:
    """Mean-field initialisation mu ~ N(0,1), omega ~ N(0,1) (Turing ``meanfield``:
    mu = randn(D), sigma = softplus.(randn(D)), theta = vcat(mu, invsoftplus.(sigma)))."""
    j = np.arange(2 * D, dtype=np.int64)
    v = lattice_normal(np.zeros_like(j), j, STREAM_INIT, 0, 0, seed)
    return v[:D].copy(), v[D:].copy()
