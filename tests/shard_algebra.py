"""numpy restatement of the sharded two-pass algebra for fitness_normal, used by
the world_size-2 gloo test: pass-1 partial sums on a barcode shard, the five all-reduced sums,
then the per-column and population-latent gradients.  Test infrastructure, mirrors csrc/bb_kernels.cuh."""
import numpy as np


def shard_ranges(N, M, rank, world):
    """Same split as csrc/bb_layout.cpp (non-genotype models)."""
    return (N * rank // world, N * (rank + 1) // world), (M * rank // world, M * (rank + 1) // world)


def unpack(z, T, N, M):
    st, lst = z[:T - 1], z[T - 1:2 * (T - 1)]
    sm, lsm = z[2 * (T - 1):2 * (T - 1) + M], z[2 * (T - 1) + M:2 * (T - 1) + 2 * M]
    ll = z[2 * (T - 1) + 2 * M:].reshape(N + M, T).T          # T x B
    return st, lst, sm, lsm, ll


def pass1_partials(z, T, N, M, rank, world):
    """[Lambda_t (T), Dn_t, D2n_t, A_t, W_t (T-1 each)] over this rank's columns."""
    (n0, n1), (m0, m1) = shard_ranges(N, M, rank, world)
    st, lst, sm, lsm, ll = unpack(z, T, N, M)
    cols = list(range(n0, n1)) + list(range(N + m0, N + m1))
    lam = np.exp(ll[:, cols]).sum(axis=1)
    d_n = np.diff(ll[:, n0:n1], axis=0)
    d_m = np.diff(ll[:, N + m0:N + m1], axis=0)
    w = np.exp(-2.0 * lsm[m0:m1])
    return np.concatenate([lam, d_n.sum(axis=1), (d_n ** 2).sum(axis=1),
                           (w * (d_m - sm[m0:m1])).sum(axis=1), np.full(T - 1, w.sum())])


def gradients_from_sums(z, sums, R, T, N, M, rank, world, pri):
    """d log pi / dz restricted to the latents this rank owns (zeros elsewhere; rank 0 owns the
    population latents).  `sums` are the all-reduced pass-1 sums; pri = dict of (mean, std)."""
    (n0, n1), (m0, m1) = shard_ranges(N, M, rank, world)
    st, lst, sm, lsm, ll = unpack(z, T, N, M)
    lam_t, dn, d2n, am, wm = sums[:T], sums[T:2 * T - 1], sums[2 * T - 1:3 * T - 2], sums[3 * T - 2:4 * T - 3], sums[4 * T - 3:]
    c = np.diff(np.log(lam_t))
    wbar = np.exp(-2.0 * lst)
    a = st - c
    U = wbar * (dn + N * a) + (am + a * wm)
    G = (np.concatenate([[0.0], U]) - np.concatenate([U, [0.0]])) / lam_t
    g = np.zeros_like(z)
    if rank == 0:
        g[:T - 1] = -U - (st - pri["s_pop"][0]) / pri["s_pop"][1] ** 2
        g[T - 1:2 * (T - 1)] = wbar * (d2n + 2 * a * dn + N * a * a) - N - (lst - pri["lsig_pop"][0]) / pri["lsig_pop"][1] ** 2
    off_s, off_ls, off_ll = 2 * (T - 1), 2 * (T - 1) + M, 2 * (T - 1) + 2 * M
    for b in list(range(n0, n1)) + list(range(N + m0, N + m1)):
        zb = ll[:, b]
        lam = np.exp(zb)
        gb = (R[:, b] - lam) + lam * G - (zb - pri["lam"][0]) / pri["lam"][1] ** 2
        d = np.diff(zb)
        if b < N:
            u = wbar * (d - c + st)
        else:
            m = b - N
            res = d - c - (sm[m] - st)
            wv = np.exp(-2.0 * lsm[m])
            u = wv * res
            g[off_s + m] = u.sum() - (sm[m] - pri["s_bc"][0]) / pri["s_bc"][1] ** 2
            g[off_ls + m] = (u * res).sum() - (T - 1) - (lsm[m] - pri["lsig_bc"][0]) / pri["lsig_bc"][1] ** 2
        gb[:-1] += u
        gb[1:] -= u
        g[off_ll + b * T: off_ll + (b + 1) * T] = gb
    return g
