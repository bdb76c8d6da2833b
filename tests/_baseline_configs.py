"""Synthetic inputs of BASELINE.json configs[0] (C1) and configs[1] (C2), exactly as SURVEY 8(d) fixes them
(``numpy.random.Generator(PCG64(20240 + config number))``, interactions deduplicated and row-major sorted like
``utils.py:53-57`` produces).  Shared by the CPU oracle tests and the GPU parity tests -- test infrastructure only."""
import numpy as np

from oracle import mf_oracle as o


def _zipf_cells(rng, n_u, n_i, nnz):
    wu = 1.0 / np.arange(1, n_u + 1)
    wi = 1.0 / np.arange(1, n_i + 1)
    cells = np.zeros(0, dtype=np.int64)
    while cells.size < nnz:
        m = int((nnz - cells.size) * 1.5) + 1024
        u = rng.choice(n_u, size=m, p=wu / wu.sum())
        i = rng.choice(n_i, size=m, p=wi / wi.sum())
        cells = np.unique(np.concatenate([cells, u.astype(np.int64) * n_i + i]))
    if cells.size > nnz:
        cells = np.sort(rng.permutation(cells)[:nnz])
    return cells // n_i, cells % n_i


def c1():
    """toy 1k x 1k, density 0.01 (~10,000 interactions), value 1.0, rank 10, MSE, identity features, Normal init,
    lr 1e-2 (examples/benchmark_toydata.py:25-53 with BASELINE.json's shape)."""
    rng = np.random.Generator(np.random.PCG64(20241))
    n_u = n_i = 1000
    cells = np.sort(rng.choice(n_u * n_i, size=10_000, replace=False))
    rows, cols = cells // n_i, cells % n_i
    vals = np.ones(rows.size, np.float32)
    return dict(name="c1", n_u=n_u, n_i=n_i, r=10, S=0, loss="mse", lr=1e-2, rows=rows, cols=cols, vals=vals, samp=None,
                U0=o.normal_initializer(n_u, 10, rng), V0=o.normal_initializer(n_i, 10, rng))


def c2():
    """ML-100K-shaped: 943 x 1,682, 100,000 Zipf(1.0)-weighted interactions, ratings 1-5 with P = (.06,.11,.27,.34,.22), training
    on ratings >= 4 with their values (examples/benchmarking_ML.py:38,61,102), rank 32, WMRB, S = n_items // 5 = 336 (:65),
    Uniform init, lr 0.1 (:67,75)."""
    rng = np.random.Generator(np.random.PCG64(20242))
    n_u, n_i, r = 943, 1682, 32
    S = n_i // 5
    rows, cols = _zipf_cells(rng, n_u, n_i, 100_000)
    rating = rng.choice(np.arange(1, 6), size=rows.size, p=[.06, .11, .27, .34, .22]).astype(np.float32)
    keep = rating >= 4
    samp = np.stack([rng.choice(n_i, S, replace=False) for _ in range(n_u)]).astype(np.int64)
    return dict(name="c2", n_u=n_u, n_i=n_i, r=r, S=S, loss="wmrb", lr=0.1, rows=rows[keep], cols=cols[keep], vals=rating[keep],
                samp=samp, all_rows=rows, all_cols=cols, all_ratings=rating,
                U0=o.uniform_initializer(n_u, r, rng), V0=o.uniform_initializer(n_i, r, rng))
