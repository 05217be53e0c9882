"""Seeded problem generators shared by the CPU and GPU tests (pure NumPy; no reference tree needed)."""
import numpy as np


def ex1_func(x):
    """Objective of the reference driver tests/samp_ex1.py:28-35 with sigma = 0."""
    x = 6 * x
    return (4 * (1. - np.sin(x[0] + 8 * np.exp(x[0] - 7.))) - 10.) / 5.


def ex1_problem():
    """samp_ex1 at iteration 0: X = centred LHS of 3 points in 1-D (order irrelevant), Y = Ex1Func(X)."""
    X = np.array([[1 / 6], [1 / 2], [5 / 6]])
    Y = np.array([ex1_func(x) for x in X])[:, None]
    return X, Y


def dette8(x):
    """8-D Dette & Pepelyshev (2010) curved function, scaled like reference tests/dette.py:93-94."""
    x = np.atleast_2d(x)
    y = 4 * (x[:, 0] - 2 + 8 * x[:, 1] - 8 * x[:, 1] ** 2) ** 2 + (3 - 4 * x[:, 1]) ** 2 \
        + 16 * np.sqrt(x[:, 2] + 1) * (2 * x[:, 2] - 1) ** 2
    for i in range(4, 9):
        y = y + i * np.log(1 + np.sum(x[:, 2:i], axis=1))
    return (y - 50.) / 50.


def ex4_func6(x):
    """6-D test function of reference tests/samp_ex4.py:28-39 (`Ex1Func`, noise-free): a scaled DTLZ-type function from
    Knowles (2006): g = 100 (sum_{i=1..5} ((x_i - 1/2)^2 - cos(2 pi (x_i - 1/2))) + 5), k = (1 - x_0) (g + 1) / 2, (k - 150) / 80."""
    x = np.asarray(x, dtype=float)
    g = 100. * (((x[1:6] - 0.5) ** 2 - np.cos(2. * np.pi * (x[1:6] - 0.5))).sum() + 5.)
    k = 0.5 * (1 - x[0]) * (g + 1.)
    return float((k - 150.) / 80.)


def make_problem(d, n, S, Nd, seed, ell_lo=0.15, ell_hi=0.9, var_shape=2.0, noisy=False, objective=None):
    """Well-conditioned synthetic problem: uniform lengthscales in [ell_lo, ell_hi], Gamma variance."""
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    if objective is None:
        Y = np.sin(3 * X.sum(1))[:, None] + 0.3 * X[:, :1]
    else:
        Y = np.asarray([objective(x) for x in X], dtype=float).reshape(n, 1)
    Xd = rng.random((Nd, d))
    cols = [rng.gamma(var_shape, 1.0, S) + 0.2] + [rng.uniform(ell_lo, ell_hi, S) for _ in range(d)]
    if noisy:
        cols.append(rng.uniform(0.01, 0.1, S))
    hyp = np.column_stack(cols)
    return X, Y, Xd, hyp


def prior_problem(d, n, S, Nd, seed, ell_min=0.0):
    """Prior-drawn hyper-samples as in the reference drivers (variance ~ Gamma(1, scale 2), ell ~ Exp(scale 0.7),
    samp_ex2.py:107-108): long lengthscales make many samples ill-conditioned (cond(A) up to ~1e8)."""
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    Y = np.sin(X.sum(1))[:, None]
    Xd = rng.random((Nd, d))
    hyp = np.column_stack([rng.gamma(1.0, 2.0, S) + 1e-3] + [np.maximum(rng.exponential(0.7, S), ell_min) for _ in range(d)])
    return X, Y, Xd, hyp
