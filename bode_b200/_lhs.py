"""Latin-hypercube designs with pyDOE's call signature and RNG consumption (host stand-in; pyDOE is not installed).

``lhs(n, samples, criterion)`` draws from the global ``numpy.random`` state exactly as pyDOE does:
  * ``criterion=None`` (classic): one ``rand(samples, n)`` block, then one ``permutation(samples)`` per column;
  * ``criterion='center'`` / ``'c'``: one ``permutation`` of the cell centres per column.
The reference uses both (``_sampler.py:613, 631, 638``; drivers ``samp_ex1.py:51, 61``).
"""
import numpy as np


def lhs(n, samples=None, criterion=None):
    if samples is None:
        samples = n
    cut = np.linspace(0, 1, samples + 1)
    a, b = cut[:samples], cut[1:samples + 1]
    H = np.zeros((samples, n))
    if criterion is None or str(criterion).lower() in ("classic", "none"):
        u = np.random.rand(samples, n)
        rd = np.zeros_like(u)
        for j in range(n):
            rd[:, j] = u[:, j] * (b - a) + a
        for j in range(n):
            H[:, j] = rd[np.random.permutation(range(samples)), j]
        return H
    if str(criterion).lower() in ("center", "c"):
        centres = (a + b) / 2
        for j in range(n):
            H[:, j] = np.random.permutation(centres)
        return H
    raise ValueError("lhs stand-in supports criterion=None or 'center'")
