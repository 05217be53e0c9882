"""CPU: pin the oracle to vectors produced by the reference's OWN code.

``tests/golden/ekld_ref_*.npz`` hold what the reference's unmodified ``bode/_sampler.py:391-492`` (``get_sigma_1``,
``get_mu_1``, ``get_params_2``, ``avg_kld_mean``, ``make_mcmc_model``, ``mcmc_kld``) returned on seeded inputs when run
over the GPy stand-in of ``tests/stubs`` (generator: ``tests/golden/make_golden.py``).  Two things are checked here:

* the oracle's literal restatement (``oracle.ekld_oracle.avg_kld_mean`` & co.) reproduces those vectors to 1e-13 --
  sigma_x, v^2, sigma_2, the EKLD value, the mean / variance of log EKLD and the argmax are no longer "unpinned";
* the vectorised oracle the GPU tests compare against (``orc.score``: same arithmetic, BLAS calls batched over the
  candidates) agrees with the reference's per-candidate evaluation under the tolerance model of tests/tolerance.py --
  batching changes the rounding of k(X, x) by an ulp, which cond(A) amplifies exactly like any other 1e-16
  perturbation of the reference route.
"""
import os

import numpy as np
import pytest

from oracle import ekld_oracle as orc
from oracle import truth
from tests.tolerance import assert_argmax, assert_parity

TAGS = ("c1", "c2", "c3s", "noisy")


def load(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "ekld_ref_%s.npz" % tag))  # allow_pickle=False: plain arrays only
    return {k: g[k] for k in g.files}


def _rtol(model):
    """1e-13, widened only if the host BLAS rounds differently from the generating machine: then the difference is
    eps-sized in k(X, x) and amplified by cond(A) (the reference route uses the explicit inverse)."""
    A = model.K(model.X) + (model.noise_var + orc.GPY_JITTER) * np.eye(model.X.shape[0])
    return 1e-13, 64 * np.finfo(float).eps * np.linalg.cond(A)


@pytest.mark.parametrize("tag", TAGS)
def test_literal_oracle_reproduces_reference_code(golden_dir, tag):
    g = load(golden_dir, tag)
    X, Y, hyp, Xd, nugget = g["X"], g["Y"], g["hyp"], g["Xd"], float(g["nugget"])
    S, Nd = g["kld"].shape
    assert hyp.shape[1] == X.shape[1] + (2 if int(g["noisy"]) else 1)
    cols = range(0, Nd, max(1, Nd // 48))
    exact = 0
    for s in range(S):
        m = orc.make_mcmc_model(hyp[s], X, Y, nugget)
        tight, loose = _rtol(m)
        for got, ref in ((orc.get_sigma_1(m), g["sigma_1"][s]), (orc.get_mu_1(m), g["mu_1"][s])):
            assert abs(got - ref) <= tight * abs(ref) or abs(got - ref) <= loose * abs(ref)
        for c in cols:
            v2, s2 = orc.get_params_2(m, Xd[c])
            sx = m.predict_var_full_cov(np.atleast_2d(Xd[c]))[0][0]
            with np.errstate(invalid="ignore", divide="ignore"):
                k = float(np.asarray(orc.avg_kld_mean(m, Xd[c])).item())
            for got, ref, floor in ((v2, g["v2"][s, c], 0.0), (s2, g["sigma_2"][s, c], 0.0), (sx, g["sigma_x"][s, c], 0.0),
                                    (k, g["kld"][s, c], 1e-15)):
                err = abs(got - ref)
                assert err <= tight * abs(ref) + floor or err <= loose * max(abs(ref), m.variance), (tag, s, c, got, ref)
                exact += got == ref
    assert exact > 0
    # mcmc_kld itself (:483-492): np.mean / np.var of the log over the hyper-samples, per candidate
    for c in list(cols)[:12]:
        mean, var = orc.mcmc_kld_literal(Xd[c], X, Y, hyp, nugget)
        assert abs(mean - g["mcmc_mean"][c]) <= 1e-9 and abs(var - g["mcmc_var"][c]) <= 1e-9 * max(1.0, g["mcmc_var"][c])


@pytest.mark.parametrize("tag,min_trusted", [("c1", 0.0), ("c2", 0.0), ("c3s", 1.0), ("noisy", 1.0)])
def test_vectorised_oracle_agrees_with_reference_code(golden_dir, tag, min_trusted):
    g = load(golden_dir, tag)
    X, Y, hyp, Xd, nugget = g["X"], g["Y"], g["hyp"], g["Xd"], float(g["nugget"])
    r = orc.score(X, Y, Xd, hyp, nugget=nugget)
    t = truth.ekld_truth(X, Y, Xd, hyp, nugget=nugget)
    trusted = assert_parity(r["kld"], g["kld"], t["kld"], min_trusted=min_trusted, what="vectorised oracle vs reference code")
    np.testing.assert_allclose(r["sigma_1"], g["sigma_1"], rtol=1e-13)
    np.testing.assert_allclose(r["mu_1"], g["mu_1"], rtol=1e-13)
    # argmax: the reference's own (np.argmax of its mcmc_kld means); ties up to rounding resolved through the truth
    assert_argmax(r["argmax"], g["mcmc_mean"], tol=1e-6)
    assert_argmax(int(np.argmax(g["mcmc_mean"])), np.log(t["kld"]).mean(0), tol=1e-6)
    if trusted.all():
        # sequential vs pairwise summation over s and the batched BLAS rounding: far below the 1e-9 bar
        np.testing.assert_allclose(r["score"], g["mcmc_mean"], rtol=0, atol=1e-9)
        assert r["argmax"] == int(np.argmax(g["mcmc_mean"]))


def test_reference_vectors_cover_the_named_shapes(golden_dir):
    shapes = {tag: load(golden_dir, tag) for tag in TAGS}
    assert shapes["c1"]["kld"].shape == (60, 1000) and shapes["c1"]["X"].shape == (3, 1)       # config 1, fixture hyper-samples
    assert shapes["c2"]["kld"].shape == (32, 400) and shapes["c2"]["X"].shape == (30, 2)        # config-2 shape
    assert shapes["c3s"]["kld"].shape == (8, 96) and shapes["c3s"]["X"].shape == (200, 8)       # config-3 shape, reduced S / N_d
    assert shapes["noisy"]["hyp"].shape[1] == shapes["noisy"]["X"].shape[1] + 2
    comp = np.load(os.path.join(golden_dir, "comp_ex1.npz"))
    np.testing.assert_array_equal(shapes["c1"]["hyp"], comp["models_it0"])                      # the reference's own hyper-samples
    for g in shapes.values():
        assert np.isfinite(g["kld"]).all() and (g["kld"] > 0).all()
