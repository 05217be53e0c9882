"""Generate the committed golden fixtures from the reference tree (run in the build container only).

    python tests/golden/make_golden.py

Inputs (read-only, never present on the GPU box):
  /root/reference/tests/mcmc_ex1_n=3_it=1/comp.pkl   -- the reference's only shipped result fixture
        (pickle of ``comp_log = (mu_us, sigma_us, models, models_us)``, written by tests/samp_ex1.py:116-117)
  /root/reference/bode/_core.py                       -- imported by file path (needs only numpy/scipy) to
        generate known-answer vectors for ``xik`` / ``bk`` / ``ei`` / the prior log-pdfs.
  /root/reference/bode/_sampler.py                    -- imported UNMODIFIED over the stub modules of tests/stubs
        (GPy / pyDOE / emcee / seaborn / matplotlib are not installable here); its own methods get_sigma_1, get_mu_1,
        get_params_2, avg_kld_mean, make_mcmc_model and mcmc_kld (:391-492) are executed verbatim on seeded inputs.
Outputs (committed):
  tests/golden/comp_ex1.npz      hyper-samples + Q summaries of samp_ex1 at iteration 0 / 1
  tests/golden/core_xik_bk.npz   inputs and reference outputs of xik, bk, ei, priors
  tests/golden/ekld_ref_*.npz    inputs and the reference code's own sigma_1, mu_1, sigma_x, v^2, sigma_2, EKLD,
                                 mean / var of log EKLD for config 1 (fixture hyper-samples, 1000 designs), a config-2
                                 shape, a small config-3 shape and a noisy (d+2 layout) case
The pickle and the reference sources are untrusted public content: the pickle is read through a restricted
unpickler (NumPy array reconstruction and builtin containers only), and this script is meant for the sandboxed build
container only.  The committed .npz files hold plain arrays (np.load with allow_pickle=False).
"""
import importlib.util
import io
import os
import pickle
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class _ArrayOnlyUnpickler(pickle.Unpickler):
    """Only what a pickled tuple / list of float64 ndarrays needs."""
    _ALLOWED = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
                ("numpy", "ndarray"), ("numpy", "dtype"),
                ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar")}

    def find_class(self, module, name):
        if (module, name) in self._ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError("refusing to unpickle %s.%s" % (module, name))


def reference_ekld_vectors():
    """Run the reference's own EKLD code (``_sampler.py:391-492``) on seeded inputs -> dict of npz payloads."""
    from tests.stubs import install
    ref = install.import_reference()
    import GPy  # the stub
    from tests import cases

    def run(X, Y, hyp, Xd, nugget, noisy):
        kls = object.__new__(ref.KLSampler)  # no constructor: it would start the host MCMC
        kls.X, kls.Y, kls.noisy, kls.nugget, kls.model_kern, kls.mcmc_model = X, Y, noisy, nugget, GPy.kern.RBF, True
        kls.model = [hyp]
        S, Nd = hyp.shape[0], Xd.shape[0]
        out = dict(X=X, Y=Y, hyp=hyp, Xd=Xd, nugget=np.array(nugget), noisy=np.array(int(noisy)))
        sigma_1 = np.empty(S); mu_1 = np.empty(S)
        kld = np.empty((S, Nd)); sigma_x = np.empty((S, Nd)); v2 = np.empty((S, Nd)); sigma_2 = np.empty((S, Nd))
        for s in range(S):
            m = kls.make_mcmc_model(hyp[s, :], X, Y)                      # :449-462
            sigma_1[s] = kls.get_sigma_1(m, X, Y)                          # :404-410
            mu_1[s] = kls.get_mu_1(m, X, Y)                                # :413-422
            for c in range(Nd):
                x = Xd[c]
                sigma_x[s, c] = m.predict(np.atleast_2d(x), include_likelihood=False, full_cov=True)[1][0][0]  # :475
                v2[s, c], sigma_2[s, c] = kls.get_params_2(m, X, Y, x)     # :391-402
                with np.errstate(invalid="ignore", divide="ignore"):
                    kld[s, c] = np.asarray(kls.avg_kld_mean(x, X, Y, model=m)).item()  # :464-481
        mean = np.empty(Nd); var = np.empty(Nd)
        with np.errstate(invalid="ignore", divide="ignore"):
            for c in range(Nd):
                mean[c], var[c] = kls.mcmc_kld(Xd[c], kls.model)           # :483-492 (re-builds every model per candidate)
        out.update(sigma_1=sigma_1, mu_1=mu_1, kld=kld, sigma_x=sigma_x, v2=v2, sigma_2=sigma_2, mcmc_mean=mean, mcmc_var=var)
        return out

    res = {}
    # config 1: samp_ex1 at iteration 0 with the fixture's own hyper-samples, |X_design| = 1000 centred LHS
    comp = np.load(os.path.join(HERE, "comp_ex1.npz"))
    X, Y = cases.ex1_problem()
    rng = np.random.default_rng(1333)
    Xd = rng.permutation((np.arange(1000) + 0.5) / 1000.0)[:, None]
    res["c1"] = run(X, Y, comp["models_it0"], Xd, 1e-3, False)
    # config-2 shape: d=2, n=30, S=32, prior-drawn hyper-samples (samp_ex2.py:107-108) with ell >= 0.05
    X, Y, Xd, hyp = cases.prior_problem(2, 30, 32, 400, seed=1263, ell_min=0.05)
    res["c2"] = run(X, Y, hyp, Xd, 1e-3, False)
    # small config-3 shape: d=8, n=200, Dette-Pepelyshev observations, 8 hyper-samples x 96 candidates
    rng = np.random.default_rng(1223)
    X = rng.random((200, 8)); Y = cases.dette8(X)[:, None]
    hyp = np.column_stack([rng.gamma(1.0, 2.0, 8) + 0.05] + [np.maximum(rng.exponential(0.7, 8), 0.05) for _ in range(8)])
    Xd = rng.random((96, 8))
    res["c3s"] = run(X, Y, hyp, Xd, 1e-3, False)
    # noisy layout (d+2 columns, noise std in the last one): make_mcmc_model's other branch (:454-457)
    X, Y, Xd, hyp = cases.make_problem(3, 20, 6, 64, seed=77, noisy=True)
    res["noisy"] = run(X, Y, hyp, Xd, 1e-3, True)
    return res


def main():
    with open(os.path.join(REF, "tests", "mcmc_ex1_n=3_it=1", "comp.pkl"), "rb") as f:
        mu_us, sigma_us, models, models_us = _ArrayOnlyUnpickler(io.BytesIO(f.read())).load()
    np.savez(os.path.join(HERE, "comp_ex1.npz"),
             mu_us=np.asarray(mu_us, dtype=np.float64), sigma_us=np.asarray(sigma_us, dtype=np.float64),
             models_it0=models[0][0], models_it1=models[1][0],
             models_us_it0=models_us[0][0], models_us_it1=models_us[1][0])

    spec = importlib.util.spec_from_file_location("ref_core", os.path.join(REF, "bode", "_core.py"))
    core = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(core)
    rng = np.random.default_rng(20261019)
    out = {}
    for d in (1, 2, 3, 6, 8, 10):
        x = rng.random((17, d))
        x[0] = 0.0
        x[1] = 1.0
        ell = np.concatenate([[0.021], rng.exponential(0.7, size=d)])[:d] if d > 1 else np.array([0.021])
        ell2 = rng.exponential(0.7, size=d) + 0.05
        ss = np.array([rng.gamma(1.0, 2.0)])
        for tag, e in (("a", ell), ("b", ell2)):
            out["x_d%d%s" % (d, tag)] = x
            out["ell_d%d%s" % (d, tag)] = e
            out["ss_d%d%s" % (d, tag)] = ss
            out["xik_d%d%s" % (d, tag)] = core.xik(x, e, ss)
            out["bk_d%d%s" % (d, tag)] = core.bk(e, ss)
    mu = rng.normal(size=(11, 1))
    sig = rng.random((11, 1)) + 0.01
    out["ei_mu"] = mu
    out["ei_sigma"] = sig
    out["ei_ystar"] = np.array(0.3)
    out["ei_min"] = core.ei(mu, sig, 0.3, mode="min")
    out["ei_max"] = core.ei(mu, sig, 0.3, mode="max")
    p = np.array([0.05, 0.3, 0.9, 2.5])
    out["prior_in"] = p
    out["prior_jeffreys"] = core.JeffreysPrior()(p)
    out["prior_beta_2_5"] = core.BetaPrior(a=2, b=5)(p)
    out["prior_gamma_8_1"] = core.GammaPrior(a=8, scale=1)(p)
    out["prior_gamma_1_2"] = core.GammaPrior(a=1, scale=2)(p)
    out["prior_expon_07"] = core.ExponentialPrior(scale=0.7)(p)
    np.savez(os.path.join(HERE, "core_xik_bk.npz"), **out)
    for tag, payload in reference_ekld_vectors().items():
        np.savez_compressed(os.path.join(HERE, "ekld_ref_%s.npz" % tag), **payload)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
