"""Generate the committed golden fixtures from the reference tree (run in the build container only).

    python tests/golden/make_golden.py

Inputs (read-only, never present on the GPU box):
  /root/reference/tests/mcmc_ex1_n=3_it=1/comp.pkl   -- the reference's only shipped result fixture
        (pickle of ``comp_log = (mu_us, sigma_us, models, models_us)``, written by tests/samp_ex1.py:116-117)
  /root/reference/bode/_core.py                       -- imported by file path (needs only numpy/scipy) to
        generate known-answer vectors for ``xik`` / ``bk`` / ``ei`` / the prior log-pdfs.
Outputs (committed):
  tests/golden/comp_ex1.npz      hyper-samples + Q summaries of samp_ex1 at iteration 0 / 1
  tests/golden/core_xik_bk.npz   inputs and reference outputs of xik, bk, ei, priors
"""
import importlib.util
import os
import pickle

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    with open(os.path.join(REF, "tests", "mcmc_ex1_n=3_it=1", "comp.pkl"), "rb") as f:
        mu_us, sigma_us, models, models_us = pickle.load(f)
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
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
