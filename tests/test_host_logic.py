"""CPU: host-side pieces that surround the device path (designs, ensemble sampler, argmax combine, exports)."""
import numpy as np
import pytest
import torch

import bode_b200
from bode_b200 import _emcee, _lhs, dist
from bode_b200._core import JeffreysPrior, UniformPrior


def test_package_exports_reference_names():
    import bode
    for name in ['xik', 'bk', 'JeffreysPrior', 'BetaPrior', 'GammaPrior', 'UniformPrior', 'ExponentialPrior', 'ei', 'KLSampler']:
        assert name in bode_b200.__all__ and hasattr(bode, name)


def test_klsampler_signature_matches_reference():
    import inspect
    sig = inspect.signature(bode_b200.KLSampler.__init__)
    names = list(sig.parameters)
    ref = ['self', 'X', 'Y', 'x_hyp', 'obj_func', 'noisy', 'bounds', 'true_func', 'model_kern', 'num_opt_restarts',
           'num_mc_samples', 'num_quad_points', 'energy', 'nugget', 'lengthscale', 'variance', 'N_avg', 'kld_tol', 'func_name',
           'quad_points', 'quad_points_weight', 'max_it', 'ekld_nugget', 'per_sampled', 'mcmc_acc_low', 'mcmc_acc_upp',
           'mcmc_model', 'mcmc_steps', 'mcmc_final', 'ego_init_perc', 'mcmc_chains', 'mcmc_burn', 'mcmc_thin', 'mcmc_parallel',
           'ego_iter', 'initialize_from_prior', 'variance_prior', 'lengthscale_prior', 'noise_prior', 'mcmc_model_avg']
    assert names[:len(ref)] == ref  # reference order (bode/_sampler.py:93-127); extras come after
    p = sig.parameters
    assert p['nugget'].default == 1e-3 and p['max_it'].default == 50 and p['mcmc_model_avg'].default == 50
    o = inspect.signature(bode_b200.KLSampler.optimize).parameters
    assert list(o) == ['self', 'num_designs', 'verbose', 'plots', 'comp', 'comp_plots'] and o['num_designs'].default == 1000


def test_lhs_center_is_a_permutation_of_cell_centres():
    np.random.seed(1333)
    H = _lhs.lhs(3, samples=7, criterion='center')
    assert H.shape == (7, 3)
    centres = (np.arange(7) + 0.5) / 7
    for j in range(3):
        np.testing.assert_allclose(np.sort(H[:, j]), centres, rtol=0, atol=1e-15)
    # samp_ex1: lhs(1, 3, 'center') -> {1/6, 1/2, 5/6}
    np.testing.assert_allclose(np.sort(_lhs.lhs(1, samples=3, criterion='center')[:, 0]), [1 / 6, 1 / 2, 5 / 6])


def test_lhs_classic_one_point_per_cell_and_rng_order():
    np.random.seed(7)
    H = _lhs.lhs(2, samples=50)
    for j in range(2):
        assert sorted(np.floor(H[:, j] * 50).astype(int).tolist()) == list(range(50))
    # pyDOE consumption order: rand(samples, n) first, then one permutation per column
    np.random.seed(7)
    u = np.random.rand(50, 2)
    p0 = np.random.permutation(range(50))
    a = np.linspace(0, 1, 51)[:50]
    np.testing.assert_allclose(H[:, 0], (u[:, 0] / 50 + a)[p0])
    with pytest.raises(ValueError):
        _lhs.lhs(2, 5, criterion='maximin')


def test_ensemble_sampler_moments_and_shapes():
    np.random.seed(3)
    mu = np.array([1.0, -2.0]); sd = np.array([0.5, 2.0])
    lnp = lambda p: float(-0.5 * np.sum(((p - mu) / sd) ** 2))
    s = _emcee.EnsembleSampler(10, 2, lnp)
    p0 = mu + 0.1 * np.random.randn(10, 2)
    s.run_mcmc(p0, 1500)
    assert s.chain.shape == (10, 1500, 2) and s.acceptance_fraction.shape == (10,)
    flat = s.chain[:, 300:, :].reshape(-1, 2)
    assert np.all(np.abs(flat.mean(0) - mu) < 0.25 * sd)
    assert np.all(np.abs(flat.std(0) / sd - 1) < 0.25)
    assert 0.2 < s.acceptance_fraction.mean() < 0.95


def test_ensemble_sampler_vectorised_path_is_identical():
    mu = np.array([0.3, 0.7, 1.1])
    lnp1 = lambda p: float(-0.5 * np.sum((p - mu) ** 2))
    lnpv = lambda P: -0.5 * np.sum((P - mu) ** 2, axis=1)
    out = []
    for fn, vec in ((lnp1, False), (lnpv, True)):
        np.random.seed(11)
        s = _emcee.EnsembleSampler(6, 3, fn, vectorize=vec)
        s.run_mcmc(mu + 0.05 * np.random.randn(6, 3), 50)
        out.append(s.chain.copy())
    np.testing.assert_array_equal(out[0], out[1])
    with pytest.raises(ValueError):
        _emcee.EnsembleSampler(5, 2, lnp1)


def test_priors_sampling():
    u = UniformPrior(a=0.2, b=0.9)
    x = u.sample(size=100)
    assert x.min() >= 0.2 and x.max() <= 0.9
    assert u(0.5) == pytest.approx(-np.log(0.7)) and u(0.95) == -np.inf
    with pytest.raises(SystemExit):
        JeffreysPrior().sample()


def test_shard_bounds_cover_exactly():
    for n in (1, 7, 64, 100000, 12345):
        for w in (1, 2, 3, 4, 8):
            spans = [dist.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert all(hi >= lo for lo, hi in spans)


def test_combine_argmax_semantics():
    t = lambda v, i: (torch.tensor(v, dtype=torch.float64), torch.tensor(i, dtype=torch.int64))
    assert dist.combine_argmax(*t([1.0, 3.0, 2.0], [5, 17, 40])) == (3.0, 17)
    assert dist.combine_argmax(*t([3.0, 3.0], [90, 17]))[1] == 17          # tie -> lowest global index
    v, i = dist.combine_argmax(*t([float('nan'), 9.0, float('nan')], [30, 2, 12]))
    assert np.isnan(v) and i == 12                                          # NaN wins, first NaN index
    assert dist.combine_argmax(*t([float('-inf'), 1.0], [-1, 8])) == (1.0, 8)  # empty shard ignored
    # agrees with np.argmax on random data split into shards
    rng = np.random.default_rng(0)
    for trial in range(50):
        x = rng.integers(0, 6, size=37).astype(float)
        if trial % 3 == 0:
            x[rng.integers(0, 37, size=2)] = np.nan
        vals, idxs = [], []
        for r in range(4):
            lo, hi = dist.shard_bounds(37, 4, r)
            j = int(np.argmax(x[lo:hi]))
            vals.append(x[lo + j]); idxs.append(lo + j)
        assert dist.combine_argmax(*t(vals, idxs))[1] == int(np.argmax(x))


def test_xik_factor_chebyshev_series_method():
    """The numerical method behind xik_coef_block / k_ekld_reduce, restated in NumPy: the xik factor
    F(x) = sqrt(pi/2) ell (erf((1-x)c) + erf(xc)) is even about x = 1/2, so it is a series in T_j(2(2x-1)^2 - 1) computed
    from 64 Gauss-Chebyshev nodes and cut at its rounding floor; the cut series reproduces F to ~1e-15 with a handful of
    terms, and very short lengthscales are detected as not converged (erf path)."""
    from scipy.special import erf
    M, KMAX = 64, 48
    i = np.arange(M)
    theta = np.pi * (i + 0.5) / M
    xn = 0.5 * (np.cos(theta / 2) + 1.0)
    r = (np.arange(M)[:, None] * (2 * i[None, :] + 1)) % (4 * M)
    cosr = np.cos(np.pi * r / (2 * M))          # exactly reduced argument, as the kernel does with cospi
    rng = np.random.default_rng(3)
    x = np.concatenate([[0.0, 1.0, 0.5], rng.random(20000)])
    w = 2.0 * (2.0 * x - 1.0) ** 2 - 1.0
    terms = {}
    for ell in (0.05, 0.1, 0.3, 1.0, 5.0, 50.0):
        c = 1.0 / (np.sqrt(2.0) * ell)
        F = lambda t: 1.2533141373155001 * ell * (erf((1.0 - t) * c) - erf((0.0 - t) * c))
        a = (2.0 / M) * (F(xn)[None, :] * cosr).sum(1)
        big = np.nonzero(np.abs(a) > 4e-16 * abs(a[0]))[0]
        assert big.max() < KMAX, "series must converge within 48 terms for ell >= 0.05"
        n = int(big.max()) + 1
        terms[ell] = n
        b1 = np.zeros_like(w); b2 = np.zeros_like(w)
        for j in range(n - 1, 0, -1):
            b1, b2 = a[j] + 2.0 * w * b1 - b2, b1
        got = 0.5 * a[0] + w * b1 - b2
        rel = np.abs(got - F(x)) / F(x)
        assert np.percentile(rel, 99) <= 2e-15 and rel.max() <= 1e-14, (ell, rel.max())
    assert terms[5.0] <= 6 and terms[0.3] <= 14 and terms[0.05] <= 36 and terms[50.0] <= 3
    # a lengthscale far below the node spacing near the ends does not converge -> the kernel falls back to erf
    ell = 2e-3
    c = 1.0 / (np.sqrt(2.0) * ell)
    a = (2.0 / M) * ((1.2533141373155001 * ell * (erf((1.0 - xn) * c) - erf((0.0 - xn) * c)))[None, :] * cosr).sum(1)
    assert (np.abs(a[KMAX:]) > 4e-16 * abs(a[0])).any()


# ---- round 2: host-side error handling and the packed argmax record ------------------------------------------------------
def _tiny_sampler(engine, hyp, **kw):
    X0 = np.array([[0.2], [0.5], [0.9]])
    Y0 = np.sin(4 * X0)
    return bode_b200.KLSampler(X0, Y0, False, lambda x: float(np.sin(4 * x[0])), False, [(0, 1)], mcmc_model=True, max_it=1,
                               mcmc_chains=2, mcmc_model_avg=hyp.shape[0], mcmc_steps=40, mcmc_burn=10, mcmc_thin=5,
                               engine=engine, hyper_samples=lambda X, Y, it: hyp, **kw)


def test_failed_hyper_sample_raises_instead_of_selecting_nan():
    """a hyper-sample whose factorisation fails used to poison every score with NaN (and NaN wins the argmax)."""
    from bode_b200._ext import BodeError
    from tests.fake_engine import OracleEngine
    hyp = np.array([[1.0, 0.3], [2.0, 0.4], [0.7, 0.2]])
    kls = _tiny_sampler(OracleEngine(fail_samples=(1,)), hyp)
    with pytest.raises(BodeError) as ei:
        kls.optimize(num_designs=20)
    assert ei.value.status == -5 and "[1]" in str(ei.value)
    bad = hyp.copy(); bad[2, 1] = -0.1   # invalid hyper-parameters (info = -1)
    with pytest.raises(BodeError) as ei:
        _tiny_sampler(OracleEngine(), bad).score_designs(np.random.rand(5, 1))
    assert "invalid hyper-parameters" in str(ei.value)
    ok = _tiny_sampler(OracleEngine(), hyp)
    out = ok.optimize(num_designs=20)
    assert len(ok.chosen_idx) == 1 and np.isfinite(out[3]).all()


def test_vectorised_lnprob_uses_scalar_prior_contract():
    """one out-of-range walker must not reject the whole half-ensemble; scalar and batch paths agree."""
    from tests.fake_engine import OracleEngine
    hyp = np.array([[1.0, 0.3], [2.0, 0.4]])
    kls = _tiny_sampler(OracleEngine(), hyp, variance_prior=UniformPrior(0.0, 3.0), lengthscale_prior=UniformPrior(0.0, 1.0))
    block = np.array([[1.0, 0.3], [5.0, 0.3], [2.0, 0.4], [1.0, -0.2], [0.5, 1.5]])
    got = kls.lnprob(block, None, kls.X, kls.Y)
    one = np.array([kls.lnprob(row, None, kls.X, kls.Y) for row in block])
    assert np.isfinite(got[[0, 2]]).all() and np.isinf(got[[1, 3, 4]]).all()
    np.testing.assert_allclose(got, one, rtol=1e-12)


def test_packed_record_and_c_abi_combine(built_lib):
    """The 16-byte (float64, int64) record the ranks exchange, and bode_argmax_combine's np.argmax semantics (host-only)."""
    from bode_b200 import _ext
    vals = [0.5, 2.0, 2.0, -1.0]
    recs = torch.cat([dist.pack_record(torch.tensor([v]), torch.tensor([i])) for v, i in zip(vals, [40, 30, 20, 10])])
    assert recs.dtype == torch.int64 and recs.numel() == 8
    assert _ext.argmax_combine(recs.numpy()) == (2.0, 20)                       # tie -> lowest index
    nan = torch.cat([recs, dist.pack_record(torch.tensor([float("nan")]), torch.tensor([99]))])
    v, i = _ext.argmax_combine(nan.numpy())
    assert np.isnan(v) and i == 99                                               # NaN counts as the maximum
    empty = torch.cat([dist.pack_record(torch.tensor([float("-inf")]), torch.tensor([-1])), recs[:2]])
    assert _ext.argmax_combine(empty.numpy()) == (0.5, 40)                       # idx < 0: rank without candidates
    host = dist.combine_argmax(torch.tensor(vals, dtype=torch.float64), torch.tensor([40, 30, 20, 10]))
    assert host == (2.0, 20)
    # bode_ekld_argmax_p2p marks a peer that never delivered with {NaN, -2}: the combine reports it, it does not pick
    late = torch.cat([recs, dist.pack_record(torch.tensor([float("nan")]), torch.tensor([-2]))])
    with pytest.raises(_ext.BodeError) as e:
        _ext.argmax_combine(late.numpy())
    assert e.value.status == -6


def test_reference_driver_source_runs_against_this_package(tmp_path):
    """"Drops in under the existing tests/samp_ex*.py drivers": the reference's own tests/samp_ex1.py source is executed
    unmodified with `bode` resolved to this repository and the third-party stand-ins of tests/stubs first on sys.path.
    Build container only (the reference tree is not shipped); the engine is the oracle-backed stand-in because there is no
    GPU here -- the same driver settings run on the GPU through examples/samp_ex1.py (tests/test_gpu_parity.py)."""
    import os
    import pickle
    import subprocess
    import sys
    ref = "/root/reference/tests/samp_ex1.py"
    if not os.path.exists(ref):
        pytest.skip("reference tree not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    runner = tmp_path / "run_ref_driver.py"
    runner.write_text('''
import runpy, sys
sys.path.insert(0, %r)
from tests.stubs import install
install.install()
import bode                                  # this repository's drop-in alias, cached before the driver edits sys.path
from tests.fake_engine import OracleEngine
_orig = bode.KLSampler
class ShortChains(_orig):                     # same constructor call as the driver's, shorter chains, CPU test engine
    def __init__(self, *a, **k):
        k.update(mcmc_steps=120, mcmc_burn=20, mcmc_thin=10, engine=OracleEngine())
        super().__init__(*a, **k)
bode.KLSampler = ShortChains
import builtins
builtins.quit = lambda *a: None               # the driver ends with quit()
runpy.run_path(%r, run_name="__main__")
''' % (root, ref))
    env = dict(os.environ, PYTHONPATH=root, OMP_NUM_THREADS="1")
    subprocess.check_call([sys.executable, str(runner)], cwd=str(tmp_path), env=env, timeout=600,
                          stdout=subprocess.DEVNULL)
    out = tmp_path / "mcmc_ex1_n=3_it=1"
    X = np.load(out / "X.npy"); kld = np.load(out / "kld.npy"); mu = np.load(out / "mu_qoi.npy")
    assert X.shape == (4, 1) and kld.shape == (1, 1000) and np.isfinite(kld).all() and mu.shape == (2,)
    with open(out / "comp.pkl", "rb") as f:
        mu_us, sigma_us, models, models_us = pickle.load(f)
    assert len(mu_us) == 2 and models[0][0].shape == (60, 2)


def test_log1p_form_default_is_loud_when_it_changes_the_choice(capsys):
    """the drop-in class evaluates the log1p form by default; when that changes the chosen design against the
    reference's literal expression, it says so."""
    from tests.fake_engine import OracleEngine

    class TwoFaced(OracleEngine):          # literal form (0) and log1p form (1) disagree on purpose
        def score(self, X_design, idx_offset=0, form=0, want_var=True, **kw):
            r = super().score(X_design, idx_offset=idx_offset, form=form, want_var=want_var, **kw)
            if form == 0:
                j = (int(torch.argmax(r["score"])) + 1) % r["score"].numel()   # a NaN next to the true winner
                sc = r["score"].clone(); sc[j] = float("nan")
                r.update(score=sc, best=sc[j:j + 1].clone(), best_idx=torch.tensor([idx_offset + j]),
                         rec=torch.cat([sc[j:j + 1].clone().view(torch.int64), torch.tensor([idx_offset + j])]))
            return r

    hyp = np.array([[1.0, 0.3], [2.0, 0.4]])
    kls = _tiny_sampler(TwoFaced(), hyp)
    res = kls.score_designs(np.linspace(0.05, 0.95, 12)[:, None])
    out = capsys.readouterr().out
    assert "WARNING: ekld_form=1" in out and res["idx_best_literal"] == (res["idx_best"] + 1) % 12
    quiet = _tiny_sampler(OracleEngine(), hyp)
    quiet.score_designs(np.linspace(0.05, 0.95, 12)[:, None])
    assert "WARNING" not in capsys.readouterr().out
    off = _tiny_sampler(TwoFaced(), hyp, ekld_form_check=False)
    assert "idx_best_literal" not in off.score_designs(np.linspace(0.05, 0.95, 12)[:, None])
    # the check is one scoring call plus a re-run of the reduction on its planes, not two scoring calls
    assert quiet.engine.calls["score"] == 1 and quiet.engine.calls["rescore"] == 1
    assert off.engine.calls["score"] == 1 and off.engine.calls["rescore"] == 0
    with pytest.raises(RuntimeError, match="rescore"):
        off.engine.rescore()
