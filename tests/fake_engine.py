"""CPU stand-in for ``EkldEngine`` backed by the oracle -- TEST INFRASTRUCTURE ONLY (host-logic tests of ``KLSampler`` and
``bode_b200.dist`` in a container without a GPU).  Same method surface and return types (CPU torch tensors); never
imported by the product."""
import numpy as np
import torch

from oracle import ekld_oracle as orc


class OracleEngine(object):
    device = torch.device("cpu")

    def __init__(self, fail_samples=()):
        self.fail_samples = tuple(fail_samples)  # hyper-sample rows reported as not positive definite
        self.jitter_used = None
        self._comm = None
        self.calls = dict(prepare=0, score=0, rescore=0, loglik=0, us=0)
        self._planes = None

    def prepare(self, X, Y, hyp, nugget_var, jitter=1e-8, quad=None, jitchol_retries=0):
        self.X, self.Y, self.hyp = np.array(X, dtype=float), np.array(Y, dtype=float), np.array(hyp, dtype=float)
        self.nugget, self.quad = float(np.sqrt(nugget_var)), quad
        self.calls["prepare"] += 1
        self._planes = None
        info = np.zeros(self.hyp.shape[0], dtype=np.int32)
        for s in self.fail_samples:
            info[s] = 1
        for s, row in enumerate(self.hyp):
            if not (np.isfinite(row).all() and (row[:1 + self.X.shape[1]] > 0).all()):
                info[s] = -1
        self.info = torch.from_numpy(info)
        return self.info

    def mu_sigma(self):
        return orc.get_mu_sigma(self.X, self.Y, self.hyp, self.nugget)

    def rescore(self, idx_offset=0, form=0, want_var=True):
        # like EkldEngine.rescore: only after score(want_var=False, want_logk=False) on the same prepared state
        if self._planes is None:
            raise RuntimeError("rescore() must follow score(want_var=False, want_logk=False) on the same prepared state")
        self.calls["rescore"] += 1
        self.calls["score"] -= 1
        return self.score(self._planes, idx_offset=idx_offset, form=form, want_var=want_var, want_logk=not want_var)

    def score(self, X_design, idx_offset=0, form=0, want_var=True, **kw):
        self.calls["score"] += 1
        Xd = np.atleast_2d(np.asarray(X_design, dtype=float))
        self._planes = Xd if (not want_var and not kw.get("want_logk", True)) else None
        r = orc.score(self.X, self.Y, Xd, self.hyp, nugget=self.nugget, quad=self.quad)
        sc = torch.from_numpy(r["score"].copy())
        j = int(np.argmax(r["score"]))
        rec = torch.cat([sc[j:j + 1].clone().view(torch.int64), torch.tensor([idx_offset + j])])
        with np.errstate(invalid="ignore", divide="ignore"):
            logk = torch.from_numpy(np.log(r["kld"]))
        return dict(score=sc, var=torch.from_numpy(r["var"].copy()) if want_var else None, logk=logk,
                    best=sc[j:j + 1].clone(), best_idx=torch.tensor([idx_offset + j]), rec=rec)

    def us_variance(self, X_grid, idx_offset=0):
        self.calls["us"] += 1
        pv, j = orc.us_variance(self.X, self.Y, np.asarray(X_grid, dtype=float), self.hyp, self.nugget)
        pv = torch.from_numpy(pv)
        return dict(pred_var=pv, best=pv[j:j + 1].clone(), best_idx=torch.tensor([idx_offset + j]))

    def loglik(self, X, Y, hyp, nugget_var, jitter=1e-8):
        self.calls["loglik"] += 1
        return orc.log_likelihoods(np.asarray(X, dtype=float), np.asarray(Y, dtype=float), np.atleast_2d(hyp), float(np.sqrt(nugget_var)))
