"""Device-resident front end of the EKLD hot path: torch tensors in HBM -> C ABI (``*_dev`` entry points).

``EkldEngine`` owns the per-iteration device state that replaces the reference's per-(candidate, sample) GPy
objects (``make_mcmc_model``, ``_sampler.py:449-462``): one workspace holding, for every hyper-sample, the
Cholesky-derived factors the scoring kernel streams.  PyTorch is plumbing only (allocation, streams, H2D/D2H).

No CPU fallback: constructing an engine without a CUDA device raises.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _ext

GPY_JITTER = 1e-8  # GPy ExactGaussianInference adds 1e-8 to the diagonal (fixture-verified, see oracle/)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class EkldEngine:
    """prepare(X, Y, hyp) once per outer iteration, then score(X_design) / us_variance(X_grid) / mu_sigma()."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("bode_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _ext.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n = self.d = self.S = 0
        self.ws = None
        self.info = None
        self._scratch = {}
        self.last_kernel_ms = None
        self.quad = None
        self._ndim = 0
        self._prepared = False
        self._comm = None          # our own NCCL communicator (init_nccl), world size / rank
        self._xchg = None          # peer-memory mailboxes (init_p2p): the argmax exchange as one kernel over NVLink
        self._world = 1
        self._rank = 0
        self._gathered = None
        self._group = None
        self.jitter_used = None    # per-sample jitter of the last prepare() when jitchol retries were needed

    # ---- multi-GPU argmax exchange ------------------------------------------------------------------------
    def init_nccl(self, group=None):
        """Create the communicator ``bode_ekld_argmax_nccl`` runs on: rank 0 draws an ncclUniqueId, torch.distributed
        (the caller's process group) carries its 128 bytes to the other ranks.  One communicator per engine."""
        import torch.distributed as tdist
        world, rank = tdist.get_world_size(group), tdist.get_rank(group)
        with torch.cuda.device(self.device):
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                buf = (ctypes.c_ubyte * 128)()
                _ext.check(self.lib.bode_nccl_unique_id(buf))
                uid = torch.tensor(list(buf), dtype=torch.uint8)
            backend = tdist.get_backend(group)
            t = uid.to(self.device) if backend == "nccl" else uid
            tdist.broadcast(t, src=tdist.get_global_rank(group, 0) if group is not None else 0, group=group)
            raw = bytes(t.cpu().tolist())
            comm = ctypes.c_void_p()
            _ext.check(self.lib.bode_nccl_comm_create(ctypes.c_char_p(raw), world, rank, ctypes.byref(comm)))
            self._comm, self._world, self._rank = comm, world, rank
            self._group = group
            self._gathered = torch.empty(2 * world, dtype=torch.int64, device=self.device)
        return self

    def init_p2p(self, group=None):
        """Peer-memory form of the exchange (``bode_ekld_argmax_p2p``): every rank's mailbox is exported as a CUDA IPC
        handle, torch.distributed carries the 64-byte handles, every rank maps its peers'.  One process per GPU, same
        node, peer access between the GPUs.  Collective; raises BodeError on the rank where the mapping fails."""
        import torch.distributed as tdist
        world, rank = tdist.get_world_size(group), tdist.get_rank(group)
        with torch.cuda.device(self.device):
            handle = (ctypes.c_ubyte * 64)()
            x = ctypes.c_void_p()
            _ext.check(self.lib.bode_xchg_create(world, rank, ctypes.byref(x), handle))
            mine = torch.tensor(list(handle), dtype=torch.uint8)
            on_dev = tdist.get_backend(group) == "nccl"
            mine = mine.to(self.device) if on_dev else mine
            allh = torch.empty(64 * world, dtype=torch.uint8, device=mine.device)
            tdist.all_gather_into_tensor(allh, mine, group=group)
            raw = bytes(allh.cpu().tolist())
            try:
                _ext.check(self.lib.bode_xchg_connect(x, ctypes.c_char_p(raw)))
            except _ext.BodeError:
                self.lib.bode_xchg_destroy(x)
                raise
            self._xchg, self._world, self._rank, self._group = x, world, rank, group
            self._gathered = torch.empty(2 * world, dtype=torch.int64, device=self.device)
            torch.cuda.synchronize(self.device)
        tdist.barrier(group)  # nobody stores into a mailbox that is not mapped everywhere yet
        return self

    def init_exchange(self, group=None):
        """The multi-GPU argmax exchange: peer memory when every rank can map every mailbox (one node), else NCCL.
        ``BODE_EXCHANGE=nccl|p2p`` forces one.  Collective."""
        import os
        import torch.distributed as tdist
        want = os.environ.get("BODE_EXCHANGE", "auto")
        if want == "nccl":
            return self.init_nccl(group)
        if want == "p2p":
            return self.init_p2p(group)
        ok = 1
        try:
            self.init_p2p(group)
        except _ext.BodeError:
            ok = 0
            tdist.barrier(group)  # (the barrier the failed init_p2p did not reach)
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device if tdist.get_backend(group) == "nccl" else "cpu")
        tdist.all_reduce(flag, op=tdist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            self.close_p2p()
            return self.init_nccl(group)
        return self

    def has_exchange(self):
        return self._comm is not None or self._xchg is not None

    def _enqueue_exchange(self, rec):
        if self._xchg is not None:
            _ext.check(self.lib.bode_ekld_argmax_p2p(self._xchg, _ptr(rec), _ptr(self._gathered), self._stream()))
        else:
            _ext.check(self.lib.bode_ekld_argmax_nccl(self._comm, _ptr(rec), _ptr(self._gathered), self._stream()))

    def close_p2p(self):
        if self._xchg is not None:
            torch.cuda.synchronize(self.device)
            self.lib.bode_xchg_destroy(self._xchg)
            self._xchg = None

    def close_nccl(self):
        """Release the exchange.  StepGraph objects built on this engine must be deleted first: NCCL does not destroy a
        communicator while a CUDA graph that captured one of its collectives is alive."""
        self.close_p2p()
        if self._comm is not None:
            torch.cuda.synchronize(self.device)
            self.lib.bode_nccl_comm_destroy(self._comm)
            self._comm = None

    def exchange_best(self, rec):
        """All ranks: one 16-byte ncclAllGather of the packed (best, idx) record on the compute stream, ONE device->host
        copy, np.argmax semantics on the host.  ``rec``: the int64[2] tensor of ``score()['rec']`` (or None for a rank
        without candidates)."""
        if not self.has_exchange():
            raise RuntimeError("call init_exchange() / init_nccl() / init_p2p() first")
        with torch.cuda.device(self.device):
            if rec is None:
                rec = torch.tensor([0, -1], dtype=torch.int64, device=self.device)
            self._enqueue_exchange(rec)
            host = self._gathered.cpu().numpy()
        return _ext.argmax_combine(host)

    # ---- helpers -------------------------------------------------------------------------------------------
    def _dev(self, a, shape=None):
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.device, dtype=torch.float64).contiguous()
        else:
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device, non_blocking=True)
        if shape is not None:
            t = t.reshape(shape)
        return t

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _buf(self, name, numel, dtype):
        t = self._scratch.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(int(numel), dtype=dtype, device=self.device)
            self._scratch[name] = t
        return t

    # ---- K1 ------------------------------------------------------------------------------------------------
    def prepare(self, X, Y, hyp, nugget_var, jitter=GPY_JITTER, quad=None, jitchol_retries=0):
        """Factorise every hyper-sample (``bode_ekld_prepare_dev``).  hyp: (S, d+1) or (S, d+2) reference layout.

        quad=(X_quad, weights|None) switches to quadrature mode: the kernel integrals xik / bk are replaced by
        (weighted) means over the quadrature points, here and in the following ``score`` calls.

        jitchol_retries > 0 mirrors GPy's ``jitchol`` (what ``make_mcmc_model`` does implicitly, ``_sampler.py:453``): a
        hyper-sample whose Cholesky fails is re-factorised with an extra ``mean(diag) * 1e-6 * 10^k``, k = 0.. on its
        diagonal (the healthy samples keep their bits); costs one device->host read of ``info`` per call."""
        with torch.cuda.device(self.device):
            Xd = self._dev(X)
            n, d = Xd.shape
            Yd = self._dev(Y).reshape(-1)
            if Yd.numel() != n:
                raise ValueError("Y must have one row per observation")
            H = self._dev(hyp)
            if H.dim() != 2:
                raise ValueError("hyp must be 2-D (S, d+1 | d+2)")
            S, ndim = H.shape
            nbytes = self.lib.bode_ekld_workspace_bytes(n, d, S)
            if nbytes == 0:
                _ext.check(-2)
            if self.ws is None or self.ws.numel() < nbytes:
                self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.info = self._buf("info", S, torch.int32)[:S]
            self.quad = None
            if quad is None:
                _ext.check(self.lib.bode_ekld_prepare_dev(_ptr(Xd), _ptr(Yd), _ptr(H), n, d, S, ndim, float(nugget_var),
                                                          float(jitter), _ptr(self.ws), self.ws.numel(), _ptr(self.info),
                                                          self._stream()))
            else:
                Xq = self._dev(quad[0])
                if Xq.dim() == 1:
                    Xq = Xq.reshape(-1, 1)
                if Xq.shape[1] != d:
                    raise ValueError("quadrature points have %d columns, observations have %d" % (Xq.shape[1], d))
                wq = None if quad[1] is None else self._dev(quad[1]).reshape(-1)
                if wq is not None and wq.numel() != Xq.shape[0]:
                    raise ValueError("one weight per quadrature point")
                Nq = Xq.shape[0]
                qs = self._buf("quad_scratch", self.lib.bode_quad_scratch_bytes(n, S, Nq), torch.uint8)
                if self.has_exchange() and self._world > 1 and Nq >= 64 * self._world:
                    # several GPUs (init_nccl): the O(Nq^2 S) row means of the quadrature points are computed in row shards
                    # and all-gathered (torch.distributed plumbing); sigma0_q is then formed from identical numbers in the
                    # same order on every rank -- bit-identical to the single-GPU preparation
                    import torch.distributed as tdist
                    per = (Nq + self._world - 1) // self._world
                    lo = min(Nq, self._rank * per)
                    hi = min(Nq, lo + per)
                    loc = torch.zeros((S, per), dtype=torch.float64, device=self.device)
                    if hi > lo:
                        part = torch.empty((S, hi - lo), dtype=torch.float64, device=self.device)
                        _ext.check(self.lib.bode_quad_rowmean_dev(_ptr(H), S, ndim, d, _ptr(Xq[lo:hi].contiguous()), hi - lo, _ptr(Xq),
                                                                  _ptr(wq), Nq, _ptr(part), _ptr(qs), self._stream()))
                        loc[:, : hi - lo] = part
                    allr = torch.empty((self._world, S, per), dtype=torch.float64, device=self.device)
                    tdist.all_gather_into_tensor(allr, loc, group=self._group)
                    rmq = allr.permute(1, 0, 2).reshape(S, self._world * per)[:, :Nq].contiguous()
                    _ext.check(self.lib.bode_ekld_prepare_quad_rm_dev(_ptr(Xd), _ptr(Yd), _ptr(H), n, d, S, ndim, float(nugget_var),
                                                                      float(jitter), _ptr(Xq), _ptr(wq), Nq, _ptr(rmq), _ptr(self.ws),
                                                                      self.ws.numel(), _ptr(self.info), _ptr(qs), self._stream()))
                    self._keep_rm = rmq
                else:
                    _ext.check(self.lib.bode_ekld_prepare_quad_dev(_ptr(Xd), _ptr(Yd), _ptr(H), n, d, S, ndim, float(nugget_var),
                                                                   float(jitter), _ptr(Xq), _ptr(wq), Nq, _ptr(self.ws),
                                                                   self.ws.numel(), _ptr(self.info), _ptr(qs), self._stream()))
                self.quad = (Xq, wq, Nq)
            self.n, self.d, self.S = n, d, S
            self._ndim = ndim
            self._prepared = True
            self._planes_ok = False
            self._keep = (Xd, Yd, H)  # keep inputs alive until the stream has consumed them
            self.jitter_used = None
            if jitchol_retries > 0 and quad is None:
                info = self.info.cpu().numpy()
                if (info > 0).any():
                    noise = (H[:, -1] ** 2).cpu().numpy() if ndim == d + 2 else np.full(S, float(nugget_var))
                    diag = H[:, 0].cpu().numpy() + noise + jitter
                    jit = np.full(S, float(jitter))
                    for k in range(int(jitchol_retries)):
                        bad = info > 0
                        if not bad.any():
                            break
                        jit[bad] = jitter + diag[bad] * 1e-6 * 10.0 ** k
                        J = self._dev(jit)
                        _ext.check(self.lib.bode_ekld_prepare_jitter_dev(_ptr(Xd), _ptr(Yd), _ptr(H), n, d, S, ndim,
                                                                         float(nugget_var), _ptr(J), _ptr(self.ws),
                                                                         self.ws.numel(), _ptr(self.info), self._stream()))
                        info = self.info.cpu().numpy()
                    self.jitter_used = jit
        return self.info

    def loglik(self, X, Y, hyp, nugget_var, jitter=GPY_JITTER):
        """GP log marginal likelihood of every row of ``hyp`` (``get_likelihood``, ``_sampler.py:221-232``) as a host
        NumPy array: one likelihood-only launch of the prepare kernel.  Observations are cached on the device between
        calls with the same (X, Y) objects (the MCMC calls this thousands of times per outer iteration)."""
        with torch.cuda.device(self.device):
            key = (id(X), id(Y))
            if getattr(self, "_ll_key", None) != key:
                self._ll_X = self._dev(X)
                self._ll_Y = self._dev(Y).reshape(-1)
                self._ll_key = key
                self._ll_hold = (X, Y)
            Xd, Yd = self._ll_X, self._ll_Y
            n, d = Xd.shape
            H = self._dev(np.atleast_2d(hyp))
            S, ndim = H.shape
            nbytes = self.lib.bode_ekld_workspace_bytes(n, d, S)
            if nbytes == 0:
                _ext.check(-2)
            if self.ws is None or self.ws.numel() < nbytes:
                self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            info = self._buf("info", S, torch.int32)[:S]
            out = self._buf("loglik", S, torch.float64)[:S]
            _ext.check(self.lib.bode_gp_loglik_dev(_ptr(Xd), _ptr(Yd), _ptr(H), n, d, S, ndim, float(nugget_var), float(jitter),
                                                   _ptr(self.ws), self.ws.numel(), _ptr(info), _ptr(out), self._stream()))
            self._prepared = False  # the workspace no longer holds a scoring state
            return out.cpu().numpy()

    def sample_terms(self):
        """(S, 8) device tensor: variance, noise_var, sigma0, sigma1, mu1, logdet, Y^T A^-1 Y, loglik."""
        if not self._prepared:
            raise RuntimeError("call prepare() first")
        with torch.cuda.device(self.device):
            out = torch.empty((self.S, 8), dtype=torch.float64, device=self.device)
            _ext.check(self.lib.bode_ekld_sample_terms_dev(_ptr(self.ws), self.n, self.d, self.S, _ptr(out), self._stream()))
        return out

    def mu_sigma(self):
        """``get_mu_sigma`` (``_sampler.py:434-444``): (mean_s mu_1, mean_s sigma_1) as Python floats."""
        if not self._prepared:
            raise RuntimeError("call prepare() first")
        with torch.cuda.device(self.device):
            out = torch.empty(2, dtype=torch.float64, device=self.device)
            _ext.check(self.lib.bode_ekld_mu_sigma_dev(_ptr(self.ws), self.n, self.d, self.S, _ptr(out), self._stream()))
            mu, sg = out.cpu().tolist()
        return mu, sg

    # ---- K2-K4 ---------------------------------------------------------------------------------------------
    def score(self, X_design, idx_offset=0, form=_ext.BODE_FORM_REFERENCE, want_var=True, timed=False, events=None,
              want_logk=True):
        """Score every row of X_design (device or host array).  Returns device tensors:
        dict(score (Nd,), var (Nd,)|None, logk (S, Nd)|None, best (1,), best_idx (1,) int64, rec int64[2] = the packed
        16-byte (best, best_idx) record).  ``logk`` is a view of an engine-owned buffer, valid until the next call."""
        if not self._prepared:
            raise RuntimeError("call prepare() first")
        with torch.cuda.device(self.device):
            Xd = self._dev(X_design)
            if Xd.dim() == 1:
                Xd = Xd.reshape(1, -1)
            Nd, d = Xd.shape
            if d != self.d:
                raise ValueError("X_design has %d columns, observations have %d" % (d, self.d))
            # want_logk=False: the per-sample matrix is not materialised (the kernels keep their two S x Nd planes in scratch)
            logk = self._buf("logk", self.S * Nd, torch.float64)[: self.S * Nd].view(self.S, Nd) if want_logk else None
            score = torch.empty(Nd, dtype=torch.float64, device=self.device)
            var = torch.empty(Nd, dtype=torch.float64, device=self.device) if want_var else None
            rec = torch.empty(2, dtype=torch.int64, device=self.device)  # packed {float64 best, int64 best_idx}
            best, best_idx = rec[:1].view(torch.float64), rec[1:]
            scr = self._buf("scratch", 256 + self.lib.bode_ekld_score_scratch_bytes(Nd, self.S if want_logk else 2 * self.S), torch.uint8)
            args = [_ptr(self.ws), self.n, self.d, self.S, _ptr(Xd), Nd, int(idx_offset), int(form), _ptr(logk),
                    _ptr(score), _ptr(var), _ptr(best), _ptr(best_idx), _ptr(scr), self._stream()]
            if self.quad is not None:
                Xq, wq, Nq = self.quad
                xi = self._buf("xi", self.S * Nd, torch.float64)[: self.S * Nd].view(self.S, Nd)
                _ext.check(self.lib.bode_ekld_score_quad_dev(_ptr(self.ws), self.n, self.d, self.S, _ptr(self._keep[2]),
                                                             self._ndim, _ptr(Xd), Nd, int(idx_offset), int(form), _ptr(Xq),
                                                             _ptr(wq), Nq, _ptr(xi), _ptr(logk), _ptr(score), _ptr(var),
                                                             _ptr(best), _ptr(best_idx), _ptr(scr), self._stream()))
                self._keep_x = Xd
                return dict(score=score, var=var, logk=logk, best=best, best_idx=best_idx, rec=rec, xi=xi)
            if events is not None:
                # (start, stop) torch.cuda.Event pair recorded around the scoring kernel, no synchronisation
                _ext.check(self.lib.bode_ekld_score_dev_events(*args, ctypes.c_void_p(events[0].cuda_event),
                                                               ctypes.c_void_p(events[1].cuda_event)))
            elif timed:
                ms = ctypes.c_float(0.0)
                _ext.check(self.lib.bode_ekld_score_dev_timed(*args, ctypes.c_void_p(ctypes.addressof(ms))))
                self.last_kernel_ms = ms.value
            else:
                _ext.check(self.lib.bode_ekld_score_dev(*args))
            self._keep_x = Xd
            # the sigma_x / w.k planes stay intact in scratch only when neither logk nor var was asked for: rescore() needs that
            self._planes_ok = (not want_var) and (not want_logk)
        return dict(score=score, var=var, logk=logk, best=best, best_idx=best_idx, rec=rec)

    def rescore(self, idx_offset=0, form=_ext.BODE_FORM_REFERENCE, want_var=True):
        """The other EKLD form for the design of the previous ``score(..., want_var=False, want_logk=False)`` call: only the
        reduction kernel runs again, on the planes that call left in scratch (``bode_ekld_rescore_dev``).  Same return
        value as ``score`` without ``logk``."""
        if not getattr(self, "_planes_ok", False) or self.quad is not None:
            raise RuntimeError("rescore() must follow score(want_var=False, want_logk=False) on the same prepared state")
        with torch.cuda.device(self.device):
            Xd = self._keep_x
            Nd = Xd.shape[0]
            score = torch.empty(Nd, dtype=torch.float64, device=self.device)
            var = torch.empty(Nd, dtype=torch.float64, device=self.device) if want_var else None
            rec = torch.empty(2, dtype=torch.int64, device=self.device)
            best, best_idx = rec[:1].view(torch.float64), rec[1:]
            scr = self._buf("scratch", 256 + self.lib.bode_ekld_score_scratch_bytes(Nd, 2 * self.S), torch.uint8)
            _ext.check(self.lib.bode_ekld_rescore_dev(_ptr(self.ws), self.n, self.d, self.S, _ptr(Xd), Nd, int(idx_offset), int(form),
                                                      _ptr(score), _ptr(var), _ptr(best), _ptr(best_idx), _ptr(scr), self._stream()))
            self._planes_ok = not want_var  # (the variance pass turns the sigma_x plane into log EKLD)
        return dict(score=score, var=var, logk=None, best=best, best_idx=best_idx, rec=rec)

    def us_variance(self, X_grid, idx_offset=0):
        """Uncertainty-sampling comparator (``_sampler.py:506-514``): mean_s latent predictive variance + argmax."""
        if not self._prepared:
            raise RuntimeError("call prepare() first")
        with torch.cuda.device(self.device):
            Xg = self._dev(X_grid)
            Ng, d = Xg.shape
            if d != self.d:
                raise ValueError("X_grid has %d columns, observations have %d" % (d, self.d))
            sigx = self._buf("logk", self.S * Ng, torch.float64)[: self.S * Ng].view(self.S, Ng)
            pv = torch.empty(Ng, dtype=torch.float64, device=self.device)
            best = torch.empty(1, dtype=torch.float64, device=self.device)
            best_idx = torch.empty(1, dtype=torch.int64, device=self.device)
            scr = self._buf("scratch", 256 + self.lib.bode_ekld_score_scratch_bytes(Ng, 0), torch.uint8)
            _ext.check(self.lib.bode_us_variance_dev(_ptr(self.ws), self.n, self.d, self.S, _ptr(Xg), Ng, int(idx_offset),
                                                     _ptr(sigx), _ptr(pv), _ptr(best), _ptr(best_idx), _ptr(scr),
                                                     self._stream()))
            self._keep_x = Xg
        return dict(pred_var=pv, sigma_x=sigx, best=best, best_idx=best_idx)


class StepGraph(object):
    """One outer-iteration step -- copy X, Y, hyp (and, if it lives on the host, the design block) to the device, prepare,
    score, and on several GPUs the 16-byte argmax all-gather -- captured ONCE into a CUDA graph and replayed: a step then
    costs one graph launch instead of half a dozen host-side launches, which is what a 12 500-candidate shard of an 8-GPU
    run is otherwise bound by.  The tensors passed in are read at every replay (update them in place: pinned host
    tensors for X, Y, hyp, the design wherever it lives); results are the engine-owned tensors of ``self.out``.

        sg = StepGraph(engine, hX, hY, hH, Xd, nugget_var, idx_offset=lo)
        best, idx = sg.run()          # replay + one small device->host read (the winner)
    """

    def __init__(self, engine, X, Y, hyp, X_design, nugget_var, idx_offset=0, form=_ext.BODE_FORM_REFERENCE, quad=None,
                 want_var=False, want_logk=False, jitter=GPY_JITTER):
        self.eng = engine
        dev = engine.device

        copy_stream = torch.cuda.Stream(device=dev)
        host_design = isinstance(X_design, torch.Tensor) and X_design.device.type == "cpu"

        def body():
            to = lambda t: t.to(dev, non_blocking=True) if isinstance(t, torch.Tensor) else t
            cur = torch.cuda.current_stream()
            xd = X_design
            dX, dY, dH = to(X), to(Y), to(hyp)
            if host_design:  # the design block's host->device copy does not depend on the factorisations: a forked branch
                copy_stream.wait_stream(cur)  # of the graph (behind the small copies: one DMA queue), joined before scoring
                with torch.cuda.stream(copy_stream):
                    xd = X_design.to(dev, non_blocking=True)
            engine.prepare(dX, dY, dH, nugget_var, jitter, quad=quad)
            if host_design:
                cur.wait_stream(copy_stream)  # (xd is held by this object for the life of the graph: no record_stream needed)
            else:
                xd = to(xd)
            self._xd = xd
            r = engine.score(xd, idx_offset=idx_offset, form=form, want_var=want_var, want_logk=want_logk)
            if engine.has_exchange():
                engine._enqueue_exchange(r["rec"])
                self._host_rec.copy_(engine._gathered.view(-1), non_blocking=True)
            else:
                self._host_rec.copy_(r["rec"], non_blocking=True)
            return r

        # the winner record(s) land in pinned host memory as the graph's last node: reading them costs one stream
        # synchronisation, no separate device->host call
        nrec = 2 * (engine._gathered.numel() // 2 if engine.has_exchange() else 1)
        self._host_rec = torch.empty(nrec, dtype=torch.int64).pin_memory()
        with torch.cuda.device(dev):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up outside the capture: function attributes, allocator pools, NCCL channels
                for _ in range(2):
                    body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = body()
            # the graph replays against the engine buffers that existed at capture time: keep them alive even if the engine
            # later outgrows them (a larger design would otherwise hand their memory back to the allocator)
            self._hold = (dict(engine._scratch), engine.ws, engine.info, engine._gathered, getattr(engine, "_keep", None),
                          getattr(engine, "_keep_x", None), X, Y, hyp, X_design, self._xd, copy_stream)

    def launch(self):
        """Enqueue the step (no host synchronisation); results land in ``self.out`` / the engine's gather buffer."""
        self.graph.replay()

    def winner(self):
        """(best score, global index) after ``launch``: one device->host copy, np.argmax semantics across ranks."""
        torch.cuda.current_stream(self.eng.device).synchronize()
        rec = self._host_rec.numpy()
        if self.eng.has_exchange():
            return _ext.argmax_combine(rec.reshape(-1, 2))
        return float(rec[:1].view(np.float64)[0]), int(rec[1])

    def run(self):
        self.launch()
        return self.winner()


def ekld_score(X, Y, hyp, X_design, nugget=1e-3, jitter=GPY_JITTER, form=_ext.BODE_FORM_REFERENCE, engine=None,
               want_kld=False, quad=None):
    """One-call public API with host (NumPy) buffers: prepare + score + summary, results back on the host.

    Returns dict(score, var, argmax, best, mu_Q, sigma_Q, info[, kld]).  This is the dense generalisation of
    ``KLSampler.mcmc_kld`` (``_sampler.py:483-492``) over all rows of ``X_design``.
    """
    eng = EkldEngine() if engine is None else engine
    info = eng.prepare(X, Y, hyp, nugget ** 2, jitter, quad=quad)
    r = eng.score(X_design, form=form)
    mu, sg = eng.mu_sigma()
    out = dict(score=r["score"].cpu().numpy(), var=r["var"].cpu().numpy(), argmax=int(r["best_idx"].item()),
               best=float(r["best"].item()), mu_Q=mu, sigma_Q=sg, info=info.cpu().numpy())
    if want_kld:
        out["kld"] = torch.exp(r["logk"]).cpu().numpy()
    return out
