"""Multi-GPU candidate sharding: one process per GPU, independent candidate blocks, ONE tiny collective.

The EKLD of different candidates shares only the read-only per-sample factors, so the path shards over
candidate blocks with no data-path collective: every rank runs the prepare kernel for all S
hyper-samples, scores rows ``[lo, hi)`` of ``X_design`` (each candidate's sum over hyper-samples stays inside one
reduction thread, so scores are bit-identical for any world size), and the ranks exchange ``(best score, best
global index)`` -- 16 bytes each -- with one ``all_gather`` over NCCL/NVLink to pick the next design with
``np.argmax`` semantics (first index of the maximum; NaN counts as the maximum).  The full score vector is
gathered only when the caller needs it (``kld_all`` row of ``optimize``).
"""
from __future__ import annotations

import math

import numpy as np
import torch


def shard_bounds(n_items, world_size, rank):
    """Rows [lo, hi) of rank ``rank``: contiguous blocks of ceil(n/world) rows (the last ranks may be short/empty)."""
    per = (n_items + world_size - 1) // world_size
    lo = min(n_items, rank * per)
    hi = min(n_items, lo + per)
    return lo, hi


def combine_argmax(vals, idxs):
    """Reduce per-rank (best value, best global index) pairs with np.argmax semantics.

    vals: float64 tensor (W,), idxs: int64 tensor (W,); ranks without candidates report idx < 0.
    NaN counts as the largest value (lowest NaN index wins); ties go to the lowest index.
    """
    best_v, best_i = None, -1
    for v, i in zip(vals.tolist(), idxs.tolist()):
        if i < 0:
            continue
        if best_i < 0:
            best_v, best_i = v, i
            continue
        vn, bn = math.isnan(v), math.isnan(best_v)
        if vn or bn:
            take = (vn and bn and i < best_i) or (vn and not bn)
        else:
            take = v > best_v or (v == best_v and i < best_i)
        if take:
            best_v, best_i = v, i
    return best_v, best_i


def pack_record(best_val, best_idx):
    """(float64 value, int64 index) -> int64[2] tensor: the 16-byte record the ranks exchange."""
    return torch.cat([best_val.reshape(1).to(torch.float64).view(torch.int64), best_idx.reshape(1).to(torch.int64)])


def exchange_best(best_val, best_idx, group=None):
    """torch.distributed form of the exchange (gloo on CPU in the tests; the GPU path uses the C-ABI
    ``bode_ekld_argmax_nccl`` through ``EkldEngine.exchange_best``): ONE all_gather of the packed 16-byte record, one
    copy to the host, np.argmax semantics."""
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    rec = pack_record(best_val, best_idx)
    out = torch.empty(2 * ws, dtype=torch.int64, device=rec.device)
    dist.all_gather_into_tensor(out, rec, group=group)
    host = out.cpu()
    return combine_argmax(host[0::2].contiguous().view(torch.float64), host[1::2].contiguous())


def gather_rows(local, n_total, group=None):
    """all_gather equally padded 1-D shards back into a length-n_total vector (shards follow shard_bounds)."""
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    per = (n_total + ws - 1) // ws
    pad = torch.full((per,), float("nan"), dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad, group=group)
    return torch.cat(out)[:n_total]


def sharded_score(engine, X_design, form=0, group=None, want_scores=True, want_var=True, keep_planes=False, rescore=False):
    """Score ``X_design`` (host array, identical on every rank) on this rank's shard and combine across ranks.

    Shards only when a process group is passed explicitly (``torch.distributed.group.WORLD`` for the default one): a
    caller that merely runs inside an initialised job is not silently split.  Returns dict(score (Nd,) NumPy or None,
    var, idx_best (global), best).  ``keep_planes`` (with ``want_var=False``) leaves the engine able to ``rescore``;
    ``rescore=True`` evaluates ``form`` from the planes of such a preceding call on the same design (reduction kernel only).
    """

    def run(xd, off):
        if rescore:
            return engine.rescore(idx_offset=off, form=form, want_var=want_var)
        if keep_planes:
            return engine.score(xd, idx_offset=off, form=form, want_var=False, want_logk=False)
        return engine.score(xd, idx_offset=off, form=form, want_var=want_var)

    import torch.distributed as dist
    Xd = np.ascontiguousarray(X_design, dtype=np.float64)
    Nd = Xd.shape[0]
    if group is None or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        r = run(Xd, 0)
        if r.get("rec") is not None:
            rec = r["rec"].cpu()  # (best, best_idx) in one device->host read
            best, idx = float(rec[:1].view(torch.float64).item()), int(rec[1].item())
        else:
            best, idx = float(r["best"].item()), int(r["best_idx"].item())
        return dict(score=r["score"].cpu().numpy() if want_scores else None,
                    var=r["var"].cpu().numpy() if (want_scores and r.get("var") is not None) else None, idx_best=idx, best=best)
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(Nd, ws, rank)
    dev = engine.device
    use_capi = bool(getattr(engine, "has_exchange", lambda: getattr(engine, "_comm", None) is not None)())  # C-ABI exchange on the compute stream
    if hi > lo:
        r = run(Xd[lo:hi], lo)
        bv, bi, sc, vr, rec = r["best"], r["best_idx"], r["score"], r["var"], r.get("rec")
    else:
        bv = torch.full((1,), float("-inf"), dtype=torch.float64, device=dev)
        bi = torch.full((1,), -1, dtype=torch.int64, device=dev)
        sc = torch.empty(0, dtype=torch.float64, device=dev)
        vr = torch.empty(0, dtype=torch.float64, device=dev)
        rec = None
    best, idx = engine.exchange_best(rec) if use_capi else exchange_best(bv, bi, group)
    out = dict(idx_best=int(idx), best=float(best), score=None, var=None)
    if want_scores:
        out["score"] = gather_rows(sc, Nd, group).cpu().numpy()
        if want_var and vr is not None:
            out["var"] = gather_rows(vr, Nd, group).cpu().numpy()
    return out


def broadcast_array(a, shape_hint=None, group=None, src=0, device=None):
    """Rank ``src``'s float64 array on every rank (host MCMC, LHS designs and noisy objective values are drawn on one
    rank only so that the ranks of a sharded ``KLSampler.optimize`` can never diverge)."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    dev = device if (device is not None and dist.get_backend(group) == "nccl") else torch.device("cpu")
    gsrc = dist.get_global_rank(group, src) if group is not None and group is not dist.group.WORLD else src
    if rank == src:
        arr = np.ascontiguousarray(a, dtype=np.float64)
        meta = torch.tensor([arr.ndim] + list(arr.shape) + [0] * (7 - arr.ndim), dtype=torch.int64, device=dev)
    else:
        meta = torch.zeros(8, dtype=torch.int64, device=dev)
    dist.broadcast(meta, src=gsrc, group=group)
    m = meta.cpu().tolist()
    shape = tuple(m[1:1 + m[0]])
    if rank == src:
        t = torch.from_numpy(arr).to(dev)
    else:
        t = torch.empty(shape, dtype=torch.float64, device=dev)
    dist.broadcast(t, src=gsrc, group=group)
    return t.cpu().numpy()
