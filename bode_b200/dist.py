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


def exchange_best(best_val, best_idx, group=None):
    """all_gather the 16-byte (value, index) record of every rank and combine.  Tensors live on the caller's
    device (CUDA with NCCL; CPU with gloo in the tests)."""
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    v = best_val.reshape(1).to(torch.float64)
    i = best_idx.reshape(1).to(torch.int64)
    vs = [torch.empty_like(v) for _ in range(ws)]
    is_ = [torch.empty_like(i) for _ in range(ws)]
    dist.all_gather(vs, v, group=group)
    dist.all_gather(is_, i, group=group)
    return combine_argmax(torch.cat(vs).cpu(), torch.cat(is_).cpu())


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


def sharded_score(engine, X_design, form=0, group=None, want_scores=True):
    """Score ``X_design`` (host array, identical on every rank) on this rank's shard and combine across ranks.

    Returns dict(score (Nd,) NumPy or None, var, idx_best (global), best).  Works unsharded when
    torch.distributed is not initialised.
    """
    import torch.distributed as dist
    Xd = np.ascontiguousarray(X_design, dtype=np.float64)
    Nd = Xd.shape[0]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        r = engine.score(Xd, form=form)
        return dict(score=r["score"].cpu().numpy(), var=r["var"].cpu().numpy(), idx_best=int(r["best_idx"].item()),
                    best=float(r["best"].item()))
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(Nd, ws, rank)
    dev = engine.device
    if hi > lo:
        r = engine.score(Xd[lo:hi], idx_offset=lo, form=form)
        bv, bi, sc, vr = r["best"], r["best_idx"], r["score"], r["var"]
    else:
        bv = torch.full((1,), float("-inf"), dtype=torch.float64, device=dev)
        bi = torch.full((1,), -1, dtype=torch.int64, device=dev)
        sc = torch.empty(0, dtype=torch.float64, device=dev)
        vr = torch.empty(0, dtype=torch.float64, device=dev)
    best, idx = exchange_best(bv, bi, group)
    out = dict(idx_best=int(idx), best=float(best), score=None, var=None)
    if want_scores:
        out["score"] = gather_rows(sc, Nd, group).cpu().numpy()
        out["var"] = gather_rows(vr, Nd, group).cpu().numpy()
    return out
