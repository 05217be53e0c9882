#!/usr/bin/env python
"""bench.py -- EKLD candidate-evaluations/sec on the BASELINE.json configurations.

One "step" = one pass of the hot path over one batch of synthetic input: prepare (batched Cholesky of all S hyper-samples
on the tensor pipe) + dense scoring of every candidate of this rank's shard + closed-form EKLD, in-order reduction over
hyper-samples and argmax (+ the 16-byte NCCL exchange when N > 1).  Prints ONE JSON line (rank 0).

  python bench.py [--config c2|c3|c4|c4quad|c5] [--gpus N --steps K --warmup W]  # our arm (default config: c3)
  python bench.py --impl reference [--config ...] [--gpus N ...]                 # CPU arm: the oracle port of the reference path
  torchrun --nproc-per-node N ... bench.py --gpus N ...                          # N > 1: one rank per GPU, candidate blocks sharded

Configurations (BASELINE.json `configs`)
  c2      d=2  n=30  S=32  |X_design|=1e4, closed-form EKLD: launch-latency / transcendental bound, report the absolute time
  c3      d=8  n=200 S=64  |X_design|=1e5 TOTAL, closed-form EKLD (the headline; "scaling": "strong" -- the design is fixed and
          split over the ranks; a weak-scaling side figure, 1e5 candidates per rank, is attached when N > 1)
  c4      d=10 n=500 S=128 |X_design|=1e6 total, closed-form EKLD (log1p form: at 1.28e8 evaluations the literal expression's
          2e-16 noise floor turns a handful of entries into NaN)
  c4quad  c4 with the kernel integrals replaced by means over N_q = 1e5 quadrature points (north-star quadrature mode)
  c5      the full sequential loop (KLSampler.optimize, max_it=50) on the 6-D function of tests/samp_ex4.py:28-39, n grows
          20 -> 70, S=16, |X_design|=1e4; a step is one whole 50-iteration loop; hyper-samples come from a seeded generator
          shared with the CPU oracle loop so that the chosen design sequences can be compared index by index

Contract notes
  value      whole-job evals/s (= |X_design| * S / time; c5: 50 * |X_design| * S / loop time) with X_design resident in HBM
  e2e        the same metric through the public host-buffer path: every step copies X, Y, hyp and the rank's X_design shard
             from pinned host memory and reads the score shard + the winner back
  roofline   dominant kernel k_score: algorithmic FP64 FLOP (F = n^2 + n(3d+5) + 8d + 20 per eval) / CUDA-event
             duration on the launching stream, against the FP64 peak MEASURED IN THIS RUN by a register-resident
             DMMA.8x8x4 loop (MEASURED_PEAKS.json has no FP64 entry; its HBM figure is quoted too); `traffic` is read from
             the newest `profiles/*k_score*ncu_raw.txt` (ncu --set full of this command)
  cpu_baseline  oracle/ekld_oracle.py (kind "port": the reference is pure Python on GPy and cannot run here) on a bounded
             sample of the same workload, one single-threaded-BLAS worker process per host core (A: vectorised route);
             `literal` is baseline B, the reference-structured per-candidate loop (`mcmc_kld`, re-factorising per evaluation)
  parity     the GPU scores of this very run against the oracle scores the cpu_baseline leg just computed
"""
import argparse
import glob
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ekld_candidate_evals_per_sec"
UNIT = "evals/s"
CONFIGS = {
    "c2": dict(d=2, n=30, S=32, Nd=10000, Nq=0, seed=1263, nugget=1e-3, form=0, cpu_cand=10000, cpu_samp=32,
               name="config2: d=2 n=30 S=32 |X_design|=1e4 (closed-form EKLD, dense scoring; launch-latency bound: absolute time)"),
    "c3": dict(d=8, n=200, S=64, Nd=100000, Nq=0, seed=1223, nugget=1e-3, form=0, cpu_cand=40000, cpu_samp=64,
               name="config3: d=8 n=200 S=64 |X_design|=1e5 total (closed-form EKLD, dense scoring)"),
    "c4": dict(d=10, n=500, S=128, Nd=1000000, Nq=0, seed=1223, nugget=1e-3, form=1, cpu_cand=16000, cpu_samp=128,
               name="config4: d=10 n=500 S=128 |X_design|=1e6 total (closed-form EKLD, log1p form, dense scoring)"),
    "c4quad": dict(d=10, n=500, S=128, Nd=1000000, Nq=100000, seed=1223, nugget=1e-3, form=1, cpu_cand=64, cpu_samp=16,
                   name="config4 quadrature mode: d=10 n=500 S=128 |X_design|=1e6 total, N_q=1e5 quadrature points"),
    "c5": dict(d=6, n=20, S=16, Nd=10000, Nq=0, seed=5, nugget=1e-3, form=1, iters=50,
               name="config5: sequential loop max_it=50, 6-D function, n 20->70, S=16, |X_design|=1e4, per-iteration re-factorisation"),
}


def flops_per_eval(n, d, Nq=0):
    return n * n + n * (3 * d + 5) + 8 * d + 20 + Nq * (3 * d + 2)


def make_workload(d, n, S, Nd, seed, rank=0):
    """Synthetic inputs: observations on a random design (8-D Dette-Pepelyshev values at d = 8), candidates uniform in the
    unit cube, hyper-samples drawn from the drivers' priors (variance ~ Gamma(1, scale 2), ell ~ Exp(scale 0.7), reference
    tests/samp_ex2.py:107-108) with ell clipped to >= 0.05.  `rank` selects the candidate block of the
    weak-scaling side run; rank 0's block is the fixed design of the strong-scaling headline."""
    from tests import cases
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    Y = cases.dette8(X)[:, None] if d == 8 else (np.sin(3 * X.sum(1)) + X[:, 0] ** 2)[:, None]
    hyp = np.column_stack([rng.gamma(1.0, 2.0, S) + 0.05] + [np.maximum(rng.exponential(0.7, S), 0.05) for _ in range(d)])
    if n <= 64:  # few observations: keep the prior draws well conditioned (reject cond(A) > 1e6 -> cap ell)
        hyp[:, 1:] = np.minimum(hyp[:, 1:], 0.6)
    rng_c = np.random.default_rng(seed + 1000 * (rank + 1))
    Xd = rng_c.random((Nd, d))
    return X, Y, hyp, Xd


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append(line.strip())
                if self.stop_flag.is_set():
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = sorted(sm)[len(sm) // 4:]  # median over the samples taken under load (clocks idle low before the first kernel)
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU arm: the oracle on all host cores (one single-threaded-BLAS worker per core) -----------------------------------
_POOL_STATE = {}


def _pool_init(X, Y, Xd, hyp, nugget, quad):
    from oracle import ekld_oracle  # noqa: F401  (loads SciPy's BLAS / LAPACK before the thread limit is applied to them)
    try:
        from threadpoolctl import threadpool_limits
        _POOL_STATE["limit"] = threadpool_limits(1)
    except Exception:
        os.environ["OMP_NUM_THREADS"] = "1"
    _POOL_STATE.update(X=X, Y=Y, Xd=Xd, hyp=hyp, nugget=nugget, quad=quad)


def _pool_rows(idx):
    """EKLD rows (and mu_1, sigma_1) of the hyper-samples `idx` for the worker's candidate block."""
    from oracle import ekld_oracle as orc
    st = _POOL_STATE
    out = []
    for s in idx:
        m = orc.make_mcmc_model(st["hyp"][s], st["X"], st["Y"], st["nugget"])
        sig1, mu1, _ = orc.sample_terms(m, st["quad"])
        Xd = st["Xd"]
        # cache-sized candidate blocks (k(X, x) of a block: n x 2048 doubles): the same arithmetic, just not 64 MB temporaries
        row = np.concatenate([orc.ekld_row(m, Xd[i:i + 2048], quad=st["quad"]) for i in range(0, Xd.shape[0], 2048)])
        out.append((s, row, mu1, sig1))
    return out


class CpuOracle(object):
    """The vectorised oracle (baseline A of BASELINE.md) over a process pool: hyper-samples are dealt to the workers,
    every worker evaluates its samples for the whole candidate block with BLAS pinned to one thread."""

    def __init__(self, X, Y, Xd, hyp, nugget, quad=None, procs=None):
        import multiprocessing as mp
        self.procs = int(procs or os.cpu_count() or 1)
        self.S, self.Nd = hyp.shape[0], Xd.shape[0]
        self.procs = max(1, min(self.procs, self.S))
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.procs, initializer=_pool_init, initargs=(X, Y, Xd, hyp, nugget, quad))
        self.pool.map(_pool_rows, [[] for _ in range(self.procs)])  # start the workers before anything is timed

    def run(self):
        from oracle import ekld_oracle as orc
        t0 = time.perf_counter()
        parts = self.pool.map(_pool_rows, [list(range(p, self.S, self.procs)) for p in range(self.procs)])
        kld = np.empty((self.S, self.Nd))
        for part in parts:
            for s, row, _, _ in part:
                kld[s] = row
        with np.errstate(invalid="ignore", divide="ignore"):
            mean, var = orc.seq_mean_var(np.log(kld))
        dt = time.perf_counter() - t0
        return dict(seconds=dt, score=mean, kld=kld, argmax=int(np.argmax(mean)))

    def close(self):
        self.pool.close()
        self.pool.join()


def literal_baseline(X, Y, hyp, Xd, nugget, n_cand):
    """Baseline B: the reference-structured loop (`mcmc_kld` per candidate, re-factorising per (candidate, sample),
    `_sampler.py:489-491`); a lower bound on the true reference cost (no GPy object overhead)."""
    from oracle import ekld_oracle as orc
    t0 = time.perf_counter()
    for c in range(n_cand):
        orc.mcmc_kld_literal(Xd[c], X, Y, hyp, nugget)
    dt = time.perf_counter() - t0
    return n_cand * hyp.shape[0] / dt, n_cand * hyp.shape[0]


def quad_points(cfg):
    if not cfg["Nq"]:
        return None
    rng = np.random.default_rng(cfg["seed"] + 77)
    return rng.random((cfg["Nq"], cfg["d"])), None  # classic-LHS-like points, unit weights (samp_ex1.py:61-62)


def hyper_c5(S, d):
    def hyper(X, Y, it):
        r = np.random.default_rng(1000 + it)
        return np.column_stack([r.gamma(2.0, 1.0, S) + 0.2] + [r.uniform(0.25, 0.9, S) for _ in range(d)])
    return hyper


def c5_problem(cfg):
    from tests import cases
    rng = np.random.default_rng(cfg["seed"])
    X0 = rng.random((cfg["n"], cfg["d"]))
    Y0 = np.array([cases.ex4_func6(x) for x in X0])[:, None]
    Xd = rng.random((cfg["Nd"], cfg["d"]))
    return X0, Y0, Xd, cases.ex4_func6


def oracle_loop_c5(cfg, iters):
    """The CPU side of config 5: the oracle's dense argmax loop on the same seeded hyper-samples."""
    from oracle import ekld_oracle as orc
    X0, Y0, Xd, f = c5_problem(cfg)
    hyper = hyper_c5(cfg["S"], cfg["d"])
    X, Y = X0.copy(), Y0.copy()
    chosen, mus = [], []
    t0 = time.perf_counter()
    for it in range(iters):
        ref = orc.score(X, Y, Xd, hyper(X, Y, it), nugget=cfg["nugget"])
        chosen.append(int(ref["argmax"])); mus.append(float(ref["mu_Q"]))
        x = Xd[ref["argmax"]]
        X = np.vstack([X, x[None]]); Y = np.vstack([Y, [[f(x)]]])
    return chosen, mus, time.perf_counter() - t0


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def newest_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the k_score launch from the newest committed ncu capture."""
    best = None
    for path in glob.glob(os.path.join(ROOT, "profiles", "*k_score*ncu_raw*.txt")):
        if best is None or os.path.getmtime(path) > os.path.getmtime(best):
            best = path
    # prefer the highest round number in the name over mtime (git checkouts reset mtimes)
    named = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[0-9]*k_score*ncu_raw*.txt")))
    if named:
        best = named[-1]
    if best is None:
        return None, None
    tot = 0.0
    for line in open(best):
        m = re.match(r"\s*dram__bytes_(read|write)\.sum\s+(\S+)\s+([0-9.eE+-]+)", line)
        if m:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(2), 1.0)
            tot += float(m.group(3)) * scale
    return (int(tot) if tot else None), os.path.relpath(best, ROOT)


def run_reference(args, cfg):
    """--impl reference: the reference's own CPU path (oracle port) on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config == "c5":
        times = []
        for i in range(max(1, min(args.steps, 2))):
            chosen, mus, dt = oracle_loop_c5(cfg, cfg["iters"])
            times.append(dt)
        ms = 1e3 * float(np.mean(times))
        value = cfg["iters"] * cfg["Nd"] * cfg["S"] / (ms * 1e-3)
        sample = "the whole %d-iteration loop (|X_design|=%d, S=%d), vectorised NumPy/LAPACK oracle, BLAS threads" % (cfg["iters"], cfg["Nd"], cfg["S"])
        cores, steps = blas_threads(), len(times)
    else:
        X, Y, hyp, Xd = make_workload(cfg["d"], cfg["n"], cfg["S"], cfg["cpu_cand"], cfg["seed"])
        quad = quad_points(cfg)
        if quad is not None:
            quad = (quad[0][:2000], None)  # bounded: 2000 of the 1e5 quadrature points (the cost is linear in N_q)
        orc_pool = CpuOracle(X, Y, Xd, hyp[:cfg["cpu_samp"]], cfg["nugget"], quad)
        times = []
        for i in range(args.warmup + args.steps):
            r = orc_pool.run()
            if i >= args.warmup:
                times.append(r["seconds"])
        orc_pool.close()
        ms = 1e3 * float(np.mean(times))
        evals = cfg["cpu_cand"] * cfg["cpu_samp"]
        value = evals / (ms * 1e-3)
        if quad is not None:
            value *= 2000.0 / cfg["Nq"]  # scaled to N_q = 1e5 (stated in `sample`)
        sample = "%d candidates x %d hyper-samples per step%s, vectorised NumPy/LAPACK oracle (GPy route), one single-threaded-BLAS worker per core" % (
            cfg["cpu_cand"], cfg["cpu_samp"], "" if quad is None else " with 2000 of the 1e5 quadrature points (evals/s scaled by 2000/1e5)")
        cores, steps = orc_pool.procs, args.steps
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["name"], "step": "bounded sample: " + sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nd", type=int, default=0, help="developer: override |X_design| (e.g. one rank's shard of an 8-GPU run on one GPU)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.nd:
        cfg["Nd"] = args.nd
        cfg["cpu_cand"] = min(cfg.get("cpu_cand", args.nd), args.nd)
        cfg["name"] += " [developer override: |X_design|=%d]" % args.nd
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args, cfg)

    # the CPU oracle's worker processes are forked BEFORE CUDA / NCCL are initialised in this process
    cpu_pool = None
    if int(os.environ.get("WORLD_SIZE", "1")) == 1 and not args.no_cpu_baseline and args.config != "c5":
        Xc, Yc, hc, Xdc = make_workload(cfg["d"], cfg["n"], cfg["S"], cfg["cpu_cand"], cfg["seed"], 0)
        qc = quad_points(cfg)
        cpu_pool = CpuOracle(Xc, Yc, Xdc, hc[:cfg["cpu_samp"]], cfg["nugget"], None if qc is None else (qc[0][:2000], None))

    import torch
    import torch.distributed as tdist
    from bode_b200 import _ext, dist as bdist
    from bode_b200.engine import EkldEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: NCCL prints its version banner to stdout when the communicator is created,
        # so file descriptor 1 points at stderr while that happens
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            tdist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    eng = EkldEngine()
    if world > 1:
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            eng.init_exchange(tdist.group.WORLD)  # the C-ABI argmax exchange: peer-memory mailboxes on one node, else NCCL
            eng.exchange_best(None)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    dev = eng.device

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    if args.config == "c5":
        line = bench_loop(args, cfg, eng, world, rank, local_rank, barrier)
    else:
        line = bench_dense(args, cfg, eng, world, rank, local_rank, barrier, cpu_pool)
    if rank == 0 and line is not None:
        print(json.dumps(line))
    if world > 1:
        eng.close_nccl()
        tdist.destroy_process_group()


def bench_dense(args, cfg, eng, world, rank, local_rank, barrier, cpu_pool):
    import torch
    import torch.distributed as tdist
    from bode_b200 import _ext, dist as bdist
    dev = eng.device
    d, n, S, Nd, form = cfg["d"], cfg["n"], cfg["S"], cfg["Nd"], cfg["form"]
    nug2 = cfg["nugget"] ** 2
    X, Y, hyp, Xd_full = make_workload(d, n, S, Nd, cfg["seed"], 0)  # the fixed design, identical on every rank
    quad = quad_points(cfg)
    lo, hi = bdist.shard_bounds(Nd, world, rank)                     # strong scaling: this rank's rows of the fixed design
    Xd_host = np.ascontiguousarray(Xd_full[lo:hi])
    Xd_dev = torch.from_numpy(Xd_host).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # 256 MB > 126 MB L2
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()
    hX, hY, hH, hXd = pin(X), pin(Y), pin(hyp), pin(Xd_host)
    qdev = None if quad is None else (torch.from_numpy(quad[0]).to(dev), None)

    def pick(r):
        if world > 1:
            return eng.exchange_best(r["rec"])
        rec = r["rec"].cpu()
        return float(rec[:1].view(torch.float64).item()), int(rec[1].item())

    def step_resident(events=None, xd=None, off=None):
        # observations and hyper-samples cross PCIe every step (tiny), from pinned memory; X_design stays resident
        eng.prepare(hX.to(dev, non_blocking=True), hY.to(dev, non_blocking=True), hH.to(dev, non_blocking=True), nug2, quad=qdev)
        r = eng.score(Xd_dev if xd is None else xd, idx_offset=lo if off is None else off, form=form, want_var=False,
                      want_logk=False, events=events)
        return pick(r), r

    peak = _ext.fp64_peak()
    info = eng.prepare(X, Y, hyp, nug2, quad=qdev)
    assert int(info.abs().max()) == 0, "synthetic hyper-samples must all factorise"

    # ---- device-resident arm: the step as one captured CUDA graph (EkldEngine / StepGraph public API) ----
    from bode_b200.engine import StepGraph
    use_graph = quad is None
    sg = StepGraph(eng, hX, hY, hH, Xd_dev, nug2, idx_offset=lo, form=form) if use_graph else None

    def timed_step():
        if sg is not None:
            return sg.run(), sg.out
        return step_resident()

    for _ in range(args.warmup):
        timed_step()
        flush.zero_()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
        time.sleep(0.3)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    best = None
    for i in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the per-step event pair)
        ev[i][0].record()
        best, r_last = timed_step()
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    t_local = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(t_local, op=tdist.ReduceOp.MAX)
    ms_per_step = float(t_local.item()) / args.steps
    value = Nd * S / (ms_per_step * 1e-3)
    score_gpu = r_last["score"].cpu().numpy() if rank == 0 else None  # rank 0's shard starts at candidate 0

    # ---- the dominant kernel's own duration: event pairs the C ABI records around k_score (plain launches, L2 flushed) ----
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 5))]
    for a_, b_ in kev:
        a_.record(); b_.record()
    torch.cuda.synchronize()
    if quad is None:
        for a_b in kev:
            flush.zero_()
            bk_, _ = step_resident(events=a_b)
            assert bk_[1] == best[1]
    torch.cuda.synchronize()
    kernel_ms = [a_.elapsed_time(b_) for a_, b_ in kev]

    # ---- end-to-end arm: host buffers in, host results out, every step ----
    h_score = torch.empty(hi - lo, dtype=torch.float64).pin_memory()
    sg2 = StepGraph(eng, hX, hY, hH, hXd, nug2, idx_offset=lo, form=form) if use_graph else None

    def step_e2e():
        if sg2 is not None:
            sg2.launch()
            h_score.copy_(sg2.out["score"], non_blocking=True)
            b = sg2.winner()
            torch.cuda.synchronize()
            return b
        eng.prepare(hX.to(dev, non_blocking=True), hY.to(dev, non_blocking=True), hH.to(dev, non_blocking=True), nug2, quad=qdev)
        r = eng.score(hXd.to(dev, non_blocking=True), idx_offset=lo, form=form, want_var=False, want_logk=False)
        h_score.copy_(r["score"], non_blocking=True)
        b = pick(r)
        torch.cuda.synchronize()
        return b

    for _ in range(2):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_ms = 0.0
    for i in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        e0.record()
        b2 = step_e2e()
        e1.record()
        torch.cuda.synchronize()
        e2e_ms += e0.elapsed_time(e1)
    t2 = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(t2, op=tdist.ReduceOp.MAX)
    e2e_value = Nd * S / (float(t2.item()) / args.steps * 1e-3)
    assert b2[1] == best[1]
    clock_info = clocks.finish() if clocks else None

    # ---- weak-scaling side measurement (|X_design| per rank fixed at the single-GPU size; c3 only) ----
    weak = None
    if world > 1 and args.config == "c3":
        Xw = torch.from_numpy(make_workload(d, n, S, Nd, cfg["seed"], rank)[3]).to(dev)
        for _ in range(3):
            step_resident(xd=Xw, off=rank * Nd)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            bw, _ = step_resident(xd=Xw, off=rank * Nd)
        s1.record(); torch.cuda.synchronize()
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        tdist.all_reduce(ts, op=tdist.ReduceOp.MAX)
        weak = {"value": world * Nd * S / (float(ts.item()) / args.steps * 1e-3), "unit": UNIT, "best_idx": bw[1],
                "note": "weak scaling: |X_design|=%d per rank (%d total), no L2 flush between steps" % (Nd, world * Nd)}

    if rank != 0:
        return None
    F = flops_per_eval(n, d)
    k_ms = float(np.mean(kernel_ms)) if quad is None else None
    evals_launch = (hi - lo) * S
    peaks_file = {}
    try:
        peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    cpu = parity = None
    if cpu_pool is not None:
        cand, samp = min(cfg["cpu_cand"], Nd), min(cfg["cpu_samp"], S)
        orc_pool = cpu_pool
        res = orc_pool.run()
        orc_pool.close()
        v = cand * samp / res["seconds"] * (1.0 if quad is None else 2000.0 / cfg["Nq"])
        lit_v, lit_n = literal_baseline(X, Y, hyp, Xd_full, cfg["nugget"], 4 if n <= 200 else 1) if quad is None else (None, 0)
        cpu = {"value": v, "unit": UNIT, "cores": orc_pool.procs, "kind": "port",
               "sample": "%d candidates x %d hyper-samples of the same workload (%.1f s wall)%s; vectorised NumPy/LAPACK oracle of the "
                         "reference path (GPy route), one single-threaded-BLAS worker process per core" % (
                             cand, samp, res["seconds"], "" if quad is None else ", 2000 of the 1e5 quadrature points, evals/s scaled by 2000/1e5"),
               "literal": None if lit_v is None else {"value": lit_v, "unit": UNIT, "cores": blas_threads(), "evals": lit_n,
                                                      "note": "baseline B: reference-structured loop, re-factorising per (candidate, sample) "
                                                              "(_sampler.py:489-491), no GPy object overhead: a lower bound on the reference's cost"}}
        if samp == S and quad is None:
            parity = parity_record(X, Y, hyp, Xd_full[:cand], cfg, form, score_gpu[:cand], res, eng)
    traffic, traffic_src = newest_traffic()
    h2d = int((hX.numel() + hY.numel() + hH.numel() + hXd.numel()) * 8)
    d2h = int(h_score.numel() * 8 + 16)
    launches = (3 if quad is None else 9) * args.steps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["name"],
                   "step": "prepare (S Cholesky, tensor pipe) + k_score over the rank's candidates (sigma_x, w.k) + xik/EKLD/log + in-order "
                           "reduction over hyper-samples + argmax" + ((" + 16-byte argmax exchange over NVLink peer memory (bode_ekld_argmax_p2p, one warp)" if eng._xchg is not None
                                                                    else " + one 16-byte ncclAllGather (C ABI)") if world > 1 else ""),
                   "l2": "flushed between timed iterations (256 MB write)",
                   "launch": "the step's copies and kernels are captured once into a CUDA graph (bode_b200.engine.StepGraph) and replayed" if use_graph else "plain launches",
                   "sharding": "fixed design of %d candidates split into contiguous blocks of %d per rank" % (Nd, hi - lo),
                   "ekld_form": "reference expression (_sampler.py:480)" if form == 0 else "log1p form (analytically identical)",
                   "wall_s_timed_region": wall},
        "roofline": None if k_ms is None else {
            "bound": "tensor", "achieved": evals_launch * F / (k_ms * 1e-3) * 1e-12, "peak": peak["dmma_tflops"], "unit": "TFLOP/s",
            "frac": evals_launch * F / (k_ms * 1e-3) * 1e-12 / peak["dmma_tflops"], "traffic": traffic,
            "traffic_unit": "bytes per k_score launch at config 3, N = 1 (dram__bytes_read.sum + dram__bytes_write.sum)",
            "traffic_source": traffic_src,
            "kernel": "bode::k_score (FP64 tensor pipe, DMMA.8x8x4)", "kernel_ms": k_ms,
            "flop_per_eval": F, "evals_per_launch": evals_launch,
            "peak_source": "measured in this run: register-resident DMMA.8x8x4 loop (bode_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
            "dfma_peak_tflops": peak["dfma_tflops"], "hbm_gbs_measured_peaks_json": peaks_file.get("hbm_gbs"),
            "algorithmic_bytes_per_eval": (8 * d + 8) / S,
            "measured_to_algorithmic_traffic": None if not traffic or args.config != "c3" else traffic / ((8 * d + 8) / S * Nd * S)},
        "cpu_baseline": cpu,
        "parity": parity,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,  # k_prepare_tiles (+ xik series blocks), k_score, k_ekld_reduce (+ final argmax) per step,
                                   # replayed from one captured CUDA graph
        "clocks": clock_info,
        "best_idx": best[1],
    }
    if quad is not None:
        line["roofline"] = {"bound": "tensor", "achieved": None, "peak": peak["dmma_tflops"], "unit": "TFLOP/s", "frac": None, "traffic": None,
                            "note": "quadrature mode: the step is dominated by k_rowmean (N_q kernel evaluations per (sample, candidate)); "
                                    "pairs/s = %.3e" % ((hi - lo) * S * cfg["Nq"] * world / (ms_per_step * 1e-3))}
    if weak:
        line["weak_scaling"] = weak
    return line


def parity_record(X, Y, hyp, Xd, cfg, form, score_gpu, res, eng):
    """GPU result of this run against the oracle result of the cpu_baseline leg (tolerance model: tests/tolerance.py)."""
    import torch
    from oracle import truth
    from tests.tolerance import per_sample_err
    sub = np.random.default_rng(0).choice(Xd.shape[0], size=min(256, Xd.shape[0]), replace=False)
    t = truth.ekld_truth(X, Y, Xd[sub], hyp, nugget=cfg["nugget"])
    eng.prepare(X, Y, hyp, cfg["nugget"] ** 2)
    r = eng.score(Xd[sub], form=form, want_var=False)
    got = np.exp(r["logk"].cpu().numpy())
    eo, eg = per_sample_err(res["kld"][:, sub], t["kld"]), per_sample_err(got, t["kld"])
    trusted = eo <= 2.5e-10
    big = res["kld"][:, sub] >= 1e-3 * res["kld"][:, sub].max(axis=1, keepdims=True)
    rel = np.abs(got - res["kld"][:, sub]) / np.abs(res["kld"][:, sub])
    finite = np.isfinite(res["score"]) & np.isfinite(score_gpu)
    return {
        "checked_candidates": int(Xd.shape[0]), "hyper_samples": int(hyp.shape[0]),
        "argmax_equal": bool(int(np.nanargmax(np.where(finite, score_gpu, -np.inf))) == int(np.nanargmax(np.where(finite, res["score"], -np.inf)))),
        "max_abs_score_diff": float(np.max(np.abs(score_gpu[finite] - res["score"][finite]))),
        "trusted_frac": float(trusted.mean()),
        "max_rel_kld_trusted": float(rel[trusted][big[trusted]].max()) if trusted.any() else None,
        "t2_ok": bool((eg <= np.maximum(1e-9, 4.0 * eo)).all()),
        "gpu_max_err_vs_long_double": float(eg.max()), "oracle_max_err_vs_long_double": float(eo.max()),
        "note": "trusted = hyper-samples on which the reference route itself is within 2.5e-10 of an 80-bit evaluation (256-candidate "
                "subset); 1e-9 relative is asserted on those, everywhere the CUDA path must be at least as accurate as the reference route",
    }


def bench_loop(args, cfg, eng, world, rank, local_rank, barrier):
    """config 5: the whole KLSampler.optimize loop; a step = one 50-iteration loop."""
    import contextlib
    import io
    import torch
    import torch.distributed as tdist
    import bode_b200
    X0, Y0, Xd, f = c5_problem(cfg)
    S, d, Nd, iters = cfg["S"], cfg["d"], cfg["Nd"], cfg["iters"]
    hyper = hyper_c5(S, d)
    group = tdist.group.WORLD if world > 1 else None

    def one_loop():
        with contextlib.redirect_stdout(io.StringIO()):
            kls = bode_b200.KLSampler(X0, Y0, False, f, False, [(0, 1)] * d, mcmc_model=True, max_it=iters, kld_tol=0.0,
                                      mcmc_chains=2, mcmc_model_avg=S, mcmc_steps=100, mcmc_burn=10, mcmc_thin=5, X_design=Xd,
                                      engine=eng, hyper_samples=hyper, ekld_form=cfg["form"], process_group=group)
            out = kls.optimize(num_designs=Nd)
        return kls.chosen_idx, out[5]

    for _ in range(min(args.warmup, 2)):
        one_loop()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
        time.sleep(0.3)
    barrier()
    steps = max(1, min(args.steps, 5))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        chosen, mus = one_loop()
    ev1.record()
    barrier()
    t_local = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=eng.device)
    if world > 1:
        tdist.all_reduce(t_local, op=tdist.ReduceOp.MAX)
    ms = float(t_local.item()) / steps
    clock_info = clocks.finish() if clocks else None
    if rank != 0:
        return None
    value = iters * Nd * S / (ms * 1e-3)
    cpu = parity = None
    if world == 1 and not args.no_cpu_baseline:
        ochosen, omus, dt = oracle_loop_c5(cfg, iters)
        cpu = {"value": iters * Nd * S / dt, "unit": UNIT, "cores": blas_threads(), "kind": "port",
               "sample": "the whole %d-iteration loop on the same seeded hyper-samples (%.1f s wall), vectorised NumPy/LAPACK oracle" % (iters, dt)}
        parity = {"chosen_idx_equal": bool(list(chosen) == list(ochosen)), "iterations": iters,
                  "max_rel_mu_Q_diff": float(np.max(np.abs(np.array(mus[:iters]) - np.array(omus)) / np.abs(np.array(omus)))),
                  "gpu_loop_s": ms * 1e-3, "cpu_loop_s": dt}
    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": min(args.warmup, 2),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["name"],
                   "step": "one whole sequential loop: %d x (prepare + dense score + argmax + objective + data update), host buffers every "
                           "iteration (KLSampler.optimize)" % iters,
                   "l2": "inputs change every iteration (n grows); no flush", "hyper_samples": "seeded generator shared with the CPU oracle loop"},
        "roofline": None, "cpu_baseline": cpu, "parity": parity,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(iters * (Nd * d + S * (d + 1)) * 8), "d2h_bytes_per_step": int(iters * (Nd * 8 + 16)),
                "note": "the loop is end to end by construction: host arrays in, host scores out, every iteration"},
        "gpu_launches": 4 * iters * steps,  # k_prepare, k_mu_sigma, k_score, k_ekld_reduce per iteration
        "clocks": clock_info, "chosen_idx_head": list(map(int, chosen[:8])),
    }


if __name__ == "__main__":
    main()
