#!/usr/bin/env python
"""bench.py -- EKLD candidate-evaluations/sec on BASELINE config 3 (d=8, n=200, S=64, |X_design|=1e5 per GPU).

One "step" = one pass of the hot path over one batch of synthetic input: prepare (batched Cholesky of all S
hyper-samples) + dense scoring of every candidate + sequential reduction over hyper-samples + argmax (+ the 16-byte
NCCL exchange when N > 1).  Prints ONE JSON line (rank 0).

  python bench.py [--gpus N --steps K --warmup W]          # our arm
  python bench.py --impl reference [--gpus N ...]          # CPU arm: the oracle port of the reference path
  torchrun --nproc-per-node N ... bench.py --gpus N ...    # N > 1 (one rank per GPU, candidates sharded, weak scaling)

Contract notes
  value      whole-job evals/s (= N * |X_design| * S / time) with X_design already resident in HBM
  e2e        the same metric through the public host-buffer API (bode_b200.engine.ekld_score-style call): every
             step copies X, Y, hyp and the step's X_design from pinned host memory and reads score + argmax back
  roofline   dominant kernel k_score: algorithmic FP64 FLOP (F = n^2 + n(3d+5) + 8d + 20 per eval) /
             CUDA-event duration on the launching stream, against the FP64 peak MEASURED IN THIS RUN by a
             register-resident DMMA.8x8x4 loop (MEASURED_PEAKS.json has no FP64 entry; its HBM figure is quoted too)
  cpu_baseline  oracle/ekld_oracle.py (kind "port": the reference is pure Python on GPy and cannot run here) on a
             bounded sample of the same workload, all host cores via BLAS threads
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ekld_candidate_evals_per_sec"
UNIT = "evals/s"
WORKLOAD = dict(d=8, n=200, S=64, Nd=100000, seed=1223, nugget=1e-3)


def flops_per_eval(n, d):
    return n * n + n * (3 * d + 5) + 8 * d + 20


def make_workload(d, n, S, Nd, seed, rank=0):
    """Synthetic config-3 inputs: 8-D Dette-Pepelyshev observations on a classic LHS-like design, centred-LHS-like
    candidates, hyper-samples drawn from the drivers' priors (variance ~ Gamma(1, scale 2), ell ~ Exp(scale 0.7),
    reference tests/samp_ex2.py:107-108) with ell clipped to >= 0.05."""
    from tests import cases
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    Y = cases.dette8(X)[:, None] if d == 8 else np.sin(3 * X.sum(1))[:, None]
    hyp = np.column_stack([rng.gamma(1.0, 2.0, S) + 0.05] + [np.maximum(rng.exponential(0.7, S), 0.05) for _ in range(d)])
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
        # median over the samples taken under load (clocks idle low before the first kernel)
        busy = sorted(sm)[len(sm) // 4:]
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_evals_per_sec(X, Y, hyp, Xd, nugget, cand, samp, repeats=1):
    """Time the vectorised oracle (baseline (A) of BASELINE.md) on a bounded slice; returns (evals/s, description)."""
    from oracle import ekld_oracle as orc
    cand = min(cand, Xd.shape[0]); samp = min(samp, hyp.shape[0])
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        res = orc.score(X, Y, Xd[:cand], hyp[:samp], nugget=nugget)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    assert np.isfinite(res["score"]).all()
    return cand * samp / best, "%d candidates x %d hyper-samples of the same workload (%.1f s)" % (cand, samp, best)


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference's own CPU path (oracle port) on host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOAD
    X, Y, hyp, Xd = make_workload(w["d"], w["n"], w["S"], w["Nd"], w["seed"])
    cand, samp = 20000, 16
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        v, desc = cpu_oracle_evals_per_sec(X, Y, hyp, Xd, w["nugget"], cand, samp)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    value = cand * samp / (ms * 1e-3)
    cores = blas_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config3: d=8 n=200 S=64 |X_design|=1e5 (closed-form EKLD, dense scoring)",
                   "step": "bounded sample: %d candidates x %d hyper-samples per step" % (cand, samp)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d candidates x %d hyper-samples per step, vectorised NumPy/LAPACK oracle (GPy route)" % (cand, samp)},
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

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

    w = WORKLOAD
    d, n, S, Nd = w["d"], w["n"], w["S"], w["Nd"]
    X, Y, hyp, Xd_host = make_workload(d, n, S, Nd, w["seed"], rank)
    eng = EkldEngine()
    if world > 1:
        eng.init_nccl(tdist.group.WORLD)
    dev = eng.device
    Xd_dev = torch.from_numpy(Xd_host).to(dev)
    idx_offset = rank * Nd  # weak scaling: every rank owns |X_design| = 1e5 candidates of a N*1e5 design
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # 256 MB > 126 MB L2

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    # observations and hyper-samples cross PCIe every step (tiny), from pinned memory so that the copies
    # do not block the host; X_design stays resident
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()
    hX, hY, hH, hXd = pin(X), pin(Y), pin(hyp), pin(Xd_host)

    def step_resident(events=None):
        eng.prepare(hX.to(dev, non_blocking=True), hY.to(dev, non_blocking=True), hH.to(dev, non_blocking=True), w["nugget"] ** 2)
        r = eng.score(Xd_dev, idx_offset=idx_offset, events=events)
        if world > 1:
            return eng.exchange_best(r["rec"])
        return float(r["best"].item()), int(r["best_idx"].item())

    # ---- measured FP64 peak (roofline denominator) ----
    peak = _ext.fp64_peak()
    info = eng.prepare(X, Y, hyp, w["nugget"] ** 2)
    assert int(info.abs().max()) == 0, "synthetic hyper-samples must all factorise"

    # ---- device-resident arm ----
    for _ in range(args.warmup):
        step_resident()
        flush.zero_()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
        time.sleep(0.3)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    # event pairs the C ABI records around the scoring kernel (no synchronisation inside the timed region); recorded
    # once here so that the underlying cudaEvent_t handles exist
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a_, b_ in kev:
        a_.record(); b_.record()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    best = None
    for i in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the per-step event pair)
        ev[i][0].record()
        best = step_resident(events=kev[i])
        ev[i][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    kernel_ms = [a_.elapsed_time(b_) for a_, b_ in kev]
    t_local = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(t_local, op=tdist.ReduceOp.MAX)
    total_ms = float(t_local.item())
    ms_per_step = total_ms / args.steps
    value = world * Nd * S / (ms_per_step * 1e-3)

    # ---- end-to-end arm: host buffers in, host results out, every step ----
    h_score = torch.empty(Nd, dtype=torch.float64).pin_memory()

    def step_e2e():
        eng.prepare(hX.to(dev, non_blocking=True), hY.to(dev, non_blocking=True), hH.to(dev, non_blocking=True), w["nugget"] ** 2)
        r = eng.score(hXd.to(dev, non_blocking=True), idx_offset=idx_offset)
        h_score.copy_(r["score"], non_blocking=True)
        if world > 1:
            b = eng.exchange_best(r["rec"])
        else:
            b = (float(r["best"].item()), int(r["best_idx"].item()))
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
    e2e_value = world * Nd * S / (float(t2.item()) / args.steps * 1e-3)
    assert b2[1] == best[1]
    clock_info = clocks.finish() if clocks else None

    # ---- strong-scaling side measurement (total |X_design| = 1e5 split over the ranks) ----
    strong = None
    if world > 1:
        lo, hi = bdist.shard_bounds(Nd, world, rank)
        Xs = torch.from_numpy(make_workload(d, n, S, Nd, w["seed"], 0)[3][lo:hi]).to(dev)  # same global design on all ranks
        for _ in range(3):
            eng.prepare(X, Y, hyp, w["nugget"] ** 2); r = eng.score(Xs, idx_offset=lo); eng.exchange_best(r["rec"])
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            eng.prepare(X, Y, hyp, w["nugget"] ** 2); r = eng.score(Xs, idx_offset=lo); eng.exchange_best(r["rec"])
        s1.record(); torch.cuda.synchronize()
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        tdist.all_reduce(ts, op=tdist.ReduceOp.MAX)
        strong = {"value": Nd * S / (float(ts.item()) / args.steps * 1e-3), "unit": UNIT, "note": "fixed total |X_design|=1e5 split over ranks"}

    if rank == 0:
        F = flops_per_eval(n, d)
        k_ms = float(np.mean(kernel_ms))
        achieved = Nd * S * F / (k_ms * 1e-3) * 1e-12
        peaks_file = {}
        try:
            peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, desc = cpu_oracle_evals_per_sec(X, Y, hyp, Xd_host, w["nugget"], 20000, 64)
            cpu = {"value": v, "unit": UNIT, "cores": blas_threads(), "kind": "port",
                   "sample": desc + "; vectorised NumPy/LAPACK oracle of the reference path (GPy route)"}
        h2d = int((hX.numel() + hY.numel() + hH.numel() + hXd.numel()) * 8)
        d2h = int(h_score.numel() * 8 + 16)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config3: d=8 n=200 S=64 |X_design|=1e5 per GPU (closed-form EKLD, dense scoring)",
                       "step": "prepare (S Cholesky) + score all candidates (sigma_x, w.k) + xik/EKLD + reduce over hyper-samples + argmax" +
                               (" + NCCL all_gather of (best, idx)" if world > 1 else ""),
                       "l2": "flushed between timed iterations (256 MB write)", "sharding": "candidate blocks, %d per rank" % Nd,
                       "wall_s_timed_region": wall},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak["dmma_tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peak["dmma_tflops"], "traffic": 66816512,
                         "traffic_unit": "bytes per launch (dram__bytes_read.sum 19.43 MB + dram__bytes_write.sum 47.39 MB; part of the 102 MB of "
                                         "(sigma_x, w.k) output is still in the 126 MB L2 when the launch ends)",
                         "traffic_source": "profiles/r1_final_k_score_bench_ncu_raw.txt: ncu --set full of this command, k_score launch",
                         "kernel": "bode::k_score (FP64 tensor pipe, DMMA.8x8x4)", "kernel_ms": k_ms,
                         "flop_per_eval": F, "evals_per_launch": Nd * S,
                         "peak_source": "measured in this run: register-resident DMMA.8x8x4 loop (bode_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry",
                         "dfma_peak_tflops": peak["dfma_tflops"], "hbm_gbs_measured_peaks_json": peaks_file.get("hbm_gbs"),
                         "algorithmic_bytes_per_eval": (8 * d + 8) / S + 16},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": 3 * args.steps,  # k_prepare (+ xik series blocks), k_score, k_ekld_reduce (+ final argmax) per step
            "clocks": clock_info,
            "best_idx": best[1],
        }
        if strong:
            line["strong_scaling"] = strong
        print(json.dumps(line))
    if world > 1:
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
