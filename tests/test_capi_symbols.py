"""CPU: the C-ABI library builds, loads and exports every symbol include/bode_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "bode_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bode_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_expected_entry_points():
    syms = header_symbols()
    for s in ("bode_ekld_prepare_dev", "bode_ekld_score_dev", "bode_ekld_score_host", "bode_ekld_mu_sigma_dev",
              "bode_us_variance_dev", "bode_gp_loglik_host", "bode_ekld_workspace_bytes", "bode_fp64_peak"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for s in header_symbols():
        assert hasattr(lib, s), "missing export %s" % s
    from bode_b200 import _ext
    assert set(_ext.EXPORTS) == set(header_symbols())


def test_pure_host_entry_points(built_lib):
    from bode_b200 import _ext
    lib = _ext.load()
    assert lib.bode_version().decode().startswith("bode_b200")
    assert lib.bode_strerror(0) == b"ok" and b"unsupported" in lib.bode_strerror(-2)
    # workspace sizing: monotone in S, zero for unsupported shapes
    w1 = lib.bode_ekld_workspace_bytes(200, 8, 64)
    w2 = lib.bode_ekld_workspace_bytes(200, 8, 128)
    assert 0 < w1 < w2 and w1 % 256 == 0
    assert lib.bode_ekld_workspace_bytes(2000, 8, 4) == 0
    assert lib.bode_ekld_workspace_bytes(10, 17, 4) == 0
    assert lib.bode_ekld_score_scratch_bytes(100000, 64) >= (100000 // 256) * 16 + 8 * 64 * 100000
    # argument validation happens before any CUDA call
    z = np.zeros(4)
    p = ctypes.c_void_p(z.ctypes.data)
    assert lib.bode_ekld_prepare_dev(None, p, p, 2, 1, 1, 2, 1e-6, 1e-8, p, 0, p, None) == -1
    assert lib.bode_ekld_prepare_dev(p, p, p, 2, 1, 1, 5, 1e-6, 1e-8, p, 0, p, None) == -1
    assert lib.bode_ekld_prepare_dev(p, p, p, 2000, 1, 1, 2, 1e-6, 1e-8, p, 0, p, None) == -2
    assert lib.bode_ekld_prepare_dev(p, p, p, 2, 1, 1, 2, 1e-6, 1e-8, p, 16, p, None) == -3


def test_no_cpu_fallback_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from bode_b200 import _ext
    from bode_b200.engine import EkldEngine
    with pytest.raises(RuntimeError):
        EkldEngine()
    with pytest.raises(_ext.BodeError) as ei:
        _ext.ekld_score_host(np.random.rand(5, 2), np.random.rand(5), np.random.rand(3, 3) + 0.1, np.random.rand(7, 2), 1e-6)
    assert ei.value.status == -4
    import bode_b200
    with pytest.raises(RuntimeError):
        bode_b200.KLSampler(np.random.rand(3, 1), np.random.rand(3, 1), False, lambda x: 0.0, False, [(0, 1)], mcmc_model=True)


@pytest.mark.parametrize("n", [1, 3, 8, 9, 30, 64, 200, 256, 257, 500, 512, 513, 700, 1000, 1024])
def test_launch_plans_are_consistent(built_lib, n):
    """Host-side planning of the scoring kernel (no GPU needed): every row tile is assigned exactly once, ring chunks are
    whole pairs of k4 groups covering all groups, the shared-memory request fits a B200 CTA, work items cover S."""
    from bode_b200 import _ext
    S, Nd = 64, 100000
    p = _ext.plan_describe(n, 8, S, Nd)
    nrt = (n + 7) // 8
    assert p["nrt"] == nrt and p["ng"] == 2 * nrt
    tiles = sorted(t for row in p["tile_id"] for t in row if t >= 0)
    assert tiles == list(range(nrt))
    for row in p["tile_id"]:  # right-aligned, ascending
        live = [t for t in row if t >= 0]
        assert live == sorted(live) and row[len(row) - len(live):] == live
    loads = [sum(t + 1 for t in row if t >= 0) for row in p["tile_id"]]
    assert max(loads) <= 1.25 * (sum(loads) / len(loads)) + nrt  # LPT balance of the triangular work
    g0 = p["chunk_g0"]
    assert g0[0] == 0 and g0[-1] == 2 * nrt and all(b > a and b % 2 == 0 for a, b in zip(g0, g0[1:]))
    blocks = lambda a, b: sum(nrt - (g >> 1) for g in range(a, b)) * 32
    assert all(blocks(a, b) <= p["chunk_doubles"] for a, b in zip(g0, g0[1:]))
    assert p["smem_bytes"] <= 232448 and p["stages"] >= 2 and p["threads"] in (256, 512)
    assert p["candidates_per_cta"] == p["col_groups"] * p["col_tiles_per_warp"] * 8
    assert p["tiles"] == -(-Nd // p["candidates_per_cta"])
    _check_sample_schedule(p, S)
    assert p["items"] >= 148


def _check_sample_schedule(p, S):
    """Work items = tiles x hyper-sample chunks; the chunks (equal ones, then the tapered tail) cover 0..S once, largest first."""
    sc, nm, tail = p["samples_per_item"], p["main_chunks"], p["tail_s0"]
    chunks = [(i * sc, min(S, (i + 1) * sc)) for i in range(nm)] + list(zip(tail, tail[1:]))
    assert p["items"] == p["tiles"] * len(chunks)
    assert chunks[0][0] == 0 and chunks[-1][1] == S and all(a[1] == b[0] for a, b in zip(chunks, chunks[1:]))
    sizes = [b - a for a, b in chunks]
    assert all(x >= 1 for x in sizes) and sizes == sorted(sizes, reverse=True)


@pytest.mark.parametrize("S,Nd", [(64, 100000), (64, 12500), (60, 100000), (128, 1000000), (1, 100000), (2, 20000), (33, 9472), (64, 9471)])
def test_sample_chunks_taper_when_there_are_more_tiles_than_sms(built_lib, S, Nd):
    from bode_b200 import _ext
    p = _ext.plan_describe(200, 8, S, Nd)
    _check_sample_schedule(p, S)
    if p["tiles"] >= 148 and S > 1:   # halving chunks: the last wave of CTAs is made of one-sample items
        tail = p["tail_s0"]
        sizes = [b - a for a, b in zip(tail, tail[1:])]
        assert p["main_chunks"] == 0 and sizes[0] == (S + 1) // 2 and sizes[-1] == 1
    else:
        assert p["tail_s0"] == []


def test_launch_plan_small_problem_uses_all_sms(built_lib):
    from bode_b200 import _ext
    p = _ext.plan_describe(200, 8, 64, 1000)      # 16 tiles only: split the hyper-samples to fill the GPU
    assert p["items"] >= 148 and p["samples_per_item"] < 64
    _check_sample_schedule(p, 64)
    q = _ext.plan_describe(200, 8, 2, 100)        # tiny: cannot fill, must still be valid
    assert q["items"] == 2 * 2 and q["samples_per_item"] == 1
    _check_sample_schedule(q, 2)
    lib = _ext.load()
    import ctypes
    buf = ctypes.create_string_buffer(4096)
    assert lib.bode_ekld_plan_describe(2000, 8, 4, 10, 148, 232448, buf, 4096) == -2
    assert lib.bode_ekld_plan_describe(10, 8, 4, 10, 148, 232448, buf, 8) == -1


def test_sample_schedule_invariants_random_shapes(built_lib):
    """The work-item schedule over random (n, S, Nd): chunks cover 0..S exactly once in descending size, items = tiles x
    chunks, and a design with fewer tiles than SMs is still split into enough items to occupy most SMs (one full wave of
    two-sample items can beat two waves of one-sample items, so not necessarily all of them)."""
    from bode_b200 import _ext
    rng = np.random.default_rng(77)
    for _ in range(300):
        n = int(rng.choice([1, 7, 30, 64, 200, 256, 500, 1000]))
        S = int(rng.choice([1, 2, 3, 5, 16, 31, 60, 64, 100, 128, 257, 1000, 70000]))
        Nd = int(rng.choice([1, 63, 64, 65, 1000, 9471, 9472, 12500, 100000, 1000000]))
        p = _ext.plan_describe(n, 8, S, Nd)
        _check_sample_schedule(p, S)
        assert len(p["tail_s0"]) <= 25
        assert 2 * p["items"] >= min(148, p["tiles"] * S) or p["tiles"] >= 148, (n, S, Nd, p["items"])
