"""Host emulation of the peer-memory argmax exchange (`k_argmax_exchange`, bode_b200/csrc/capi.cu): N "ranks" run as
threads against shared mailboxes with random delays.  What the protocol promises and this checks: every rank gathers, at
every step, exactly the records its peers published AT THAT STEP -- although nothing but the mailboxes synchronises
them -- because a slot is rewritten only two epochs later, and a rank cannot get two steps ahead of a peer whose flag of
the step in between it needs.  TEST INFRASTRUCTURE (no GPU): the CUDA kernel itself is exercised by
tests/test_gpu_parity.py::test_two_ranks_over_nccl_match_single_rank_bitwise[p2p]."""
import random
import threading
import time

import pytest


class Mailbox(object):
    """2 (epoch parity) x world slots of [best, idx, epoch]; one per rank, written by everybody."""

    def __init__(self, world):
        self.slots = [[[0.0, 0, 0] for _ in range(world)] for _ in range(2)]
        self.epoch = 0  # the rank's own step counter (device memory in the kernel)


def exchange(rank, world, boxes, rec, jitter):
    """One launch of the kernel on `rank`: lane r = loop index r."""
    mine = boxes[rank]
    e = mine.epoch + 1
    par = e & 1
    for r in range(world):                       # stores into every peer's mailbox: payload first, then the epoch word
        slot = boxes[r].slots[par][rank]
        slot[0], slot[1] = rec
        jitter()
        slot[2] = e                              # (st.release.sys)
    out = []
    for r in range(world):                       # waits on its own mailbox
        slot = mine.slots[par][r]
        t0 = time.monotonic()
        while slot[2] != e:                      # (ld.acquire.sys)
            assert time.monotonic() - t0 < 20.0, "peer %d never delivered epoch %d to rank %d" % (r, e, rank)
            time.sleep(0)
        out.append((slot[0], slot[1]))
    mine.epoch = e
    return out


@pytest.mark.parametrize("world", [2, 3, 8])
def test_every_step_gathers_that_steps_records(world):
    steps = 200
    boxes = [Mailbox(world) for _ in range(world)]
    got = [[None] * steps for _ in range(world)]
    errors = []

    def run(rank):
        rng = random.Random(1000 + rank)
        jitter = lambda: time.sleep(rng.choice([0, 0, 0, 1e-5, 2e-4]))
        try:
            for t in range(steps):
                if rng.random() < 0.05:
                    time.sleep(rng.random() * 2e-3)          # a rank that falls behind
                got[rank][t] = exchange(rank, world, boxes, (float(1000 * t + rank), 7 * t + rank), jitter)
        except BaseException as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [th.start() for th in threads]
    [th.join(timeout=120) for th in threads]
    assert not errors, errors
    assert all(not th.is_alive() for th in threads)
    for rank in range(world):
        for t in range(steps):
            assert got[rank][t] == [(float(1000 * t + r), 7 * t + r) for r in range(world)], (rank, t)
