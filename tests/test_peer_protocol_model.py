"""A host model of the peer-memory collectives' protocol (csrc/kernels/peer.cu): N threads play the
ranks, numpy arrays play the mailboxes, every rank runs the same sequence of all-reduces and slab halo
updates with random delays.  What is checked is the logic the CUDA kernels rely on -- two parities per
slot and monotone sequence flags are enough for ranks that drift by up to one exchange, without any
barrier -- for 8 ranks (the N the GPU tests do not reach: they run 2 and 4 ranks).  Memory ordering is
not modelled (Python is sequentially consistent); the kernels use system-scope fences for that."""
import random
import threading
import time

import numpy as np
import pytest

WORDS = 8
STOP = threading.Event()      # releases ranks that are still spinning once a run has been judged


def spin():
    if STOP.is_set():
        raise RuntimeError("run abandoned")
    time.sleep(0)


class Mailbox:
    def __init__(self, size, cap):
        self.ar_val = np.zeros((2, size, WORDS))
        self.ar_flag = np.zeros((2, size), dtype=np.int64)
        self.halo_flag = np.zeros((2, 2), dtype=np.int64)       # [slot][parity]
        self.halo_data = np.zeros((2, 2, cap))


def allreduce(boxes, rank, data, seq, jitter, done, parities=2):
    """done[p] = last all-reduce rank p has finished reading.  Writing exchange `seq` into p's mailbox
    reuses the slot of exchange seq - parities: p must be past it (the invariant the parities buy)."""
    size, par = len(boxes), seq % parities
    for p in range(size):                                        # one "thread" per peer in the kernel
        assert done[p] >= seq - parities, ("slot reused before it was read", rank, p, seq, done[p])
        boxes[p].ar_val[par, rank, :data.size] = data
        jitter()
        boxes[p].ar_flag[par, rank] = seq                        # published after the data
    for p in range(size):
        while boxes[rank].ar_flag[par, p] != seq:
            spin()
    out = sum(boxes[rank].ar_val[par, q, :data.size].copy() for q in range(size))   # rank order
    done[rank] = seq
    return out


def halo(boxes, rank, x, plane, seq, jitter, done, parities=2):
    """x = [ghost_lo | owned ... | ghost_hi] in planes of `plane` doubles"""
    size, par = len(boxes), seq % parities
    peers = [p for p in (rank - 1, rank + 1) if 0 <= p < size]
    for p in peers:
        assert done[p] >= seq - parities, ("halo slot reused before it was read", rank, p, seq, done[p])
        remote_slot = 0 if rank < p else 1                       # receiver's slot for its lower / higher neighbour
        src = x[-2 * plane:-plane] if p > rank else x[plane:2 * plane]
        boxes[p].halo_data[remote_slot, par, :plane] = src
        jitter()
    for p in peers:
        boxes[p].halo_flag[0 if rank < p else 1, par] = seq      # the last block publishes
    for p in peers:
        local_slot = 0 if p < rank else 1
        while boxes[rank].halo_flag[local_slot, par] != seq:
            spin()
        dst = x[:plane] if p < rank else x[-plane:]
        dst[:] = boxes[rank].halo_data[local_slot, par, :plane]
    done[rank] = seq


def run_ranks(size, parities, rounds=60, budget=60.0):
    plane = 5
    boxes = [Mailbox(size, plane) for _ in range(size)]
    errors = []
    done_ar, done_halo = [0] * size, [0] * size
    STOP.clear()

    def rank_main(rank):
        rng = random.Random(rank)

        def jitter():
            # rank 0 is the straggler: it stalls between publishing and reading, the others run ahead
            if rank == 0 and rng.random() < 0.5:
                time.sleep(3e-4)
            elif rng.random() < 0.3:
                time.sleep(rng.random() * 2e-4)
        try:
            x = np.zeros(4 * plane)                              # ghost, two owned planes, ghost
            ar_seq = halo_seq = 0
            for it in range(rounds):
                x[plane:-plane] = 1000.0 * it + 10.0 * rank + np.arange(2 * plane)
                halo_seq += 1
                halo(boxes, rank, x, plane, halo_seq, jitter, done_halo, parities)
                if rank > 0:
                    want = 1000.0 * it + 10.0 * (rank - 1) + plane + np.arange(plane)    # neighbour's last owned plane
                    assert np.array_equal(x[:plane], want), ("halo lo", rank, it)
                if rank < size - 1:
                    want = 1000.0 * it + 10.0 * (rank + 1) + np.arange(plane)            # neighbour's first owned plane
                    assert np.array_equal(x[-plane:], want), ("halo hi", rank, it)
                for n in (1, 2):                                 # the two all-reduces of a half step
                    ar_seq += 1
                    got = allreduce(boxes, rank, np.full(n, float(it * size + rank)), ar_seq, jitter, done_ar, parities)
                    assert np.array_equal(got, np.full(n, float(sum(it * size + r for r in range(size))))), ("ar", rank, it)
                jitter()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=rank_main, args=(r,), daemon=True) for r in range(size)]
    for t in threads:
        t.start()
    deadline = time.time() + budget
    for t in threads:
        t.join(timeout=max(0.1, deadline - time.time()))
    stuck = any(t.is_alive() for t in threads)
    STOP.set()
    for t in threads:
        t.join(timeout=5)
    return [e for e in errors if "run abandoned" not in e], stuck


@pytest.mark.parametrize("size", [2, 3, 8])
def test_two_parities_suffice_without_a_barrier(size):
    errors, stuck = run_ranks(size, parities=2)
    assert not errors, errors[:3]
    assert not stuck, "protocol deadlocked"


def test_a_single_parity_would_not(size=4):
    """Negative control: with one slot per peer a rank that runs ahead overwrites data (or a flag) its
    straggling peer has not consumed yet -- the model notices (invariant, wrong sums, or a hang)."""
    errors, stuck = run_ranks(size, parities=1, rounds=60, budget=8.0)
    assert errors or stuck
